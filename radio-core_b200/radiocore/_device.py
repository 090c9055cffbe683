"""Device plumbing: PyTorch owns device memory and streams, nothing else."""
import numpy as np
import torch


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("radiocore (B200): no CUDA device visible; this package has no CPU fallback")


def device_index():
    require_cuda()
    return torch.cuda.current_device()


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def is_tensor(x):
    return isinstance(x, torch.Tensor)


def to_device(x, dtype):
    """NumPy array / torch tensor / sequence -> contiguous CUDA tensor of `dtype`."""
    require_cuda()
    if isinstance(x, torch.Tensor):
        t = x
    else:
        a = np.asarray(x)
        if dtype.is_complex and not np.iscomplexobj(a):
            a = a.astype(np.complex64)
        elif dtype.is_complex:
            a = a.astype(np.complex64, copy=False)
        else:
            a = a.astype(np.float32, copy=False)
        t = torch.from_numpy(np.ascontiguousarray(a))
    if t.device.type != "cuda":
        t = t.to("cuda", non_blocking=False)
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def is_complex_input(x):
    if isinstance(x, torch.Tensor):
        return x.is_complex()
    return np.iscomplexobj(np.asarray(x))


def to_host(t):
    return t.detach().cpu().numpy()
