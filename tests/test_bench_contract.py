"""The JSON line `bench.py` prints (driver contract): checked here on the CPU arm with a small
workload -- one line on stdout, the required keys, the reference package as the thing timed.  The
B200 arm prints the same keys plus `roofline`, `kernels`, `clocks` (profiles/r02_final_*.json)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

REQUIRED = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline", "gpu_launches"]


def test_reference_arm_line():
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "small",
                           "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [ln for ln in proc.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines                      # ONE JSON line on stdout
    d = json.loads(lines[0])
    for k in REQUIRED:
        assert k in d, k
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert set(d["config"]) == {"workload", "demodulator"}          # the same object the B200 arm prints
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"]
    assert cb["single_core"]["cores"] == 1 and cb["single_core"]["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    if os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "radiocore")) or os.path.isdir("/root/reference/radiocore"):
        assert cb["kind"] == "reference"               # the reference package itself, not the port


def test_b200_arm_refuses_without_gpu():
    import torch
    if torch.cuda.is_available():
        return
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "small"], capture_output=True,
                          text=True, timeout=300, cwd=ROOT)
    assert proc.returncode != 0 and "no CUDA device" in (proc.stderr + proc.stdout)    # no CPU fallback
