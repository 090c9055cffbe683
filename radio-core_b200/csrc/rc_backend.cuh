// rc_backend.cuh -- memory / launch shims.
//
// Product build (default): CUDA device memory, CUDA kernels on a stream.
// Replay build (-DRC_EMULATE, tests/native only): the same per-thread phase
// functions are run serially on the host so the index arithmetic of every
// kernel can be checked in the GPU-less build container.  The replay library
// is test infrastructure: the Python package never loads it.
#pragma once

#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rc_fft.cuh"

namespace rc {

#define RC_CHECK(expr)                                     \
    do {                                                   \
        cudaError_t _e = (expr);                           \
        if (_e != cudaSuccess) return _e;                  \
    } while (0)

#ifdef RC_EMULATE
constexpr bool kOnDevice = false;
inline cudaError_t dev_malloc(void** p, size_t n) {
    *p = calloc(1, n ? n : 1);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
inline cudaError_t dev_free(void* p) { free(p); return cudaSuccess; }
inline cudaError_t dev_copy(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) {
    memcpy(d, s, n);
    return cudaSuccess;
}
inline cudaError_t dev_zero(void* d, size_t n, cudaStream_t) { memset(d, 0, n); return cudaSuccess; }
inline cudaError_t dev_sync(cudaStream_t) { return cudaSuccess; }

template <class F> cudaError_t launch_ew(long long n, int batch, const F& f, cudaStream_t,
                                         const char* = "ew", double = 0.0, int = 0) {
    for (int b = 0; b < batch; b++)
        for (long long i = 0; i < n; i++) f(b, i);
    return cudaSuccess;
}

#else
constexpr bool kOnDevice = true;
inline cudaError_t dev_malloc(void** p, size_t n) { return cudaMalloc(p, n ? n : 1); }
inline cudaError_t dev_free(void* p) { return cudaFree(p); }
inline cudaError_t dev_copy(void* d, const void* s, size_t n, cudaMemcpyKind k, cudaStream_t st) {
    return cudaMemcpyAsync(d, s, n, k, st);
}
inline cudaError_t dev_zero(void* d, size_t n, cudaStream_t st) { return cudaMemsetAsync(d, 0, n, st); }
inline cudaError_t dev_sync(cudaStream_t st) { return cudaStreamSynchronize(st); }

template <class F> __global__ void __launch_bounds__(256) ew_kernel(const F f, long long n) {
    const int b = blockIdx.y;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        f(b, i);
}

// Elementwise launch: f(batch, i) for i in [0, n).  Grid is sized in whole
// waves of the 148 SMs (8 resident 256-thread CTAs each) and grid-strided.
// ctas_per_sm > 0 caps the grid at that many CTAs per SM (persistent grid-stride): for kernels bound
// by something other than HBM (NVLink stores) that should leave SM slots to concurrent streams.
template <class F> cudaError_t launch_ew(long long n, int batch, const F& f, cudaStream_t stream,
                                         const char* tag = "ew", double bytes = 0.0, int ctas_per_sm = 0) {
    if (n <= 0 || batch <= 0) return cudaSuccess;
    ProfileScope scope(tag, bytes, stream);
    long long blocks = (n + 255) / 256;
    long long cap = ((ctas_per_sm > 0 ? 148LL * ctas_per_sm : 148LL * 8 * 4) + batch - 1) / batch;
    if (cap < 1) cap = 1;
    if (blocks > cap) blocks = cap;
    dim3 grid((unsigned)blocks, (unsigned)batch);
    ew_kernel<F><<<grid, 256, 0, stream>>>(f, n);
    return cudaGetLastError();
}
#endif

}  // namespace rc
