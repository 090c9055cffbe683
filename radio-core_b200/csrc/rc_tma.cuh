// rc_tma.cuh -- Tensor Memory Accelerator plumbing for the FFT pass kernels:
// tensor-map encoding on the host (driver entry point fetched through the
// runtime, so the library does not link libcuda) and the mbarrier / bulk-tensor
// PTX used on the device.  A pass stages its [R rows x 16 columns] complex64
// tile with a handful of cp.async.bulk.tensor loads that complete on one
// mbarrier; no thread computes a global load address.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace rc {

// Geometry of one tile source: element (row t, column j, batch b) of a complex64
// array lives at base + b*batch_stride + t*row_stride + j (in float2 units).
struct TileSource {
    const float2* base;
    long long row_stride;     // n / R
    long long batch_stride;
    long long cols;           // number of valid columns (== row_stride for a pass)
    int rows;                 // R
    int batch;
};

constexpr int kTmaMaxBoxRows = 256;

// Largest divisor of R that fits one TMA box.
inline int tma_box_rows(int R) {
    for (int d = kTmaMaxBoxRows; d >= 1; d--)
        if (R % d == 0) return d;
    return 1;
}

#ifndef RC_EMULATE
typedef CUresult (*rc_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                       const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                       CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline rc_encode_tiled_fn tma_encode_fn() {
    static rc_encode_tiled_fn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
        return (rc_encode_tiled_fn)p;
    }();
    return fn;
}

// true when `src` can be described by a tensor map (alignment rules of cuTensorMapEncodeTiled)
inline bool tma_source_ok(const TileSource& s) {
    return tma_encode_fn() != nullptr && ((uintptr_t)s.base % 16 == 0) && (s.row_stride % 2 == 0) &&
           (s.batch_stride % 2 == 0) && s.rows >= 1 && s.cols >= 1;
}

// rank-3 map over float32 pairs: dim0 = 2*cols floats (contiguous), dim1 = rows, dim2 = batch;
// box = {2*tile_cols floats, box_rows, 1}.  Columns past `cols` read as zero.
inline bool tma_encode_tile_map(CUtensorMap* map, const TileSource& s, int box_rows, int tile_cols) {
    cuuint64_t dims[3] = {(cuuint64_t)s.cols * 2, (cuuint64_t)s.rows, (cuuint64_t)(s.batch > 0 ? s.batch : 1)};
    cuuint64_t strides[2] = {(cuuint64_t)s.row_stride * 8, (cuuint64_t)(s.batch_stride > 0 ? s.batch_stride : s.row_stride * s.rows) * 8};
    cuuint32_t box[3] = {(cuuint32_t)(2 * tile_cols), (cuuint32_t)box_rows, 1u};
    cuuint32_t es[3] = {1u, 1u, 1u};
    CUresult r = tma_encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)s.base, dims, strides, box, es,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// rank-2 map over float32: dim0 contiguous, dim1 at `row_stride_bytes` (a multiple of 16)
inline bool tma_encode_2d_f32(CUtensorMap* map, const void* base, unsigned long long dim0, unsigned long long dim1,
                              unsigned long long row_stride_bytes, unsigned box0, unsigned box1) {
    if (tma_encode_fn() == nullptr || ((uintptr_t)base % 16) != 0 || (row_stride_bytes % 16) != 0 || box0 > 256 || box1 > 256 ||
        (box0 * 4) % 16 != 0) return false;
    cuuint64_t dims[2] = {dim0, dim1};
    cuuint64_t strides[1] = {row_stride_bytes};
    cuuint32_t box[2] = {box0, box1};
    cuuint32_t es[2] = {1u, 1u};
    CUresult r = tma_encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, es,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}
#endif

#if defined(__CUDACC__) && !defined(RC_EMULATE)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int z) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
#endif

}  // namespace rc
