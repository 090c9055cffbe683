// membench.cu -- data-movement skeletons of one FFT pass (no math) on B200:
// how fast can a [R rows x T columns] strided tile be staged through shared
// memory and written back?  Decides tile width / pipelining of the pass kernels.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o membench tools/membench.cu
//   ./membench            (prints GB/s per variant; read+write bytes / time)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

// ---------------------------------------------------------------- plain LDG/STG tile copy
template <int T, int NT, int ROWS_PER_THREAD, bool L2P = false>
__global__ void __launch_bounds__(NT) ldg_tile_copy(const float2* __restrict__ in, float2* __restrict__ out,
                                                   long long stride, int R) {
    extern __shared__ float2 sm[];
    const int c = threadIdx.x % T, r0 = threadIdx.x / T;
    const long long j = (long long)blockIdx.x * T + c;
    constexpr int RS = NT / T;
    for (int base = 0; base < R; base += RS * ROWS_PER_THREAD) {
        float2 v[ROWS_PER_THREAD];
#pragma unroll
        for (int i = 0; i < ROWS_PER_THREAD; i++) {
            int t = base + r0 + i * RS;
            if (t < R) {
                if (L2P) asm volatile("ld.global.nc.L2::256B.v2.f32 {%0, %1}, [%2];" : "=f"(v[i].x), "=f"(v[i].y) : "l"(in + j + (long long)t * stride));
                else v[i] = __ldg(in + j + (long long)t * stride);
            }
        }
#pragma unroll
        for (int i = 0; i < ROWS_PER_THREAD; i++) {
            int t = base + r0 + i * RS;
            if (t < R) sm[t * T + c] = v[i];
        }
    }
    __syncthreads();
    for (int t = r0; t < R; t += RS) {
        float2 v = sm[t * T + c];
        out[j + (long long)t * stride] = make_float2(v.x + 1.0f, v.y);
    }
}

// ---------------------------------------------------------------- TMA-staged persistent tile copy
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
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
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
}

template <int T, int NT, int NBUF, int BOXROWS>
__global__ void __launch_bounds__(NT) tma_tile_copy(const __grid_constant__ CUtensorMap map, float2* __restrict__ out,
                                                   long long stride, int R, int ntiles, int mode) {
    extern __shared__ __align__(128) unsigned char smraw[];
    float2* bufs = (float2*)smraw;
    uint64_t* bars = (uint64_t*)(smraw + (size_t)NBUF * R * T * sizeof(float2));
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int b = 0; b < NBUF; b++) mbar_init(&bars[b], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t tile_bytes = (uint32_t)R * T * sizeof(float2);
    auto issue = [&](int tile, int b) {
        mbar_expect_tx(&bars[b], tile_bytes);
        for (int r = 0; r < R; r += BOXROWS)
            tma_load_2d(bufs + (size_t)b * R * T + (size_t)r * T, &map, &bars[b], tile * T * 2, r);
    };
    int my = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) my++;
    if (tid == 0 && (mode & 3) != 2)
        for (int p = 0; p < NBUF - 1 && p < my; p++) issue(blockIdx.x + p * gridDim.x, p);
    const int c = tid % T, r0 = tid / T;
    constexpr int RS = NT / T;
    for (int i = 0; i < my; i++) {
        const int tile = blockIdx.x + i * gridDim.x;
        const int b = i % NBUF;
        // the buffer refilled now was consumed in iteration i-1 (all threads passed the barrier below)
        if (mode & 4) {
            asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
        }
        if ((mode & 3) != 2) {
            if (tid == 0 && i + NBUF - 1 < my) issue(blockIdx.x + (i + NBUF - 1) * gridDim.x, (i + NBUF - 1) % NBUF);
            mbar_wait(&bars[b], (i / NBUF) & 1);
        }
        const float2* sm = bufs + (size_t)b * R * T;
        const long long j = (long long)tile * T + c;
        if ((mode & 3) == 1) {
            float acc = 0.f;
            for (int t = r0; t < R; t += RS) acc += sm[t * T + c].x;
            if (acc == 123.456f) out[j] = make_float2(acc, acc);
        } else {
#pragma unroll 4
            for (int t = r0; t < R; t += RS) {
                float2 v = (mode & 3) == 2 ? make_float2((float)t, (float)c) : sm[t * T + c];
                out[j + (long long)t * stride] = make_float2(v.x + 1.0f, v.y);
            }
        }
        __syncthreads();
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn get_encode() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    return (EncodeFn)fn;
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; CK(cudaEventElapsedTime(&ms, a, b)); return ms; }

template <int T, int NT, int RPT, bool L2P = false>
void run_ldg(const char* name, const float2* in, float2* out, long long n, int R, int iters) {
    const long long stride = n / R;
    const int grid = (int)(stride / T);
    size_t smem = (size_t)R * T * sizeof(float2);
    CK(cudaFuncSetAttribute(ldg_tile_copy<T, NT, RPT, L2P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    ldg_tile_copy<T, NT, RPT, L2P><<<grid, NT, smem>>>(in, out, stride, R);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(a);
    for (int i = 0; i < iters; i++) ldg_tile_copy<T, NT, RPT, L2P><<<grid, NT, smem>>>(in, out, stride, R);
    cudaEventRecord(b); CK(cudaDeviceSynchronize());
    float ms = time_ms(a, b) / iters;
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ldg_tile_copy<T, NT, RPT, L2P>, NT, smem);
    printf("%-34s R=%4d T=%2d NT=%3d occ=%d  %.3f ms  %.0f GB/s\n", name, R, T, NT, occ, ms, 16.0 * n / ms / 1e6);
}

template <int T, int NT, int NBUF, int BOXROWS>
void run_tma(const char* name, const float2* in, float2* out, long long n, int R, int iters, int ctas_per_sm, int promo = 0, int mode = 0) {
    static EncodeFn enc = get_encode();
    const long long stride = n / R;
    CUtensorMap map;
    cuuint64_t dims[2] = {(cuuint64_t)stride * 2, (cuuint64_t)R};
    cuuint64_t strides[1] = {(cuuint64_t)stride * 8};
    cuuint32_t box[2] = {(cuuint32_t)T * 2, (cuuint32_t)BOXROWS};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)in, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", name, (int)r); return; }
    if (R % BOXROWS) { printf("%s: R %% BOXROWS\n", name); return; }
    const int ntiles = (int)(stride / T);
    size_t smem = (size_t)NBUF * R * T * sizeof(float2) + 64;
    CK(cudaFuncSetAttribute(tma_tile_copy<T, NT, NBUF, BOXROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = 148 * ctas_per_sm;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (mode & 4) ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, tma_tile_copy<T, NT, NBUF, BOXROWS>, map, out, stride, R, ntiles, mode));
    CK(cudaDeviceSynchronize());
    cudaEventRecord(a);
    for (int i = 0; i < iters; i++) CK(cudaLaunchKernelEx(&cfg, tma_tile_copy<T, NT, NBUF, BOXROWS>, map, out, stride, R, ntiles, mode));
    cudaEventRecord(b); CK(cudaDeviceSynchronize());
    float ms = time_ms(a, b) / iters;
    printf("%-34s R=%4d T=%2d NT=%3d nbuf=%d cta/sm=%d promo=%d mode=%d  %.3f ms  %.0f GB/s\n", name, R, T, NT, NBUF, ctas_per_sm, promo, mode, ms,
           ((mode & 3) == 0 ? 16.0 : 8.0) * n / ms / 1e6);
}

__global__ void plain_copy(const float4* __restrict__ in, float4* __restrict__ out, long long n4) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) out[i] = in[i];
}

int main() {
    const long long n = 256000000LL;
    float2 *in, *out;
    CK(cudaMalloc(&in, n * sizeof(float2)));
    CK(cudaMalloc(&out, n * sizeof(float2)));
    CK(cudaMemset(in, 0, n * sizeof(float2)));
    CK(cudaMemset(out, 0, n * sizeof(float2)));
    const int iters = 5;
    {
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        plain_copy<<<148 * 8, 512>>>((const float4*)in, (float4*)out, n / 2);
        cudaEventRecord(a);
        for (int i = 0; i < iters; i++) plain_copy<<<148 * 8, 512>>>((const float4*)in, (float4*)out, n / 2);
        cudaEventRecord(b); CK(cudaDeviceSynchronize());
        printf("%-34s %.3f ms  %.0f GB/s\n", "plain float4 copy", time_ms(a, b) / iters, 16.0 * n / (time_ms(a, b) / iters) / 1e6);
        cudaEventRecord(a);
        for (int i = 0; i < iters; i++) CK(cudaMemcpyAsync(out, in, n * 8, cudaMemcpyDeviceToDevice));
        cudaEventRecord(b); CK(cudaDeviceSynchronize());
        printf("%-34s %.3f ms  %.0f GB/s\n", "cudaMemcpy D2D", time_ms(a, b) / iters, 16.0 * n / (time_ms(a, b) / iters) / 1e6);
    }
    run_tma<16, 256, 2, 128>("tma T16 2buf", in, out, n, 640, iters, 1, 0, 0);
    run_tma<16, 256, 2, 128>("tma T16 2buf cluster2 lockstep", in, out, n, 640, iters, 1, 0, 4);
    run_tma<16, 256, 2, 128>("tma T16 read-only cluster2", in, out, n, 640, iters, 1, 0, 5);
    run_tma<16, 256, 2, 128>("T16 write-only cluster2", in, out, n, 640, iters, 1, 0, 6);
    run_tma<16, 256, 1, 128>("tma T16 1buf x2cta cluster2", in, out, n, 640, iters, 2, 0, 4);
    run_tma<32, 512, 1, 128>("tma T32 1buf R640", in, out, n, 640, iters, 1, 0, 0);
    run_tma<32, 512, 1, 128>("tma T32 1buf R640 read-only", in, out, n, 640, iters, 1, 0, 1);
    run_tma<32, 512, 2, 80>("tma T32 2buf R320", in, out, n, 320, iters, 1, 0, 0);
    run_tma<64, 512, 2, 80>("tma T64 2buf R160", in, out, n, 160, iters, 1, 0, 0);
    run_tma<64, 512, 2, 80>("tma T64 2buf R160 read-only", in, out, n, 160, iters, 1, 0, 1);
    run_tma<16, 512, 1, 125>("tma T16 1buf R1000", in, out, n, 1000, iters, 1, 0, 0);
    return 0;
}
