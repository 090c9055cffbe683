// rc_fft2_inst.cuh -- kernels and launchers of the register-radix passes for one
// group of schedules.  Included by rc_fft2_g<k>.cu with RC_V2_GROUP_LIST set to
// the group's X-macro list (rc_fft.cuh), so the groups compile in parallel.
#pragma once

#include "rc_exec.cuh"
#include "rc_fft2.cuh"

namespace rc {

#if defined(__CUDACC__) && !defined(RC_EMULATE)
// first pass of a plan: fused LoadOp, column runs re-ordered through shared memory
template <class S, int SIGN>
__global__ void __launch_bounds__(S::NT, S::MINB) v2_first_kernel(const FftPass P, const LoadAny ld, const StoreC64 st) {
    extern __shared__ float2 rc_v2_smem[];
    float2* sm = rc_v2_smem;
    float2* tw = rc_v2_smem + S::R * V2Smem<S, true>::PITCH;
    const int batch = blockIdx.y + blockIdx.z * gridDim.y;
    const long long j0 = (long long)blockIdx.x * kV2T;
    const int tid = threadIdx.x;
    v2_load_table<S, SIGN>(tw, P, tid);
    __syncthreads();
    switch (ld.kind) {
        case kLdGather: v2_stage0<S, SIGN, true>(sm, tw, P, ld.gather, batch, j0, tid); break;
        case kLdDisc: v2_stage0<S, SIGN, true>(sm, tw, P, ld.disc, batch, j0, tid); break;
        default: v2_stage0<S, SIGN, true>(sm, tw, P, ld.c64, batch, j0, tid); break;
    }
    __syncthreads();
    if constexpr (S::R1 > 1) {
        v2_stage1<S, SIGN, true>(sm, tw, tid);
        __syncthreads();
    }
    float2 hold[S::HOLD];
    v2_last_first_a<S, SIGN>(sm, hold, tid);
    __syncthreads();
    v2_last_first_b<S>(sm, hold, tid);
    __syncthreads();
    v2_first_copy_out<S>(sm, P, st, batch, j0, tid);
}

// later passes: plain complex64 input, inter-pass twiddles, fused StoreOp on the last one
template <class S, int SIGN>
__global__ void __launch_bounds__(S::NT, S::MINB) v2_later_kernel(const FftPass P, const LoadC64 ld, const StoreAny st) {
    extern __shared__ float2 rc_v2_smem[];
    float2* sm = rc_v2_smem;
    float2* tw = rc_v2_smem + S::R * V2Smem<S, false>::PITCH;
    const int batch = blockIdx.y + blockIdx.z * gridDim.y;
    const long long j0 = (long long)blockIdx.x * kV2T;
    const int tid = threadIdx.x;
    v2_load_table<S, SIGN>(tw, P, tid);
    __syncthreads();
    v2_stage0<S, SIGN, false>(sm, tw, P, ld, batch, j0, tid);
    __syncthreads();
    if constexpr (S::R1 > 1) {
        v2_stage1<S, SIGN, false>(sm, tw, tid);
        __syncthreads();
    }
    if (st.kind == kStLmr) v2_last_direct<S, SIGN>(sm, P, st.lmr, batch, j0, tid);
    else v2_last_direct<S, SIGN>(sm, P, st.c64, batch, j0, tid);
}
#endif

template <class S, int SIGN>
cudaError_t v2_run_first(const FftPass& P, const LoadAny& ld, const StoreC64& st, int batch, cudaStream_t stream) {
    const long long tiles = (P.stride + kV2T - 1) / kV2T;
#ifdef RC_EMULATE
    (void)stream;
    std::vector<float2> smv(V2Smem<S, true>::ELEMS), hold((size_t)S::NT * S::HOLD);
    float2* sm = smv.data();
    float2* tw = sm + S::R * V2Smem<S, true>::PITCH;
    for (int tid = 0; tid < S::NT; tid++) v2_load_table<S, SIGN>(tw, P, tid);
    for (int b = 0; b < batch; b++)
        for (long long tile = 0; tile < tiles; tile++) {
            const long long j0 = tile * kV2T;
            for (int tid = 0; tid < S::NT; tid++) {
                if (ld.kind == kLdGather) v2_stage0<S, SIGN, true>(sm, tw, P, ld.gather, b, j0, tid);
                else if (ld.kind == kLdDisc) v2_stage0<S, SIGN, true>(sm, tw, P, ld.disc, b, j0, tid);
                else v2_stage0<S, SIGN, true>(sm, tw, P, ld.c64, b, j0, tid);
            }
            if constexpr (S::R1 > 1) for (int tid = 0; tid < S::NT; tid++) v2_stage1<S, SIGN, true>(sm, tw, tid);
            for (int tid = 0; tid < S::NT; tid++) v2_last_first_a<S, SIGN>(sm, hold.data() + (size_t)tid * S::HOLD, tid);
            for (int tid = 0; tid < S::NT; tid++) v2_last_first_b<S>(sm, hold.data() + (size_t)tid * S::HOLD, tid);
            for (int tid = 0; tid < S::NT; tid++) v2_first_copy_out<S>(sm, P, st, b, j0, tid);
        }
    return cudaSuccess;
#else
    int by, bz;
    if (!fft_grid_dims(batch, by, bz)) return cudaErrorInvalidValue;
    dim3 grid((unsigned)tiles, (unsigned)by, (unsigned)bz);
    const size_t smem = (size_t)V2Smem<S, true>::ELEMS * sizeof(float2);
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(v2_first_kernel<S, SIGN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured[dev & 63] = true;
    }
    v2_first_kernel<S, SIGN><<<grid, S::NT, smem, stream>>>(P, ld, st);
    return cudaGetLastError();
#endif
}

template <class S, int SIGN>
cudaError_t v2_run_later(const FftPass& P, const LoadC64& ld, const StoreAny& st, int batch, cudaStream_t stream) {
    const long long tiles = (P.stride + kV2T - 1) / kV2T;
#ifdef RC_EMULATE
    (void)stream;
    std::vector<float2> smv(V2Smem<S, false>::ELEMS);
    float2* sm = smv.data();
    float2* tw = sm + S::R * V2Smem<S, false>::PITCH;
    for (int tid = 0; tid < S::NT; tid++) v2_load_table<S, SIGN>(tw, P, tid);
    for (int b = 0; b < batch; b++)
        for (long long tile = 0; tile < tiles; tile++) {
            const long long j0 = tile * kV2T;
            for (int tid = 0; tid < S::NT; tid++) v2_stage0<S, SIGN, false>(sm, tw, P, ld, b, j0, tid);
            if constexpr (S::R1 > 1) for (int tid = 0; tid < S::NT; tid++) v2_stage1<S, SIGN, false>(sm, tw, tid);
            for (int tid = 0; tid < S::NT; tid++) {
                if (st.kind == kStLmr) v2_last_direct<S, SIGN>(sm, P, st.lmr, b, j0, tid);
                else v2_last_direct<S, SIGN>(sm, P, st.c64, b, j0, tid);
            }
        }
    return cudaSuccess;
#else
    int by, bz;
    if (!fft_grid_dims(batch, by, bz)) return cudaErrorInvalidValue;
    dim3 grid((unsigned)tiles, (unsigned)by, (unsigned)bz);
    const size_t smem = (size_t)V2Smem<S, false>::ELEMS * sizeof(float2);
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(v2_later_kernel<S, SIGN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured[dev & 63] = true;
    }
    v2_later_kernel<S, SIGN><<<grid, S::NT, smem, stream>>>(P, ld, st);
    return cudaGetLastError();
#endif
}

#define RC_V2_CASE_FIRST(id, r0, r1, r2, nt, mb)                                                              \
    case id: return sign < 0 ? v2_run_first<V2Sched<r0, r1, r2, nt, mb>, -1>(P, ld, st, batch, stream)        \
                             : v2_run_first<V2Sched<r0, r1, r2, nt, mb>, +1>(P, ld, st, batch, stream);
#define RC_V2_CASE_LATER(id, r0, r1, r2, nt, mb)                                                              \
    case id: return sign < 0 ? v2_run_later<V2Sched<r0, r1, r2, nt, mb>, -1>(P, ld, st, batch, stream)        \
                             : v2_run_later<V2Sched<r0, r1, r2, nt, mb>, +1>(P, ld, st, batch, stream);

#define RC_V2_DEFINE_GROUP(k, LIST)                                                                           \
    cudaError_t v2_first_g##k(int id, int sign, const FftPass& P, const LoadAny& ld, const StoreC64& st,      \
                              int batch, cudaStream_t stream) {                                               \
        switch (id) { LIST(RC_V2_CASE_FIRST) default: return cudaErrorInvalidValue; }                         \
    }                                                                                                         \
    cudaError_t v2_later_g##k(int id, int sign, const FftPass& P, const LoadC64& ld, const StoreAny& st,      \
                              int batch, cudaStream_t stream) {                                               \
        switch (id) { LIST(RC_V2_CASE_LATER) default: return cudaErrorInvalidValue; }                         \
    }

}  // namespace rc
