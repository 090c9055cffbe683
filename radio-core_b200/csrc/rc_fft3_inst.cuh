// rc_fft3_inst.cuh -- kernels and launchers of the register-radix passes for one
// group of schedules.  Included by rc_fft3_g<k>.cu with the group's X-macro list
// (rc_fft.cuh), so the groups compile in parallel.
#pragma once

#include "rc_exec.cuh"
#include "rc_fft3.cuh"

namespace rc {

#if defined(__CUDACC__) && !defined(RC_EMULATE)
// Elected thread: arm the mbarrier and issue the TMA boxes of this CTA's tile.
template <class S>
__device__ __forceinline__ void v3_issue_tile(float4* tile, uint64_t* bar, const CUtensorMap* map, int box_rows,
                                              long long j0, int batch) {
    mbar_init(bar, 1);
    mbar_expect_tx(bar, (uint32_t)(S::R * S::T * sizeof(float2)));
    for (int r = 0; r < S::R; r += box_rows)
        tma_load_3d(tile + (size_t)r * S::CP, map, bar, (int)(j0 * 2), r, batch);
}

// first pass of a plan: fused LoadOp (or TMA-staged complex64), column runs re-ordered through shared memory
template <class S, int SIGN>
__global__ void __launch_bounds__(S::NT, S::MINB_FIRST)
v3_first_kernel(const FftPass P, const LoadAny ld, const StoreC64 st, const __grid_constant__ CUtensorMap tmap,
                const __grid_constant__ CUtensorMap tmap2) {
    extern __shared__ __align__(128) unsigned char rc_v3_smem[];
    float4* tile = (float4*)rc_v3_smem;
    float2* tw = (float2*)(rc_v3_smem + (size_t)S::TILE_F4 * 16);
    uint64_t* bar = (uint64_t*)(tw + S::R);
    int* flag = (int*)(bar + 1);
    float* win = (float*)(rc_v3_smem + S::WIN_OFF);           // gather by TMA: [R][T] Hann weights; angle by TMA: [R] samples
    const int batch = blockIdx.y + blockIdx.z * gridDim.y;
    const long long j0 = (long long)blockIdx.x * S::T;
    if (j0 >= P.stride) return;                    // padding CTA of the last cluster
    const int tid = threadIdx.x;
    const bool tma = ld.kind == kLdTma || ld.kind == kLdAngTma;
    if (tma && tid == 0) v3_issue_tile<S>(tile, bar, &tmap, ld.box_rows, j0, batch);
    if (ld.kind == kLdGatherTma && tid == 0) {
        // rows [0, R/2): bins (j0 + t*S - r) mod n;  rows [R/2, R): (j0 + t*S + n - num - r) mod n
        const LoadTunerGather& gth = ld.gather;
        const unsigned nx = (unsigned)gth.n_x, r = (unsigned)ldg(gth.roll + batch), Sd = (unsigned)P.stride;
        unsigned sa = (unsigned)j0 + nx - r;
        if (sa >= nx) sa -= nx;
        unsigned sb = (unsigned)j0 + (unsigned)gth.half + (unsigned)(gth.n_x - gth.num) + nx - r;
        if (sb >= nx) sb -= nx;
        if (sb >= nx) sb -= nx;
        const unsigned span = (unsigned)(S::R / 2 - 1) * Sd + (unsigned)S::T;
        // no wrap-around inside either half, and box starts on 16-byte boundaries (even bins)
        const int ok = (sa + span <= nx) && (sb + span <= nx) && (((sa | sb) & 1u) == 0u);
        *flag = ok;
        if (ok) {
            mbar_init(bar, 1);
            mbar_expect_tx(bar, (uint32_t)(S::R * S::T * (sizeof(float2) + sizeof(float))));
            tma_load_2d(tile, &tmap, bar, (int)(2u * (sa % Sd)), (int)(sa / Sd));
            tma_load_2d(tile + (size_t)(S::R / 2) * S::CP, &tmap, bar, (int)(2u * (sb % Sd)), (int)(sb / Sd));
            tma_load_2d(win, &tmap2, bar, (int)j0, 0);
            if (S::R > 256) tma_load_2d(win + (size_t)(S::R / 2) * S::T, &tmap2, bar, (int)j0, S::R / 2);
        }
    }
    if (ld.kind == kLdAngTma) {
        // the sample before each row's first column lies in the neighbouring tile: fetch the R of
        // them now, behind the TMA load, instead of one dependent load per row inside stage 0
        const float* a = ld.angle.ang + batch * ld.angle.batch_stride;
        for (int r = tid; r < S::R; r += S::NT) {
            const long long i = 2 * (j0 + (long long)r * P.stride);
            win[r] = ldg(a + (i > 0 ? i - 1 : 0));
        }
    }
    v3_load_table<S, SIGN>(tw, P, tid);
    const V3Tw tws = v3_twiddle_setup<S, false>(P, j0, tid);
    __syncthreads();
    int kind = ld.kind;
    if (kind == kLdGatherTma) {
        if (*flag) {
            mbar_wait(bar, 0);
            if (j0 == 0) {
                // column 0 of row R/2 is bin num/2: the box delivered its merged partner X[n - num/2 - r]
                if (tid == 0) {
                    const LoadTunerGather& gth = ld.gather;
                    const long long r = ldg(gth.roll + batch);
                    long long sp = gth.half - r;
                    if (sp < 0) sp += gth.n_x;
                    const float2 xp = ldg(gth.X + sp);
                    float2* e = (float2*)(tile + (size_t)(S::R / 2) * S::CP);
                    const float ratio = gth.w_neg_half / win[(size_t)(S::R / 2) * S::T];
                    *e = make_float2(xp.x + ratio * e->x, xp.y + ratio * e->y);
                }
                __syncthreads();
            }
            v3_stage0<S, SIGN, false>(tile, tw, P, V3FromTileWin<S::CP>{tile, win}, batch, j0, tid, tws);
            kind = -1;
        } else {
            kind = kLdGather;                      // this tile wraps around the spectrum: per-thread loads
        }
    }
    switch (kind) {
        case -1: break;
        case kLdTma:
            mbar_wait(bar, 0);
            v3_stage0<S, SIGN, false>(tile, tw, P, V3FromTile<S::CP>{tile}, batch, j0, tid, tws);
            break;
        case kLdAngTma:
            mbar_wait(bar, 0);
            v3_stage0<S, SIGN, false>(tile, tw, P, V3FromTileAng<S::CP>{tile, win},
                                      batch, j0, tid, tws);
            break;
        case kLdGather: v3_stage0<S, SIGN, false>(tile, tw, P, V3FromOp<LoadTunerGather>{&ld.gather, P.stride}, batch, j0, tid, tws); break;
        case kLdDisc: v3_stage0<S, SIGN, false>(tile, tw, P, V3FromOp<LoadDiscriminatorPacked>{&ld.disc, P.stride}, batch, j0, tid, tws); break;
        case kLdAng: v3_stage0<S, SIGN, false>(tile, tw, P, V3FromOp<LoadAnglePacked>{&ld.angle, P.stride}, batch, j0, tid, tws); break;
        default: v3_stage0<S, SIGN, false>(tile, tw, P, V3FromOp<LoadC64>{&ld.c64, P.stride}, batch, j0, tid, tws); break;
    }
    __syncthreads();
    if constexpr (S::R1 > 1) {
        v3_stage1<S, SIGN>(tile, tw, tid);
        __syncthreads();
    }
    float4 hold[S::HOLD];
    v3_last_first_a<S, SIGN>(tile, hold, tid);
    __syncthreads();
    v3_last_first_b<S>((float2*)tile, hold, tid);
    __syncthreads();
    v3_first_copy_out<S>((const float2*)tile, P, st, batch, j0, tid);
}

// later passes: complex64 input (TMA-staged when the layout allows), inter-pass twiddles,
// fused StoreOp on the last one
template <class S, int SIGN>
__global__ void __launch_bounds__(S::NT, S::MINB)
v3_later_kernel(const FftPass P, const LoadAny ld, const StoreAny st, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(128) unsigned char rc_v3_smem[];
    float4* tile = (float4*)rc_v3_smem;
    float2* tw = (float2*)(rc_v3_smem + (size_t)S::TILE_F4 * 16);
    uint64_t* bar = (uint64_t*)(tw + S::R);
    const int batch = blockIdx.y + blockIdx.z * gridDim.y;
    const long long j0 = (long long)blockIdx.x * S::T;
    if (j0 >= P.stride) return;
    const int tid = threadIdx.x;
    const bool tma = ld.kind == kLdTma;
    if (tma && tid == 0) v3_issue_tile<S>(tile, bar, &tmap, ld.box_rows, j0, batch);
    v3_load_table<S, SIGN>(tw, P, tid);
    const V3Tw tws = v3_twiddle_setup<S, true>(P, j0, tid);      // table look-ups overlap the TMA load
    __syncthreads();
    if (tma) {
        mbar_wait(bar, 0);
        v3_stage0<S, SIGN, true>(tile, tw, P, V3FromTile<S::CP>{tile}, batch, j0, tid, tws);
    } else {
        v3_stage0<S, SIGN, true>(tile, tw, P, V3FromOp<LoadC64>{&ld.c64, P.stride}, batch, j0, tid, tws);
    }
    __syncthreads();
    if constexpr (S::R1 > 1) {
        v3_stage1<S, SIGN>(tile, tw, tid);
        __syncthreads();
    }
    if (st.kind == kStLmr) v3_last_direct<S, SIGN>(tile, P, st.lmr, batch, j0, tid);
    else if (st.kind == kStWin) v3_last_direct<S, SIGN>(tile, P, st.win, batch, j0, tid);
    else if (st.kind == kStAng) v3_last_direct<S, SIGN>(tile, P, st.angle, batch, j0, tid);
    else if (st.kind == kStScatter) v3_last_direct<S, SIGN>(tile, P, st.scatter, batch, j0, tid);
    else v3_last_direct<S, SIGN>(tile, P, st.c64, batch, j0, tid);
}

// Launch as clusters of two adjacent tiles (grid.x padded to even; the padding CTA exits).
template <class K, class... Args>
cudaError_t v3_launch(K kernel, long long tiles, int batch, int threads, size_t smem, cudaStream_t stream, Args... args) {
    int by, bz;
    if (!fft_grid_dims(batch, by, bz)) return cudaErrorInvalidValue;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((tiles + 1) / 2 * 2), (unsigned)by, (unsigned)bz);
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}
#endif

#ifdef RC_EMULATE
// host stand-in of the TMA box loads: rows of 16 columns, columns past the end read as zero
template <class S>
void v3_emulate_tma(float4* tile, const LoadC64& src, const FftPass& P, int batch, long long j0) {
    float2* t = (float2*)tile;
    for (int r = 0; r < S::R; r++)
        for (int c = 0; c < S::T; c++) {
            const long long j = j0 + c;
            t[r * S::T + c] = j < P.stride ? src(batch, j + (long long)r * P.stride) : make_float2(0.f, 0.f);
        }
}
#endif

template <class S, int SIGN>
cudaError_t v3_run_first(const FftPass& P, const LoadAny& ld, const StoreC64& st, int batch, cudaStream_t stream) {
    const long long tiles = (P.stride + S::T - 1) / S::T;
#ifdef RC_EMULATE
    (void)stream;
    std::vector<float4> smv((size_t)S::SMEM_BYTES / 16 + 1), hold((size_t)S::NT * S::HOLD);
    float4* tile = smv.data();
    float2* tw = (float2*)(tile + S::TILE_F4);
    for (int tid = 0; tid < S::NT; tid++) v3_load_table<S, SIGN>(tw, P, tid);
    for (int b = 0; b < batch; b++)
        for (long long t = 0; t < tiles; t++) {
            const long long j0 = t * S::T;
            if (ld.kind == kLdTma) v3_emulate_tma<S>(tile, ld.c64, P, b, j0);
            for (int tid = 0; tid < S::NT; tid++) {
                const V3Tw tws = v3_twiddle_setup<S, false>(P, j0, tid);
                if (ld.kind == kLdTma) v3_stage0<S, SIGN, false>(tile, tw, P, V3FromTile<S::CP>{tile}, b, j0, tid, tws);
                else if (ld.kind == kLdGather) v3_stage0<S, SIGN, false>(tile, tw, P, V3FromOp<LoadTunerGather>{&ld.gather, P.stride}, b, j0, tid, tws);
                else if (ld.kind == kLdDisc) v3_stage0<S, SIGN, false>(tile, tw, P, V3FromOp<LoadDiscriminatorPacked>{&ld.disc, P.stride}, b, j0, tid, tws);
                else if (ld.kind == kLdAng) v3_stage0<S, SIGN, false>(tile, tw, P, V3FromOp<LoadAnglePacked>{&ld.angle, P.stride}, b, j0, tid, tws);
                else v3_stage0<S, SIGN, false>(tile, tw, P, V3FromOp<LoadC64>{&ld.c64, P.stride}, b, j0, tid, tws);
            }
            if constexpr (S::R1 > 1) for (int tid = 0; tid < S::NT; tid++) v3_stage1<S, SIGN>(tile, tw, tid);
            for (int tid = 0; tid < S::NT; tid++) v3_last_first_a<S, SIGN>(tile, hold.data() + (size_t)tid * S::HOLD, tid);
            for (int tid = 0; tid < S::NT; tid++) v3_last_first_b<S>((float2*)tile, hold.data() + (size_t)tid * S::HOLD, tid);
            for (int tid = 0; tid < S::NT; tid++) v3_first_copy_out<S>((const float2*)tile, P, st, b, j0, tid);
        }
    return cudaSuccess;
#else
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(v3_first_kernel<S, SIGN>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::SMEM_BYTES_WIN);
        if (e != cudaSuccess) return e;
        configured[dev & 63] = true;
    }
    const size_t smem = ld.kind == kLdGatherTma ? (size_t)S::SMEM_BYTES_WIN
                      : ld.kind == kLdAngTma  ? (size_t)S::SMEM_BYTES_ANG : (size_t)S::SMEM_BYTES;
    return v3_launch(v3_first_kernel<S, SIGN>, tiles, batch, S::NT, smem, stream, P, ld, st, ld.tmap, ld.tmap2);
#endif
}

template <class S, int SIGN>
cudaError_t v3_run_later(const FftPass& P, const LoadAny& ld, const StoreAny& st, int batch, cudaStream_t stream) {
    const long long tiles = (P.stride + S::T - 1) / S::T;
#ifdef RC_EMULATE
    (void)stream;
    std::vector<float4> smv((size_t)S::SMEM_BYTES / 16 + 1);
    float4* tile = smv.data();
    float2* tw = (float2*)(tile + S::TILE_F4);
    for (int tid = 0; tid < S::NT; tid++) v3_load_table<S, SIGN>(tw, P, tid);
    for (int b = 0; b < batch; b++)
        for (long long t = 0; t < tiles; t++) {
            const long long j0 = t * S::T;
            if (ld.kind == kLdTma) v3_emulate_tma<S>(tile, ld.c64, P, b, j0);
            for (int tid = 0; tid < S::NT; tid++) {
                const V3Tw tws = v3_twiddle_setup<S, true>(P, j0, tid);
                if (ld.kind == kLdTma) v3_stage0<S, SIGN, true>(tile, tw, P, V3FromTile<S::CP>{tile}, b, j0, tid, tws);
                else v3_stage0<S, SIGN, true>(tile, tw, P, V3FromOp<LoadC64>{&ld.c64, P.stride}, b, j0, tid, tws);
            }
            if constexpr (S::R1 > 1) for (int tid = 0; tid < S::NT; tid++) v3_stage1<S, SIGN>(tile, tw, tid);
            for (int tid = 0; tid < S::NT; tid++) {
                if (st.kind == kStLmr) v3_last_direct<S, SIGN>(tile, P, st.lmr, b, j0, tid);
                else if (st.kind == kStWin) v3_last_direct<S, SIGN>(tile, P, st.win, b, j0, tid);
                else if (st.kind == kStAng) v3_last_direct<S, SIGN>(tile, P, st.angle, b, j0, tid);
                else if (st.kind == kStScatter) v3_last_direct<S, SIGN>(tile, P, st.scatter, b, j0, tid);
                else v3_last_direct<S, SIGN>(tile, P, st.c64, b, j0, tid);
            }
        }
    return cudaSuccess;
#else
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(v3_later_kernel<S, SIGN>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured[dev & 63] = true;
    }
    return v3_launch(v3_later_kernel<S, SIGN>, tiles, batch, S::NT, (size_t)S::SMEM_BYTES, stream, P, ld, st, ld.tmap);
#endif
}

// role (rc_fft.cuh): schedules restricted to one kind of pass instantiate only that kernel
template <class S, int ROLE>
cudaError_t v3_role_first(int sign, const FftPass& P, const LoadAny& ld, const StoreC64& st, int batch, cudaStream_t stream) {
    if constexpr (ROLE == 2) return cudaErrorInvalidValue;
    else return sign < 0 ? v3_run_first<S, -1>(P, ld, st, batch, stream) : v3_run_first<S, +1>(P, ld, st, batch, stream);
}
template <class S, int ROLE>
cudaError_t v3_role_later(int sign, const FftPass& P, const LoadAny& ld, const StoreAny& st, int batch, cudaStream_t stream) {
    if constexpr (ROLE == 1) return cudaErrorInvalidValue;
    else return sign < 0 ? v3_run_later<S, -1>(P, ld, st, batch, stream) : v3_run_later<S, +1>(P, ld, st, batch, stream);
}

#define RC_V3_CASE_FIRST(id, r0, r1, r2, nt, mb, cp, role) \
    case id: return v3_role_first<V3Sched<r0, r1, r2, nt, mb, cp>, role>(sign, P, ld, st, batch, stream);
#define RC_V3_CASE_LATER(id, r0, r1, r2, nt, mb, cp, role) \
    case id: return v3_role_later<V3Sched<r0, r1, r2, nt, mb, cp>, role>(sign, P, ld, st, batch, stream);

#define RC_V3_DEFINE_GROUP(k, LIST)                                                                           \
    cudaError_t v3_first_g##k(int id, int sign, const FftPass& P, const LoadAny& ld, const StoreC64& st,      \
                              int batch, cudaStream_t stream) {                                               \
        switch (id) { LIST(RC_V3_CASE_FIRST) default: return cudaErrorInvalidValue; }                         \
    }                                                                                                         \
    cudaError_t v3_later_g##k(int id, int sign, const FftPass& P, const LoadAny& ld, const StoreAny& st,      \
                              int batch, cudaStream_t stream) {                                               \
        switch (id) { LIST(RC_V3_CASE_LATER) default: return cudaErrorInvalidValue; }                         \
    }

}  // namespace rc
