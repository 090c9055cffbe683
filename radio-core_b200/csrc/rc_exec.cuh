// rc_exec.cuh -- runs an FftPlan: picks the register-radix passes (rc_fft3.cuh)
// when the plan has them and the fused load/store functors are ones those
// kernels are instantiated for, the generic shared-memory passes otherwise.
#pragma once

#include "rc_ops.cuh"
#include "rc_tma.cuh"

namespace rc {

// ---- the functor kinds the register-radix kernels (rc_fft3.cuh) are compiled for ----
enum { kLdC64 = 0, kLdGather = 1, kLdDisc = 2, kLdTma = 3, kLdAng = 4, kLdGatherTma = 5, kLdAngTma = 6 };
enum { kStC64 = 0, kStLmr = 1, kStWin = 2, kStAng = 3, kStScatter = 4 };

struct LoadAny {
    int kind;
    int box_rows;              // kLdTma: rows per TMA box
    LoadC64 c64;               // kLdC64; also the source described by `tmap` (kLdTma)
    LoadTunerGather gather;
    LoadDiscriminatorPacked disc;
    LoadAnglePacked angle;
    CUtensorMap tmap;          // host copy; handed to the kernel as its own __grid_constant__ argument
    CUtensorMap tmap2;         // kLdGatherTma: the Hann-weight table (tmap describes the spectrum)
};
struct StoreAny {
    int kind;
    StoreC64 c64;
    StoreLmrPacked lmr;
    StoreC64Win win;
    StoreAngle angle;
    StoreScatterC64 scatter;
};

template <class L> struct V3LoadOk { static constexpr bool value = false; };
template <> struct V3LoadOk<LoadC64> { static constexpr bool value = true; };
template <> struct V3LoadOk<LoadTunerGather> { static constexpr bool value = true; };
template <> struct V3LoadOk<LoadDiscriminatorPacked> { static constexpr bool value = true; };
template <> struct V3LoadOk<LoadAnglePacked> { static constexpr bool value = true; };
template <class S> struct V3StoreOk { static constexpr bool value = false; };
template <> struct V3StoreOk<StoreC64> { static constexpr bool value = true; };
template <> struct V3StoreOk<StoreLmrPacked> { static constexpr bool value = true; };
template <> struct V3StoreOk<StoreC64Win> { static constexpr bool value = true; };
template <> struct V3StoreOk<StoreAngle> { static constexpr bool value = true; };
template <> struct V3StoreOk<StoreScatterC64> { static constexpr bool value = true; };

// A complex64 source becomes a TMA tile load when its layout allows a tensor map.
inline LoadAny to_any(const LoadC64& l, const FftPass& P, int batch) {
    LoadAny a{};
    a.kind = kLdC64;
    a.c64 = l;
#ifdef RC_EMULATE
    (void)batch;
    if (P.stride % 2 == 0 && l.batch_stride % 2 == 0) { a.kind = kLdTma; a.box_rows = tma_box_rows(P.R); }
#else
    static const bool no_tma = getenv("RC_NO_TMA") != nullptr;
    TileSource src{l.p, P.stride, l.batch_stride, P.stride, P.R, batch};
    if (!no_tma && tma_source_ok(src)) {
        const int br = tma_box_rows(P.R);
        if (tma_encode_tile_map(&a.tmap, src, br, P.T)) { a.kind = kLdTma; a.box_rows = br; }
    }
#endif
    return a;
}
// The tuner gather becomes two TMA boxes per tile when the geometry is regular: with R even the
// first R/2 rows of a tile are positive-frequency bins and the last R/2 negative-frequency bins,
// each an arithmetic progression of stride S = num/R in the spectrum.  The spectrum is described
// as overlapping rows of S + T bins at stride S, so a box may start at any bin.  Tiles whose
// progression wraps around the end of the spectrum use the per-thread path (decided per CTA).
inline LoadAny to_any(const LoadTunerGather& l, const FftPass& P, int) {
    LoadAny a{};
    a.kind = kLdGather;
    a.gather = l;
#ifndef RC_EMULATE
    static const bool no_tma = getenv("RC_NO_TMA") != nullptr || getenv("RC_NO_TMA_GATHER") != nullptr;
    const long long S = P.stride;
    if (!no_tma && P.R % 2 == 0 && P.R / 2 <= 256 && S % 4 == 0 && l.num == S * P.R && l.half == (long long)(P.R / 2) * S &&
        l.n_x < (1LL << 30) && S + P.T < (1LL << 30)) {
        const unsigned long long rows = (unsigned long long)((l.n_x + S - 1) / S);
        const bool ok1 = tma_encode_2d_f32(&a.tmap, l.X, 2ULL * (unsigned long long)(S + P.T), rows, (unsigned long long)S * 8,
                                           2u * (unsigned)P.T, (unsigned)(P.R / 2));
        // weights: one box when the tile has <= 256 rows, else two halves (their shared-memory
        // destinations must stay 128-byte aligned: R/2 rows of 4*T bytes)
        const unsigned wrows = P.R <= 256 ? (unsigned)P.R : (unsigned)(P.R / 2);
        const bool walign = P.R <= 256 || ((P.R / 2) * P.T * 4) % 128 == 0;
        const bool ok2 = ok1 && walign && tma_encode_2d_f32(&a.tmap2, l.wtab, (unsigned long long)S, (unsigned long long)P.R,
                                                            (unsigned long long)S * 4, (unsigned)P.T, wrows);
        if (ok2) { a.kind = kLdGatherTma; a.box_rows = P.R / 2; }
    }
#endif
    return a;
}
inline LoadAny to_any(const LoadDiscriminatorPacked& l, const FftPass&, int) { LoadAny a{}; a.kind = kLdDisc; a.disc = l; return a; }
inline StoreAny to_any(const StoreC64& s) { StoreAny a{}; a.kind = kStC64; a.c64 = s; return a; }
inline StoreAny to_any(const StoreLmrPacked& s) { StoreAny a{}; a.kind = kStLmr; a.lmr = s; return a; }
// The angle samples of a packed column pair are 16 contiguous bytes: the tile is staged exactly
// like a complex64 tile (two samples = one "complex" element).
inline LoadAny to_any(const LoadAnglePacked& l, const FftPass& P, int batch) {
    LoadAny a{};
    a.kind = kLdAng;
    a.angle = l;
#ifndef RC_EMULATE
    static const bool no_tma = getenv("RC_NO_TMA") != nullptr || getenv("RC_NO_TMA_ANGLE") != nullptr;
    TileSource src{(const float2*)l.ang, P.stride, l.batch_stride / 2, P.stride, P.R, batch};
    if (!no_tma && l.batch_stride % 2 == 0 && tma_source_ok(src)) {
        const int br = tma_box_rows(P.R);
        if (tma_encode_tile_map(&a.tmap, src, br, P.T)) { a.kind = kLdAngTma; a.box_rows = br; }
    }
#endif
    return a;
}
inline StoreAny to_any(const StoreAngle& s) { StoreAny a{}; a.kind = kStAng; a.angle = s; return a; }
inline StoreAny to_any(const StoreScatterC64& s) { StoreAny a{}; a.kind = kStScatter; a.scatter = s; return a; }
inline StoreAny to_any(const StoreC64Win& s) { StoreAny a{}; a.kind = kStWin; a.win = s; return a; }

// Defined in rc_fft3_g<k>.cu (schedule ids with id % kV3Groups == k).
#define RC_V3_DECL(k)                                                                                         \
    cudaError_t v3_first_g##k(int id, int sign, const FftPass& P, const LoadAny& ld, const StoreC64& st,      \
                              int batch, cudaStream_t stream);                                                \
    cudaError_t v3_later_g##k(int id, int sign, const FftPass& P, const LoadAny& ld, const StoreAny& st,      \
                              int batch, cudaStream_t stream);
RC_V3_DECL(0) RC_V3_DECL(1) RC_V3_DECL(2) RC_V3_DECL(3)
#undef RC_V3_DECL

inline cudaError_t v3_first(const FftPass& P, int sign, const LoadAny& ld, const StoreC64& st, int batch,
                            cudaStream_t stream) {
    switch (P.fast_id % kV3Groups) {
        case 0: return v3_first_g0(P.fast_id, sign, P, ld, st, batch, stream);
        case 1: return v3_first_g1(P.fast_id, sign, P, ld, st, batch, stream);
        case 2: return v3_first_g2(P.fast_id, sign, P, ld, st, batch, stream);
        default: return v3_first_g3(P.fast_id, sign, P, ld, st, batch, stream);
    }
}
inline cudaError_t v3_later(const FftPass& P, int sign, const LoadAny& ld, const StoreAny& st, int batch,
                            cudaStream_t stream) {
    switch (P.fast_id % kV3Groups) {
        case 0: return v3_later_g0(P.fast_id, sign, P, ld, st, batch, stream);
        case 1: return v3_later_g1(P.fast_id, sign, P, ld, st, batch, stream);
        case 2: return v3_later_g2(P.fast_id, sign, P, ld, st, batch, stream);
        default: return v3_later_g3(P.fast_id, sign, P, ld, st, batch, stream);
    }
}

inline bool fft_grid_dims(int batch, int& by, int& bz) {
    by = batch; bz = 1;
    while (by > 65535) { bz *= 2; by = (batch + bz - 1) / bz; }
    return (long long)by * bz == batch;          // caller keeps batch <= 65535 or a multiple of the split
}

// one generic shared-memory pass (device launch or CPU replay)
template <int SIGN, class LD, class ST>
cudaError_t fft_generic_pass(const FftPass& P, int batch, const LD& ld, const ST& st, cudaStream_t stream) {
#ifdef RC_EMULATE
    (void)stream;
    long long tiles = (P.stride + P.T - 1) / P.T;
    std::vector<float2> sm(P.smem_elems);
    for (int b = 0; b < batch; b++)
        for (long long tile = 0; tile < tiles; tile++) {
            long long j0 = tile * P.T;
            for (int tid = 0; tid < P.threads; tid++) fft_pass_load<LD, SIGN>(sm.data(), P, ld, b, j0, tid, P.threads);
            int Lprev = 1;
            for (int s = 0; s < P.nstage; s++) {
                for (int tid = 0; tid < P.threads; tid++)
                    fft_stage_dispatch<SIGN>(sm.data(), P, P.radix[s], Lprev, tid, P.threads);
                Lprev *= P.radix[s];
            }
            for (int tid = 0; tid < P.threads; tid++) fft_pass_store<ST>(sm.data(), P, st, b, j0, tid, P.threads);
        }
    return cudaSuccess;
#else
    long long tiles = (P.stride + P.T - 1) / P.T;
    int by, bz;
    if (!fft_grid_dims(batch, by, bz)) return cudaErrorInvalidValue;
    dim3 grid((unsigned)tiles, (unsigned)by, (unsigned)bz);
    size_t smem = (size_t)P.smem_elems * sizeof(float2);
    cudaError_t e = cudaFuncSetAttribute(fft_pass_kernel<LD, ST, SIGN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    fft_pass_kernel<LD, ST, SIGN><<<grid, P.threads, smem, stream>>>(P, ld, st);
    return cudaGetLastError();
#endif
}

// Run all passes.  work0/work1: scratch of batch*n float2 each (work0 needed
// when the plan has >= 2 passes, work1 when >= 3; size by plan.max_passes()).
// LoadOp feeds pass 0, StoreOp drains the last pass.  tag / in_bytes /
// out_bytes feed the optional profiler: compulsory bytes the first pass reads
// through LoadOp and the last pass writes through StoreOp (0 -> 8 bytes per
// element, i.e. a plain complex64 array).
template <int SIGN, class LoadOp, class StoreOp>
cudaError_t fft_exec(const FftPlan& plan, int batch, const LoadOp& ld, const StoreOp& st,
                     float2* work0, float2* work1, cudaStream_t stream, const char* tag = "fft",
                     double in_bytes = 0.0, double out_bytes = 0.0) {
    if (batch <= 0) return cudaSuccess;
    constexpr bool v3ok = V3LoadOk<LoadOp>::value && V3StoreOk<StoreOp>::value;
    const bool fast = v3ok && plan.nfast >= 2 && plan.n < (1LL << 31);
    const int npass = fast ? plan.nfast : plan.npass;
    const FftPass* passes = fast ? plan.fast : plan.pass;
    const double plain = 8.0 * (double)plan.n * (double)batch;
    for (int i = 0; i < npass; i++) {
        const FftPass& P = passes[i];
        const bool first = i == 0, last = i == npass - 1;
        float2* src = ((i - 1) % 2 == 0) ? work0 : work1;
        float2* dst = (i % 2 == 0) ? work0 : work1;
        LoadC64 lmid{src, plan.n};
        StoreC64 smid{dst, plan.n, 1.0f};
        char name[96];
        snprintf(name, sizeof(name), "%s/pass%d_R%d%s", tag, i, P.R, fast ? "r" : "");
        ProfileScope scope(name, ((first && in_bytes > 0) ? in_bytes : plain) + ((last && out_bytes > 0) ? out_bytes : plain), stream);
        cudaError_t e = cudaSuccess;
        if (fast) {
            if constexpr (v3ok) {
                if (first) e = v3_first(P, SIGN, to_any(ld, P, batch), smid, batch, stream);
                else if (last) e = v3_later(P, SIGN, to_any(lmid, P, batch), to_any(st), batch, stream);
                else e = v3_later(P, SIGN, to_any(lmid, P, batch), to_any(smid), batch, stream);
            }
        } else if (first && last) e = fft_generic_pass<SIGN>(P, batch, ld, st, stream);
        else if (first) e = fft_generic_pass<SIGN>(P, batch, ld, smid, stream);
        else if (last) e = fft_generic_pass<SIGN>(P, batch, lmid, st, stream);
        else e = fft_generic_pass<SIGN>(P, batch, lmid, smid, stream);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace rc
