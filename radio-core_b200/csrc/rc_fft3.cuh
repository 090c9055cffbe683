// rc_fft3.cuh -- register-radix FFT pass kernels, two columns per thread
// (the fast path of rc_fft.cuh; replaces the one-column kernels of round 1).
//
// Pass semantics are those of rc_fft.cuh: one Stockham step of radix R on a tile
// of T = 16 adjacent columns, R = R0*R1*R2 evaluated as 2 or 3 in-register radix
// stages with shared-memory exchanges in between.  What is new:
//
//   * a thread owns a PAIR of adjacent columns: the tile is an array of float4
//     [R rows][8 column pairs], every shared-memory access is a 16-byte LDS/STS
//     (a quarter-warp covers one 128-byte row: conflict-free), global stores of
//     later passes are 16-byte STG, and row/twiddle index arithmetic is shared
//     by the two butterflies;
//   * the complex arithmetic is packed f32x2 (FADD2/FMUL2/FFMA2, rc_fft.cuh);
//   * plain complex64 tiles are staged by TMA: one elected thread issues
//     cp.async.bulk.tensor loads of [<=256 rows x 16 columns] boxes that complete
//     on an mbarrier (rc_tma.cuh); CTAs are launched as clusters of two adjacent
//     tiles so both 128-byte halves of a 256-byte DRAM segment are requested
//     together (measured: 4.5 -> 5.4 TB/s on strided tiles, tools/membench.cu);
//   * inter-pass twiddles W_M^{t*kj}: three fp64 table look-ups per thread and
//     tile, then fp64 recurrences along t (exact to ~1e-16), rounded to fp32 once.
//
// Digit bookkeeping for one column (t = input row, K = output row):
//   t = t0*(R1*R2) + t1*R2 + t2        K = k0 + R0*k1 + R0*R1*k2
//   shared row rho = d0*(R1*R2) + d1*R2 + d2, digit d_s holds t_s before
//   stage s and k_s after it.
//   W_R^{tK} = W_R0^{t0k0} * W_R^{t1*k0*R2} * W_R1^{t1k1} * W_R^{t2*(k0+R0k1)} * W_R2^{t2k2}
#pragma once

#include "rc_fft.cuh"
#include "rc_tma.cuh"

namespace rc {

// CP = column pairs per tile: 8 (16 columns, 128-byte rows) or, for short passes whose tiles
// stay small, 16 (32 columns, 256-byte rows: a whole DRAM segment per row) or 32 (64 columns:
// one warp per row; R = 50, whose last stage has only 5 butterflies per column pair and would
// leave half of 10 row groups idle at 32 columns).
template <int R0_, int R1_, int R2_, int NT_, int MINB_, int CP_ = 8>
struct V3Sched {
    static constexpr int R0 = R0_, R1 = R1_, R2 = R2_, NT = NT_, MINB = MINB_;
    static constexpr int CP = CP_, T = 2 * CP_, LOGCP = CP_ == 8 ? 3 : (CP_ == 16 ? 4 : 5);
    static constexpr int R = R0_ * R1_ * R2_;
    static constexpr int U = R1_ * R2_;
    static constexpr int NG = NT_ / CP_;                     // row groups working in parallel
    static constexpr int NB0 = U, NB1 = R0_ * R2_, NB2 = R0_ * R1_;   // butterflies (per column pair) of each stage
    static constexpr int IT0 = (NB0 + NG - 1) / NG, IT1 = (NB1 + NG - 1) / NG, IT2 = (NB2 + NG - 1) / NG;
    static constexpr int HOLD = IT2 * R2_;                   // float4 registers held across the re-order (first pass)
    static constexpr int MINB_F_ = 65536 / (NT_ * (HOLD * 4 + 48));
    static constexpr int MINB_FIRST = MINB_F_ < 1 ? 1 : (MINB_F_ < MINB_ ? MINB_F_ : MINB_);
    static constexpr int PITCH = (R % 2 == 0) ? R + 1 : R;   // float2 pitch of the transposed [column][K] layout (odd)
    static constexpr int TILE_F4 = R * CP_ + 2 * CP_;        // float4 slots: tile, and room for the transposed layout
    static constexpr int SMEM_BYTES = TILE_F4 * 16 + R * 8 + 16;   // + W_R table + mbarrier
    static constexpr int WIN_OFF = (SMEM_BYTES + 127) / 128 * 128;  // tuner gather by TMA: + [R][T] Hann weights
    static constexpr int SMEM_BYTES_WIN = WIN_OFF + R * T * 4;
    static constexpr int SMEM_BYTES_ANG = WIN_OFF + R * 4;          // angle tile by TMA: + one preceding sample per row
    static_assert(CP_ == 8 || CP_ == 16 || CP_ == 32, "8, 16 or 32 column pairs");
    static_assert(NT_ % CP_ == 0, "threads must be a multiple of the column pairs");
    static_assert(T * PITCH <= TILE_F4 * 2, "transposed layout must fit the tile buffer");
};

RC_HD float wrap_half_turns_dev(float x) {      // (-2, 2) half-turns -> (-1, 1]   (= wrap_half_turns of rc_ops.cuh)
#ifdef __CUDA_ARCH__
    return x - 2.0f * rintf(0.5f * x);
#else
    return x - 2.0f * nearbyintf(0.5f * x);
#endif
}

RC_HD float2 f4lo(float4 v) { return make_float2(v.x, v.y); }
RC_HD float2 f4hi(float4 v) { return make_float2(v.z, v.w); }
RC_HD float4 f4make(float2 a, float2 b) { return make_float4(a.x, a.y, b.x, b.y); }

// W_R table: forward sign in global memory, conjugated here for the inverse.
template <class S, int SIGN>
RC_HD void v3_load_table(float2* tw, const FftPass& P, int tid) {
    for (int i = tid; i < S::R; i += S::NT) {
        float2 w = ldg(P.twR + i);
        if (SIGN > 0) w.y = -w.y;
        tw[i] = w;
    }
}

// Sources of stage 0: the tile already staged in shared memory (TMA), or a LoadOp functor.
template <int CP> struct V3FromTile {
    static constexpr bool kTile = true;         // staged tile: columns past the end hold zeros, every lane may read
    const float4* tile;
    struct Ctx {};
    RC_HD Ctx prepare(int) const { return Ctx{}; }
    RC_HD float4 get(const Ctx&, int row, int cp, long long, bool) const { return tile[row * CP + cp]; }
};
// tile of raw spectrum bins staged by TMA next to a tile of their Hann weights (tuner gather)
template <int CP> struct V3FromTileWin {
    static constexpr bool kTile = true;
    const float4* tile;
    const float* win;          // [R][2*CP] weights
    struct Ctx {};
    RC_HD Ctx prepare(int) const { return Ctx{}; }
    RC_HD float4 get(const Ctx&, int row, int cp, long long, bool) const {
        const float4 x = tile[row * CP + cp];
        const float2 w = *(const float2*)(win + row * 2 * CP + 2 * cp);
        const float2 a = cscale(make_float2(x.x, x.y), w.x), d = cscale(make_float2(x.z, x.w), w.y);
        return make_float4(a.x, a.y, d.x, d.y);
    }
};
#if defined(__CUDACC__) && !defined(RC_EMULATE)
// tile of angle(y)/pi samples staged by TMA (two samples per column): packed FM discriminator
// d[n] = wrap(a[n] - a[n-1]) (LoadAnglePacked).  The sample before a column pair comes from the
// neighbouring lane (same row); the one before a row's first pair was fetched into `before`
// while the tile was in flight (a[-1] := a[0], so that d[0] = 0).
template <int CP> struct V3FromTileAng {
    static constexpr bool kTile = true;
    const float4* tile;
    const float* before;       // shared: the sample preceding each row's first column (v3_first_kernel)
    struct Ctx {};
    __device__ Ctx prepare(int) const { return Ctx{}; }
    __device__ float4 get(const Ctx&, int row, int cp, long long, bool) const {
        const float4 x = tile[row * CP + cp];
        const unsigned lane = threadIdx.x & 31u;
        const unsigned group = 0xFFu << (lane & 24u);                 // the 8 (or 2 x 8) lanes of this row group
        float prev = __shfl_up_sync(CP == 8 ? group : (0xFFFFu << (lane & 16u)), x.w, 1, CP);
        if (cp == 0) prev = before[row];
        const float d0 = wrap_half_turns_dev(x.x - prev);
        return make_float4(d0, wrap_half_turns_dev(x.y - x.x), wrap_half_turns_dev(x.z - x.y), wrap_half_turns_dev(x.w - x.z));
    }
};
#endif
template <class LoadOp> struct V3FromOp {
    static constexpr bool kTile = false;
    const LoadOp* ld;
    long long stride;
    typedef typename LoadOp::Ctx Ctx;
    RC_HD Ctx prepare(int batch) const { return ld->prepare(batch); }
    RC_HD float4 get(const Ctx& c, int row, int, long long j, bool has_b) const {
        return ld->load2(c, j + (long long)row * stride, has_b);
    }
};

template <int SIGN> RC_HD float2 v3_tw32(double2 w) {
    return make_float2((float)w.x, (float)(SIGN < 0 ? w.y : -w.y));
}

// Per-thread inter-pass twiddle state of a later pass: W_M^{u*kj} for the thread's first
// row group, its steps along the row groups (NG*kj) and along t0 (U*kj), for both columns.
// Three table look-ups per column; issued before the tile arrives so their latency hides
// behind the TMA load.
struct V3Tw { double2 wa, wb, sta, stb, wsa, wsb; };

// W_M^q for q < M < 2^32 (always the case here: q <= U*(Ns-1) < M/R0): 32-bit index arithmetic
RC_HD double2 v3_tw64(const FftPass& P, unsigned q) {
    const double2 a = ldg(P.tw_lo + (q & P.tw_mask));
    const double2 b = ldg(P.tw_hi + (q >> P.tw_shift));
    return cmul64(a, b);
}

template <class S, bool LATER>
RC_HD V3Tw v3_twiddle_setup(const FftPass& P, long long j0, int tid) {
    V3Tw t;
    t.wa = t.wb = t.sta = t.stb = t.wsa = t.wsb = make_double2(1.0, 0.0);
    if (LATER) {
        const int cp = tid & (S::CP - 1), g = tid >> S::LOGCP;
        const unsigned ns = (unsigned)P.Ns;
        const unsigned ka = (unsigned)(j0 + 2 * cp) % ns;
        // column A from the two-level table; column B = A's twiddle times W^{g}, W^{NG}, W^{U}
        // (kb = ka + 1), which are the same for every column: three short look-ups
        t.wa = v3_tw64(P, (unsigned)g * ka);
        t.wsa = v3_tw64(P, (unsigned)S::U * ka);
        if (S::IT0 > 1) t.sta = v3_tw64(P, (unsigned)S::NG * ka);
        if (ka + 1u != ns) {
            t.wb = cmul64(t.wa, v3_tw64(P, (unsigned)g));
            t.wsb = cmul64(t.wsa, v3_tw64(P, (unsigned)S::U));
            if (S::IT0 > 1) t.stb = cmul64(t.sta, v3_tw64(P, (unsigned)S::NG));
        }                                   // else kb = 0: all ones
    }
    return t;
}

// stage 0: inputs (+ inter-pass twiddles) -> radix R0 -> intra-pass twiddle -> tile (in place)
template <class S, int SIGN, bool LATER, class Src>
RC_HD void v3_stage0(float4* tile, const float2* tw, const FftPass& P, const Src& src, int batch, long long j0, int tid,
                     const V3Tw& tws) {
    const int cp = tid & (S::CP - 1), g = tid >> S::LOGCP;
    const long long j = j0 + 2 * cp;
    const bool act_a = j < P.stride, act_b = j + 1 < P.stride;
    double2 wa = tws.wa, wb = tws.wb;
    const double2 sta = tws.sta, stb = tws.stb, wsa = tws.wsa, wsb = tws.wsb;
    const typename Src::Ctx sctx = src.prepare(batch);
#pragma unroll
    for (int it = 0; it < S::IT0; it++) {
        const int u = g + it * S::NG;
        if (S::NB0 % S::NG != 0 && u >= S::NB0) break;
        float2 a[S::R0], b[S::R0];
        if (act_a || Src::kTile) {
#pragma unroll
            for (int t0 = 0; t0 < S::R0; t0++) {
                const float4 x = src.get(sctx, t0 * S::U + u, cp, j, act_b);
                a[t0] = f4lo(x);
                b[t0] = f4hi(x);
            }
        } else {
#pragma unroll
            for (int t0 = 0; t0 < S::R0; t0++) a[t0] = b[t0] = make_float2(0.f, 0.f);
        }
        if (LATER) {
            double2 pa = wa, pb = wb;
#pragma unroll
            for (int t0 = 0; t0 < S::R0; t0++) {
                a[t0] = cmul(a[t0], v3_tw32<SIGN>(pa));
                b[t0] = cmul(b[t0], v3_tw32<SIGN>(pb));
                if (t0 + 1 < S::R0) { pa = cmul64(pa, wsa); pb = cmul64(pb, wsb); }
            }
            if (it + 1 < S::IT0) { wa = cmul64(wa, sta); wb = cmul64(wb, stb); }
        }
        Dft<S::R0, SIGN>::run(a);
        Dft<S::R0, SIGN>::run(b);
        // intra-pass twiddle on the outputs: two-stage W_R^{t2*k0} (u = t2), three-stage W_R^{t1*k0*R2}
        const int tq = (S::R1 == 1) ? u : (u / S::R2) * S::R2;
        if (tq > 0) {
#pragma unroll
            for (int k0 = 1; k0 < S::R0; k0++) {
                const float2 w = tw[tq * k0];
                a[k0] = cmul(a[k0], w);
                b[k0] = cmul(b[k0], w);
            }
        }
#pragma unroll
        for (int k0 = 0; k0 < S::R0; k0++) tile[(k0 * S::U + u) * S::CP + cp] = f4make(a[k0], b[k0]);
    }
}

// middle stage (three-stage schedules only): radix R1 over digit d1, in place
template <class S, int SIGN>
RC_HD void v3_stage1(float4* tile, const float2* tw, int tid) {
    const int cp = tid & (S::CP - 1), g = tid >> S::LOGCP;
#pragma unroll
    for (int it = 0; it < S::IT1; it++) {
        const int q = g + it * S::NG;
        if (S::NB1 % S::NG != 0 && q >= S::NB1) break;
        const int t2 = q % S::R2, k0 = q / S::R2;
        const int base = k0 * S::U + t2;
        float2 a[S::R1], b[S::R1];
#pragma unroll
        for (int t1 = 0; t1 < S::R1; t1++) {
            const float4 x = tile[(base + t1 * S::R2) * S::CP + cp];
            a[t1] = f4lo(x);
            b[t1] = f4hi(x);
        }
        Dft<S::R1, SIGN>::run(a);
        Dft<S::R1, SIGN>::run(b);
        if (t2 > 0) {
#pragma unroll
            for (int k1 = 0; k1 < S::R1; k1++) {
                const int e = t2 * (k0 + S::R0 * k1);
                if (e > 0) {
                    const float2 w = tw[e];
                    a[k1] = cmul(a[k1], w);
                    b[k1] = cmul(b[k1], w);
                }
            }
        }
#pragma unroll
        for (int k1 = 0; k1 < S::R1; k1++) tile[(base + k1 * S::R2) * S::CP + cp] = f4make(a[k1], b[k1]);
    }
}

// Where the outputs of one column pair of a later pass go: element K of column A at
// oa + K*ns, of column B at ob + K*ns (the Stockham position in the batch entry).
struct V3Out {
    long long oa, ob, ns;
    bool act_a, act_b, pair;
};

template <class S>
RC_HD V3Out v3_out_default(const FftPass& P, long long j0, int tid) {
    const int cp = tid & (S::CP - 1);
    const long long j = j0 + 2 * cp;
    V3Out o;
    o.act_a = j < P.stride;
    o.act_b = j + 1 < P.stride;
    o.ns = P.Ns;
    const unsigned qn32 = (unsigned)j / (unsigned)P.Ns;          // n < 2^31 on this path: 32-bit division
    const long long qn = qn32, rem = j - qn * o.ns;
    o.oa = qn * o.ns * S::R + rem;
    o.ob = (rem + 1 < o.ns) ? o.oa + 1 : (qn + 1) * o.ns * S::R;
    o.pair = P.pair_ok && o.act_b;
    return o;
}

// last stage, passes after the first: radix R2 over digit d2, results straight to global
template <class S, int SIGN, class StoreOp>
RC_HD void v3_last_direct(const float4* tile, const StoreOp& st, int batch, const V3Out& o, int tid) {
    const int cp = tid & (S::CP - 1), g = tid >> S::LOGCP;
    if (!o.act_a) return;
    const long long ns = o.ns, oa = o.oa, ob = o.ob;
    const bool act_b = o.act_b, pair = o.pair;
#pragma unroll
    for (int it = 0; it < S::IT2; it++) {
        const int q = g + it * S::NG;
        if (S::NB2 % S::NG != 0 && q >= S::NB2) break;
        const int k1 = q % S::R1, k0 = q / S::R1;
        const int base = k0 * S::U + k1 * S::R2;
        float2 a[S::R2], b[S::R2];
#pragma unroll
        for (int t2 = 0; t2 < S::R2; t2++) {
            const float4 x = tile[(base + t2) * S::CP + cp];
            a[t2] = f4lo(x);
            b[t2] = f4hi(x);
        }
        Dft<S::R2, SIGN>::run(a);
        Dft<S::R2, SIGN>::run(b);
        const long long row0 = (long long)(k0 + S::R0 * k1) * ns;
        const long long rstep = (long long)(S::R0 * S::R1) * ns;
        if (pair) {
#pragma unroll
            for (int k2 = 0; k2 < S::R2; k2++) st.pair(batch, oa + row0 + k2 * rstep, a[k2], b[k2]);
        } else {
#pragma unroll
            for (int k2 = 0; k2 < S::R2; k2++) {
                st(batch, oa + row0 + k2 * rstep, a[k2]);
                if (act_b) st(batch, ob + row0 + k2 * rstep, b[k2]);
            }
        }
    }
}
template <class S, int SIGN, class StoreOp>
RC_HD void v3_last_direct(const float4* tile, const FftPass& P, const StoreOp& st, int batch, long long j0, int tid) {
    v3_last_direct<S, SIGN>(tile, st, batch, v3_out_default<S>(P, j0, tid), tid);
}

// last stage of the first pass, part A: read + radix into held registers
template <class S, int SIGN>
RC_HD void v3_last_first_a(const float4* tile, float4* hold, int tid) {
    const int cp = tid & (S::CP - 1), g = tid >> S::LOGCP;
#pragma unroll
    for (int it = 0; it < S::IT2; it++) {
        const int q = g + it * S::NG;
        if (S::NB2 % S::NG != 0 && q >= S::NB2) break;
        const int k0 = q % S::R0, k1 = q / S::R0;      // k0 fastest: the row groups of a warp write consecutive K
        const int base = k0 * S::U + k1 * S::R2;
        float2 a[S::R2], b[S::R2];
#pragma unroll
        for (int t2 = 0; t2 < S::R2; t2++) {
            const float4 x = tile[(base + t2) * S::CP + cp];
            a[t2] = f4lo(x);
            b[t2] = f4hi(x);
        }
        Dft<S::R2, SIGN>::run(a);
        Dft<S::R2, SIGN>::run(b);
#pragma unroll
        for (int k2 = 0; k2 < S::R2; k2++) hold[it * S::R2 + k2] = f4make(a[k2], b[k2]);
    }
}
// part B (after a barrier): write the held results transposed, [column][K] with an odd pitch
template <class S>
RC_HD void v3_last_first_b(float2* tr, const float4* hold, int tid) {
    const int cp = tid & (S::CP - 1), g = tid >> S::LOGCP;
#pragma unroll
    for (int it = 0; it < S::IT2; it++) {
        const int q = g + it * S::NG;
        if (S::NB2 % S::NG != 0 && q >= S::NB2) break;
        const int k0 = q % S::R0, k1 = q / S::R0;      // k0 fastest: the row groups of a warp write consecutive K
        const int K0 = k0 + S::R0 * k1;
#pragma unroll
        for (int k2 = 0; k2 < S::R2; k2++) {
            const float4 v = hold[it * S::R2 + k2];
            const int K = K0 + S::R0 * S::R1 * k2;
            tr[(2 * cp) * S::PITCH + K] = f4lo(v);
            tr[(2 * cp + 1) * S::PITCH + K] = f4hi(v);
        }
    }
}
// part C (after a barrier): each column's run [j*R, (j+1)*R) with lanes along the run
template <class S, class StoreOp>
RC_HD void v3_first_copy_out(const float2* tr, const FftPass& P, const StoreOp& st, int batch, long long j0, int tid) {
    if (S::R % 2 == 0 && P.pair_ok) {
        constexpr int H = S::R / 2;
        for (int idx = tid; idx < H * S::T; idx += S::NT) {
            const int c = idx / H, K = 2 * (idx - c * H);
            const long long j = j0 + c;
            if (j < P.stride) st.pair(batch, j * S::R + K, tr[c * S::PITCH + K], tr[c * S::PITCH + K + 1]);
        }
    } else {
        for (int idx = tid; idx < S::R * S::T; idx += S::NT) {
            const int c = idx / S::R, K = idx - c * S::R;
            const long long j = j0 + c;
            if (j < P.stride) st(batch, j * S::R + K, tr[c * S::PITCH + K]);
        }
    }
}

}  // namespace rc
