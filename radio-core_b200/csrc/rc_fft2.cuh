// rc_fft2.cuh -- register-radix FFT pass kernels (the fast path of rc_fft.cuh).
//
// Same pass semantics as the generic shared-memory kernel in rc_fft.cuh (one
// Stockham step of radix R on a tile of T = 16 adjacent columns), but the
// R-point transforms are evaluated as 2 or 3 in-register radix stages
// R = R0*R1*R2 chosen at compile time:
//
//   stage 0   global -> registers (R0 strided loads per butterfly, all issued
//             before use, inter-pass fp64 twiddle recurrence) -> radix-R0 ->
//             intra-pass twiddle -> shared memory
//   stage 1   shared -> registers -> radix-R1 -> twiddle -> shared (in place)
//   stage 2   shared -> registers -> radix-R2 -> global (through the StoreOp)
//
// so an element makes two shared-memory round trips per pass instead of one
// per radix-<=25 stage, every global access is a full 128-byte line (16
// float2 columns) and all shared-memory accesses are conflict-free (a
// half-warp always touches 16 consecutive float2 of one row).
// The first pass of a plan (Ns == 1) writes each column's R outputs as one
// contiguous run; it re-orders through shared memory (pitch 17) so that the
// run is stored with lanes along the run.
//
// Digit bookkeeping for one column (t = input row, K = output row):
//   t = t0*(R1*R2) + t1*R2 + t2        K = k0 + R0*k1 + R0*R1*k2
//   shared row rho = d0*(R1*R2) + d1*R2 + d2, digit d_s holds t_s before
//   stage s and k_s after it.
//   W_R^{tK} = W_R0^{t0k0} * W_R^{t1*k0*R2} * W_R1^{t1k1} * W_R^{t2*(k0+R0k1)} * W_R2^{t2k2}
#pragma once

#include "rc_fft.cuh"

namespace rc {

constexpr int kV2T = 16;

template <int R0_, int R1_, int R2_, int NT_, int MINB_>
struct V2Sched {
    static constexpr int R0 = R0_, R1 = R1_, R2 = R2_, NT = NT_, MINB = MINB_;
    static constexpr int R = R0_ * R1_ * R2_;
    static constexpr int U = R1_ * R2_;
    static constexpr int RL = R2_;                       // radix of the last stage
    static constexpr int NBL = R0_ * R1_ * kV2T;         // butterflies of the last stage
    static constexpr int ITL = (NBL + NT_ - 1) / NT_;
    static constexpr int HOLD = ITL * R2_;               // float2 registers held across the re-order (first pass)
    static_assert(NT_ % kV2T == 0, "threads must be a multiple of the tile width");
};

template <class S, bool FIRST> struct V2Smem {
    static constexpr int PITCH = FIRST ? kV2T + 1 : kV2T;
    static constexpr int ELEMS = S::R * PITCH + S::R;    // tile + W_R table
};

// W_R table: forward sign in global memory, conjugated here for the inverse.
template <class S, int SIGN>
RC_HD void v2_load_table(float2* tw, const FftPass& P, int tid) {
    for (int i = tid; i < S::R; i += S::NT) {
        float2 w = ldg(P.twR + i);
        if (SIGN > 0) w.y = -w.y;
        tw[i] = w;
    }
}

template <class S, int SIGN, bool FIRST, class LoadOp>
RC_HD void v2_stage0(float2* sm, const float2* tw, const FftPass& P, const LoadOp& ld, int batch,
                     long long j0, int tid) {
    constexpr int T = kV2T, PITCH = V2Smem<S, FIRST>::PITCH;
    constexpr int NB = S::U * T, ITER = (NB + S::NT - 1) / S::NT;
    const int c = tid & (T - 1);
    const long long j = j0 + c;
    const bool active = j < P.stride;
    unsigned long long kj = 0;
    double2 ws = make_double2(1.0, 0.0);
    if (!FIRST && active) {
        kj = (unsigned long long)(j % P.Ns);
        ws = fft_tw64(P, (unsigned long long)S::U * kj);
    }
#pragma unroll
    for (int it = 0; it < ITER; it++) {
        const int b = tid + it * S::NT;
        if (NB % S::NT != 0 && b >= NB) break;
        const int u = b / T;
        float2 v[S::R0];
        if (active) {
#pragma unroll
            for (int t0 = 0; t0 < S::R0; t0++) v[t0] = ld(batch, j + (long long)(t0 * S::U + u) * P.stride);
            if (!FIRST) {
                double2 w = fft_tw64(P, (unsigned long long)u * kj);
#pragma unroll
                for (int t0 = 0; t0 < S::R0; t0++) {
                    v[t0] = cmul(v[t0], make_float2((float)w.x, (float)(SIGN < 0 ? w.y : -w.y)));
                    if (t0 + 1 < S::R0) w = cmul64(w, ws);
                }
            }
        } else {
#pragma unroll
            for (int t0 = 0; t0 < S::R0; t0++) v[t0] = make_float2(0.f, 0.f);
        }
        Dft<S::R0, SIGN>::run(v);
        // intra-pass twiddle on the outputs: two-stage W_R^{t2*k0} (u = t2), three-stage W_R^{t1*k0*R2}
        const int tq = (S::R1 == 1) ? u : (u / S::R2) * S::R2;
        if (tq > 0) {
#pragma unroll
            for (int k0 = 1; k0 < S::R0; k0++) v[k0] = cmul(v[k0], tw[tq * k0]);
        }
#pragma unroll
        for (int k0 = 0; k0 < S::R0; k0++) sm[(k0 * S::U + u) * PITCH + c] = v[k0];
    }
}

// middle stage (three-stage schedules only): radix R1 over digit d1, in place
template <class S, int SIGN, bool FIRST>
RC_HD void v2_stage1(float2* sm, const float2* tw, int tid) {
    constexpr int T = kV2T, PITCH = V2Smem<S, FIRST>::PITCH;
    constexpr int NB = S::R0 * S::R2 * T, ITER = (NB + S::NT - 1) / S::NT;
    const int c = tid & (T - 1);
#pragma unroll
    for (int it = 0; it < ITER; it++) {
        const int b = tid + it * S::NT;
        if (NB % S::NT != 0 && b >= NB) break;
        const int q = b / T;
        const int t2 = q % S::R2, k0 = q / S::R2;
        const int base = k0 * S::U + t2;
        float2 v[S::R1];
#pragma unroll
        for (int t1 = 0; t1 < S::R1; t1++) v[t1] = sm[(base + t1 * S::R2) * PITCH + c];
        Dft<S::R1, SIGN>::run(v);
        if (t2 > 0) {
#pragma unroll
            for (int k1 = 0; k1 < S::R1; k1++) {
                const int e = t2 * (k0 + S::R0 * k1);
                if (e > 0) v[k1] = cmul(v[k1], tw[e]);
            }
        }
#pragma unroll
        for (int k1 = 0; k1 < S::R1; k1++) sm[(base + k1 * S::R2) * PITCH + c] = v[k1];
    }
}

// last stage, passes after the first: radix R2 over digit d2, results straight to global
template <class S, int SIGN, class StoreOp>
RC_HD void v2_last_direct(const float2* sm, const FftPass& P, const StoreOp& st, int batch, long long j0, int tid) {
    constexpr int T = kV2T, PITCH = V2Smem<S, false>::PITCH;
    constexpr int NB = S::NBL, ITER = S::ITL;
    const int c = tid & (T - 1);
    const long long j = j0 + c;
    if (j >= P.stride) return;
    const long long qn = j / P.Ns;
    const long long obase = qn * P.Ns * S::R + (j - qn * P.Ns);
#pragma unroll
    for (int it = 0; it < ITER; it++) {
        const int b = tid + it * S::NT;
        if (NB % S::NT != 0 && b >= NB) break;
        const int q = b / T;                 // q = k0*R1 + k1
        const int k1 = q % S::R1, k0 = q / S::R1;
        const int base = k0 * S::U + k1 * S::R2;
        float2 v[S::R2];
#pragma unroll
        for (int t2 = 0; t2 < S::R2; t2++) v[t2] = sm[(base + t2) * PITCH + c];
        Dft<S::R2, SIGN>::run(v);
        const int K0 = k0 + S::R0 * k1;
#pragma unroll
        for (int k2 = 0; k2 < S::R2; k2++) st(batch, obase + (long long)(K0 + S::R0 * S::R1 * k2) * P.Ns, v[k2]);
    }
}

// last stage of the first pass, part A: read + radix into held registers
template <class S, int SIGN>
RC_HD void v2_last_first_a(const float2* sm, float2* hold, int tid) {
    constexpr int T = kV2T, PITCH = V2Smem<S, true>::PITCH;
    const int c = tid & (T - 1);
#pragma unroll
    for (int it = 0; it < S::ITL; it++) {
        const int b = tid + it * S::NT;
        float2* v = hold + it * S::R2;
        if (S::NBL % S::NT != 0 && b >= S::NBL) break;
        const int q = b / T;
        const int k1 = q % S::R1, k0 = q / S::R1;
        const int base = k0 * S::U + k1 * S::R2;
#pragma unroll
        for (int t2 = 0; t2 < S::R2; t2++) v[t2] = sm[(base + t2) * PITCH + c];
        Dft<S::R2, SIGN>::run(v);
    }
}
// part B (after a barrier): write the held results at their natural row K
template <class S>
RC_HD void v2_last_first_b(float2* sm, const float2* hold, int tid) {
    constexpr int T = kV2T, PITCH = V2Smem<S, true>::PITCH;
    const int c = tid & (T - 1);
#pragma unroll
    for (int it = 0; it < S::ITL; it++) {
        const int b = tid + it * S::NT;
        const float2* v = hold + it * S::R2;
        if (S::NBL % S::NT != 0 && b >= S::NBL) break;
        const int q = b / T;
        const int k1 = q % S::R1, k0 = q / S::R1;
        const int K0 = k0 + S::R0 * k1;
#pragma unroll
        for (int k2 = 0; k2 < S::R2; k2++) sm[(K0 + S::R0 * S::R1 * k2) * PITCH + c] = v[k2];
    }
}
// part C (after a barrier): each column's run [j*R, (j+1)*R) with lanes along the run
template <class S, class StoreOp>
RC_HD void v2_first_copy_out(const float2* sm, const FftPass& P, const StoreOp& st, int batch, long long j0, int tid) {
    constexpr int T = kV2T, PITCH = V2Smem<S, true>::PITCH;
    for (int idx = tid; idx < S::R * T; idx += S::NT) {
        const int K = idx % S::R, c = idx / S::R;
        const long long j = j0 + c;
        if (j < P.stride) st(batch, j * S::R + K, sm[K * PITCH + c]);
    }
}

}  // namespace rc
