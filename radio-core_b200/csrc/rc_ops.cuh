// rc_ops.cuh -- the per-stage arithmetic of the receive chain as load/store
// functors fused into the FFT passes and as elementwise functors.
//
// Each functor restates one step of the reference (file:line given) in fp32
// (fp64 where a short accumulation makes it free).  All are __host__ __device__
// so tests/native can replay them on the CPU.
#pragma once

#include "rc_backend.cuh"

namespace rc {

constexpr float kInvPiF = 0.31830988618379067154f;
constexpr double kPi = 3.14159265358979323846264338327950288;

// Periodic cosine window after fftshift, evaluated at bin k of an n-bin
// spectrum:  W(k) = a0 + a1*cos(2*pi*k/n + phi), phi = 0 (n even) or pi/n (odd).
// (tuner.py:155-157 'hann' -> a0=a1=0.5; decimate.py:32-33 'hamm' -> .54/.46)
struct ShiftedWindow {
    float a0, a1;
    double two_over_n;    // 2/n
    float phi_over_pi;    // phi/pi
    RC_HD float at_centered(long long kc) const {   // kc in (-n/2, n/2]
        float th = (float)((double)kc * two_over_n) + phi_over_pi;
#ifdef __CUDA_ARCH__
        return a0 + a1 * cospif(th);
#else
        return a0 + a1 * (float)cos(kPi * (double)th);
#endif
    }
};

inline ShiftedWindow make_window(bool hann, long long n) {
    ShiftedWindow w;
    w.a0 = hann ? 0.5f : 0.54f;
    w.a1 = hann ? 0.5f : 0.46f;
    w.two_over_n = 2.0 / (double)n;
    w.phi_over_pi = (n % 2 == 0) ? 0.0f : (float)(1.0 / (double)n);
    return w;
}

// ---------------------------------------------------------------------------
// LoadOp: two-sided Fourier resampling in the frequency domain.
// Restates scipy.signal.resample's two-sided branch as called by
// Tuner.run (tuner.py:159-161: roll, Hann, domain='freq') and by Decimate.run
// on complex input (decimate.py:48: Hamming, no roll): logical bin j of the
// num-point inverse FFT, gathered from the n_x-point spectrum X.
// ---------------------------------------------------------------------------
struct LoadResampleGather {
    const float2* X;           // spectrum, n_x bins
    long long x_batch_stride;  // 0 when every batch entry reads the same spectrum (Tuner)
    const long long* roll;     // per-batch roll in bins (may be null -> 0)
    long long n_x, num, m, m2;
    ShiftedWindow win;
    float scale;               // num / n_x

    RC_HD float2 src(int b, long long k, long long r) const {   // X[(k-r) mod n_x] * W(k)
        long long s = k - r;
        if (s < 0) s += n_x; else if (s >= n_x) s -= n_x;
        const long long kc = (2 * k > n_x) ? k - n_x : k;
        const float w = win.at_centered(kc) * scale;
        float2 v = ldg(X + b * x_batch_stride + s);
        return make_float2(v.x * w, v.y * w);
    }
    RC_HD float2 operator()(int b, long long j) const {
        const long long r = roll ? ldg(roll + b) : 0;
        float2 v = make_float2(0.f, 0.f);
        if (j < m2) v = src(b, j, r);
        else if (j >= num - (m - m2)) v = src(b, n_x - (num - j), r);
        if ((m & 1) == 0) {
            if (num < n_x) {
                if (j == num - m / 2) v = cadd(v, src(b, n_x - m / 2, r));
            } else if (n_x < num) {
                if (j == m / 2) v = cscale(v, 0.5f);
                else if (j == num - m / 2) v = cscale(src(b, m / 2, r), 0.5f);
            }
        }
        return v;
    }
};

// ---------------------------------------------------------------------------
// LoadOp: Tuner.run's gather (tuner.py:151-161) for the engine's batched path:
// even channel size num < n_x.  Same arithmetic as LoadResampleGather, but the
// Hann weights W(kc)*scale are read from a table built once per channel size on
// the host (fp64 cosine, rounded to fp32; all channels of a bank share it):
//   i <= num/2 :  X[(i - r) mod n_x]            * wtab[i]      (kc = i)
//   i >  num/2 :  X[(n_x - num + i - r) mod n_x] * wtab[i]      (kc = i - num)
//   i == num/2 additionally + X[(n_x - num/2 - r) mod n_x] * w_neg_half   (merged +-num/2 bins)
// roll r is normalised to [0, n_x) on the host.
// ---------------------------------------------------------------------------
struct LoadTunerGather {
    const float2* X;           // spectrum, n_x bins
    const long long* roll;     // per-batch roll in bins, in [0, n_x)
    const float* wtab;         // num weights
    long long n_x, num, half;
    float w_neg_half;

    RC_HD float2 one(long long r, long long i) const {
        long long s = (i <= half ? i : i + (n_x - num)) - r;
        if (s < 0) s += n_x;
        float2 v = cscale(ldg(X + s), ldg(wtab + i));
        if (i == half) {
            long long s2 = n_x - half - r;
            if (s2 < 0) s2 += n_x;
            v = caxpy(v, w_neg_half, ldg(X + s2));
        }
        return v;
    }
    RC_HD float2 operator()(int b, long long i) const { return one(ldg(roll + b), i); }

    // per-thread context of the register-radix kernels (n_x < 2^31: 32-bit index arithmetic)
    struct Ctx { unsigned r, nx, gap, half; long long r64; };
    RC_HD Ctx prepare(int b) const {
        Ctx c;
        c.r64 = ldg(roll + b);
        c.r = (unsigned)c.r64; c.nx = (unsigned)n_x; c.gap = (unsigned)(n_x - num); c.half = (unsigned)half;
        return c;
    }
    RC_HD float4 load2(const Ctx& c, long long i64, bool has_b) const {
        if (n_x >= (1LL << 30)) {                                       // 3*n_x must fit 32 bits below
            const float2 a = one(c.r64, i64);
            const float2 d = has_b ? one(c.r64, i64 + 1) : make_float2(0.f, 0.f);
            return make_float4(a.x, a.y, d.x, d.y);
        }
        const unsigned i = (unsigned)i64;
        // s = src(i) - r (mod n_x), src(i) = i (i <= half) or i + gap
        unsigned s = i + c.nx - c.r + (i > c.half ? c.gap : 0u);        // in [1, 3 n_x)
        if (s >= c.nx) s -= c.nx;
        if (s >= c.nx) s -= c.nx;
        unsigned s1 = s + 1u + (i == c.half ? c.gap : 0u);
        if (s1 >= c.nx) s1 -= c.nx;
        // (a 16-byte fast path for the aligned common case measured no faster at config 3 and
        // slower, through divergence, for rolls of mixed parity: configs 2 and 4)
        float2 a = cscale(ldg(X + s), ldg(wtab + i));
        float2 d = make_float2(0.f, 0.f);
        if (has_b) d = cscale(ldg(X + s1), ldg(wtab + i + 1));
        if (i == c.half || i + 1u == c.half) {                          // the merged +-num/2 bin
            unsigned s2 = 2u * c.nx - c.half - c.r;
            if (s2 >= c.nx) s2 -= c.nx;
            if (s2 >= c.nx) s2 -= c.nx;
            const float2 e = ldg(X + s2);
            if (i == c.half) a = caxpy(a, w_neg_half, e);
            else if (has_b) d = caxpy(d, w_neg_half, e);
        }
        return make_float4(a.x, a.y, d.x, d.y);
    }
};

// ---------------------------------------------------------------------------
// LoadOp: FM discriminator feeding a packed real FFT.
// fm.py:60-65: angle -> unwrap -> diff -> pad(1,0) -> /pi, i.e.
// d[0] = 0, d[n] = wrap(angle(y[n]) - angle(y[n-1]))/pi = angle(y[n]*conj(y[n-1]))/pi.
// Element i of the half-length complex sequence is (d[2i], d[2i+1]).
// ---------------------------------------------------------------------------
RC_HD float2 f4lo_(float4 v) { return make_float2(v.x, v.y); }
RC_HD float2 f4hi_(float4 v) { return make_float2(v.z, v.w); }
// atan2(y, x) / pi, branch-free: octant reduction to t = min/max in [0, 1], then
// atan(t)/(pi t) as a degree-7 polynomial in t^2 (minimax fit, max error 1.2e-8 in
// units of pi before rounding; fp32 evaluation keeps it below 1e-7).
RC_HD float atan2pi_fast(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
#ifdef __CUDA_ARCH__
    float t = __fdividef(mn, mx);
#else
    float t = mn / mx;
#endif
    if (!(mx > 0.f)) t = 0.f;
    const float s = t * t;
    float p = -0.0012906081974506378f;
    p = fmaf(p, s, 0.006959192920476198f);
    p = fmaf(p, s, -0.0177974421530962f);
    p = fmaf(p, s, 0.03069206513464451f);
    p = fmaf(p, s, -0.044272541999816895f);
    p = fmaf(p, s, 0.06349188834428787f);
    p = fmaf(p, s, -0.10609224438667297f);
    p = fmaf(p, s, 0.3183096647262573f);
    float r = t * p;
    if (ay > ax) r = 0.5f - r;
    if (x < 0.f) r = 1.0f - r;
    return copysignf(r, y);
}

RC_HD float fm_step(float2 cur, float2 prev) {
    const float re = cur.x * prev.x + cur.y * prev.y;
    const float im = cur.y * prev.x - cur.x * prev.y;
    return atan2pi_fast(im, re);
}

struct LoadDiscriminatorPacked {
    const float2* y;
    long long batch_stride;
    RC_HD float2 operator()(int b, long long i) const {
        const float2* p = y + b * batch_stride + 2 * i;
        const float2 y0 = ldg(p), y1 = ldg(p + 1);
        const float d0 = (i == 0) ? 0.f : fm_step(y0, ldg(p - 1));
        return make_float2(d0, fm_step(y1, y0));
    }
    // packed elements i and i+1: d[2i .. 2i+3] from y[2i-1 .. 2i+3]
    struct Ctx { const float2* base; };
    RC_HD Ctx prepare(int b) const { return Ctx{y + b * batch_stride}; }
    RC_HD float4 load2(const Ctx& c, long long i, bool has_b) const {
        const float2* p = c.base + 2 * i;
        float2 y0, y1, y2 = make_float2(1.f, 0.f), y3 = y2;
        if ((((size_t)p) & 15) == 0) {
            const float4 q = ldg4(p);
            y0 = f4lo_(q); y1 = f4hi_(q);
            if (has_b) { const float4 q2 = ldg4(p + 2); y2 = f4lo_(q2); y3 = f4hi_(q2); }
        } else {
            y0 = ldg(p); y1 = ldg(p + 1);
            if (has_b) { y2 = ldg(p + 2); y3 = ldg(p + 3); }
        }
        const float d0 = (i == 0) ? 0.f : fm_step(y0, ldg(p - 1));
        const float d1 = fm_step(y1, y0);
        const float d2 = has_b ? fm_step(y2, y1) : 0.f;
        const float d3 = has_b ? fm_step(y3, y2) : 0.f;
        return make_float4(d0, d1, d2, d3);
    }
};

// ---------------------------------------------------------------------------
// The same discriminator split at the phase, for the engine's batched path: the
// last pass of the channel IFFT stores angle(y[n])/pi (4 bytes per sample instead
// of the 8-byte IQ sample), the first pass of the packed real FFT takes wrapped
// differences.  fm.py:60-65: d[0] = 0, d[n] = wrap(angle(y[n]) - angle(y[n-1]))/pi.
// ---------------------------------------------------------------------------
struct StoreAngle {
    float* ang;                // [batch][n] angle / pi in (-1, 1]
    long long batch_stride;
    RC_HD void operator()(int b, long long i, float2 v) const { ang[b * batch_stride + i] = atan2pi_fast(v.y, v.x); }
    RC_HD void pair(int b, long long i, float2 v, float2 w) const {
        *(float2*)(ang + b * batch_stride + i) = make_float2(atan2pi_fast(v.y, v.x), atan2pi_fast(w.y, w.x));
    }
};

RC_HD float wrap_half_turns(float x) {          // x in (-2, 2) half-turns -> (-1, 1]
#ifdef __CUDA_ARCH__
    return x - 2.0f * rintf(0.5f * x);
#else
    return x - 2.0f * nearbyintf(0.5f * x);
#endif
}

struct LoadAnglePacked {
    const float* ang;
    long long batch_stride;
    RC_HD float2 operator()(int b, long long i) const {
        const float* p = ang + b * batch_stride + 2 * i;
        const float a0 = ldg(p), a1 = ldg(p + 1);
        const float d0 = (i == 0) ? 0.f : wrap_half_turns(a0 - ldg(p - 1));
        return make_float2(d0, wrap_half_turns(a1 - a0));
    }
    struct Ctx { const float* base; };
    RC_HD Ctx prepare(int b) const { return Ctx{ang + b * batch_stride}; }
    RC_HD float4 load2(const Ctx& c, long long i, bool has_b) const {
        const float* p = c.base + 2 * i;
        float a0, a1, a2 = 0.f, a3 = 0.f;
        if (has_b && (((size_t)p) & 15) == 0) {
#ifdef __CUDA_ARCH__
            const float4 q = __ldg((const float4*)p);
#else
            const float4 q = make_float4(p[0], p[1], p[2], p[3]);
#endif
            a0 = q.x; a1 = q.y; a2 = q.z; a3 = q.w;
        } else {
            a0 = ldg(p); a1 = ldg(p + 1);
            if (has_b) { a2 = ldg(p + 2); a3 = ldg(p + 3); }
        }
        const float d0 = (i == 0) ? 0.f : wrap_half_turns(a0 - ldg(p - 1));
        const float d1 = wrap_half_turns(a1 - a0);
        const float d2 = has_b ? wrap_half_turns(a2 - a1) : 0.f;
        const float d3 = has_b ? wrap_half_turns(a3 - a2) : 0.f;
        return make_float4(d0, d1, d2, d3);
    }
};

// Elementwise variant (used for odd sizes and by tests): d[b][n].
struct DiscriminatorEw {
    const float2* y;
    float* d;
    long long n;
    RC_HD void operator()(int b, long long i) const {
        const float2* p = y + b * n + i;
        d[b * n + i] = (i == 0) ? 0.f : fm_step(ldg(p), ldg(p - 1));
    }
};

// ---------------------------------------------------------------------------
// Real-FFT algebra on a half-length complex FFT.
// Z = FFT_h(z), z[i] = x[2i] + i x[2i+1], n = 2h.  rfft bin k (0 <= k <= h):
//   X[k] = E + W_n^k * O,  E = (Z[k] + conj Z[h-k])/2,  O = -i (Z[k] - conj Z[h-k])/2
// rtw[k] = W_n^k = exp(-2 pi i k / n).
// ---------------------------------------------------------------------------
RC_HD float2 rfft_bin(const float2* Z, long long h, long long k, const float2* rtw) {
    const float2 zk = ldg(Z + (k == h ? 0 : k));
    const float2 zm = cconj(ldg(Z + (k == 0 ? 0 : h - k)));
    const float2 e = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y + zm.y));
    const float2 dlt = make_float2(0.5f * (zk.x - zm.x), 0.5f * (zk.y - zm.y));
    const float2 o = make_float2(dlt.y, -dlt.x);   // -i * dlt
    return cadd(e, cmul(ldg(rtw + k), o));
}

// Inverse: half-spectrum Y[0..h'] of a real signal of length num = 2h'  ->
// Z'[k] = E + i*O with E = (Y[k] + conj Y[h'-k])/2, O = (Y[k] - conj Y[h'-k])/2 * W_num^{-k};
// z = IFFT_h'(Z')/h' gives z[i] = x[2i] + i x[2i+1].   itw[k] = exp(+2 pi i k / num).
RC_HD float2 irfft_pack(float2 ya, float2 yb_conj, float2 itw_k) {
    const float2 e = make_float2(0.5f * (ya.x + yb_conj.x), 0.5f * (ya.y + yb_conj.y));
    const float2 dlt = make_float2(0.5f * (ya.x - yb_conj.x), 0.5f * (ya.y - yb_conj.y));
    const float2 o = cmul(dlt, itw_k);
    return make_float2(e.x - o.y, e.y + o.x);      // E + i*O
}

// Parameters of scipy.signal.resample's rfft branch (decimate.py:48 on real
// input): fold window Wf[k] = a0 + a1*cos(phi)*cos(2 pi k/n), keep m2 bins,
// double/halve the unpaired bin m/2, scale by num/n_x.
struct RealResampleSpec {
    long long n_x, num, h, hp, m, m2;   // h = n_x/2, hp = num/2
    float a0, a1c;                      // a1c = a1*cos(phi)
    float scale;                        // (num/n_x) * (1/hp)  (inverse-FFT normalisation folded in)
    float nyq;                          // factor on bin m/2 (2, 0.5 or 1)
    const float2* rtw;                  // W_{n_x}^k, k in [0, min(m2, h+1))
    const float2* itw;                  // W_{num}^{-k}, k in [0, hp)

    // windowed, scaled rfft bin k of the resampled spectrum (0 beyond m2)
    RC_HD float2 bin(const float2* Z, long long k, int taper_pow) const {
        if (k >= m2) return make_float2(0.f, 0.f);
        float2 x = rfft_bin(Z, h, k, rtw);
        float w = a0 + a1c * ldg(rtw + k).x;
        if (taper_pow == 2) w *= w;
        w *= scale;
        if ((m & 1) == 0 && k == m / 2) w *= nyq;
        x = make_float2(x.x * w, x.y * w);
        if (k == 0 || k == hp) x.y = 0.f;      // c2r ignores Im of DC / Nyquist
        return x;
    }
};

// Elementwise: Z (FFT of packed real input, per batch h bins) -> Z' (input of the
// packed inverse FFT, per batch hp bins).  One kernel = rfft post-processing +
// window + truncation + irfft pre-processing.
struct SpecResampleEw {
    RealResampleSpec s;
    const float2* Z;      // [batch][h]
    float2* Zp;           // [batch][hp]
    RC_HD void operator()(int b, long long k) const {
        const float2* z = Z + b * s.h;
        const float2 ya = s.bin(z, k, 1);
        const float2 yb = cconj(s.bin(z, s.hp - k, 1));
        Zp[b * s.hp + k] = irfft_pack(ya, yb, ldg(s.itw + k));
    }
};

// Same-size variants (num == n_x, so hp == h) that produce bins k and h-k together: both need
// exactly Z[k] and Z[h-k], and the twiddles of h-k are -conj of those of k, so one thread loads
// four values for two outputs instead of ten.  Launch over k in [0, h/2].
RC_HD float2 rfft_from(float2 zk, float2 zm_conj, float2 tw) {
    const float2 e = make_float2(0.5f * (zk.x + zm_conj.x), 0.5f * (zk.y + zm_conj.y));
    const float2 dlt = make_float2(0.5f * (zk.x - zm_conj.x), 0.5f * (zk.y - zm_conj.y));
    return cadd(e, cmul(tw, make_float2(dlt.y, -dlt.x)));
}

struct SpecTaperPairEw {          // SpecResampleEw with s.num == s.n_x  (FM(B, B) taper, wbfm.py:42-43)
    RealResampleSpec s;
    const float2* Z;
    float2* Zp;
    RC_HD void operator()(int b, long long k) const {
        const float2* z = Z + b * s.h;
        float2* o = Zp + b * s.h;
        const long long km = s.h - k;
        if (k == 0 || k == km) {                       // DC (pairs with Nyquist) and the self-paired middle bin
            o[k] = irfft_pack(s.bin(z, k, 1), cconj(s.bin(z, s.hp - k, 1)), ldg(s.itw + k));
            return;
        }
        const float2 zk = ldg(z + k), zm = ldg(z + km);
        const float2 rt = ldg(s.rtw + k), it = ldg(s.itw + k);
        const float2 rtm = make_float2(-rt.x, rt.y), itm = make_float2(-it.x, it.y);      // -conj
        const float wk = (s.a0 + s.a1c * rt.x) * s.scale, wm = (s.a0 - s.a1c * rt.x) * s.scale;
        const float2 bk = cscale(rfft_from(zk, cconj(zm), rt), wk);
        const float2 bm = cscale(rfft_from(zm, cconj(zk), rtm), wm);
        o[k] = irfft_pack(bk, cconj(bm), it);
        o[km] = irfft_pack(bm, cconj(bk), itm);
    }
};

// WBFM audio spectra (wbfm.py:86-87 with linearity of Decimate):
// L = D(mpx)+D(lmr), R = D(mpx)-D(lmr);  rfft(mpx)[k] = X_d[k]*Wf[k]  (the
// same-size FM(B,B) resample, wbfm.py:42-43), so D(mpx) carries Wf twice.
struct SpecStereoEw {
    RealResampleSpec s;
    const float2* Zd;     // [batch][h]  FFT of packed discriminator
    const float2* Zl;     // [batch][h]  FFT of packed lmr
    float2* Zp;           // [batch][2][hp]  (L then R)
    RC_HD void operator()(int b, long long k) const {
        const float2* zd = Zd + b * s.h;
        const float2* zl = Zl + b * s.h;
        const long long kb = s.hp - k;
        const float2 ma = s.bin(zd, k, 2), la = s.bin(zl, k, 1);
        const float2 mb = cconj(s.bin(zd, kb, 2)), lb = cconj(s.bin(zl, kb, 1));
        const float2 tw = ldg(s.itw + k);
        Zp[(b * 2 + 0) * s.hp + k] = irfft_pack(cadd(ma, la), cadd(mb, lb), tw);
        Zp[(b * 2 + 1) * s.hp + k] = irfft_pack(csub(ma, la), csub(mb, lb), tw);
    }
};

// Hilbert transform spectrum (pll.py:34): analytic z = p + i*hhat with
// Hhat[k] = -i P[k] (0 < k < n/2), 0 at DC and Nyquist.  Output packed for the
// half-length inverse FFT; s.scale must be 1/h, s.m2 = h+1, window a0=1,a1c=0.
struct SpecHilbertEw {
    RealResampleSpec s;
    const float2* Z;
    float2* Zp;
    RC_HD float2 hbin(const float2* z, long long k) const {
        if (k == 0 || k == s.h) return make_float2(0.f, 0.f);
        const float2 p = rfft_bin(z, s.h, k, s.rtw);
        return make_float2(p.y * s.scale, -p.x * s.scale);     // -i * P
    }
    RC_HD void operator()(int b, long long k) const {
        const float2* z = Z + b * s.h;
        Zp[b * s.h + k] = irfft_pack(hbin(z, k), cconj(hbin(z, s.h - k)), ldg(s.itw + k));
    }
};

struct SpecHilbertPairEw {        // SpecHilbertEw, bins k and h-k together; launch over k in [0, h/2]
    RealResampleSpec s;
    const float2* Z;
    float2* Zp;
    RC_HD void operator()(int b, long long k) const {
        const float2* z = Z + b * s.h;
        float2* o = Zp + b * s.h;
        const long long km = s.h - k;
        SpecHilbertEw one{s, Z, Zp};
        if (k == 0 || k == km) { one(b, k); return; }
        const float2 zk = ldg(z + k), zm = ldg(z + km);
        const float2 rt = ldg(s.rtw + k), it = ldg(s.itw + k);
        const float2 rtm = make_float2(-rt.x, rt.y), itm = make_float2(-it.x, it.y);
        const float2 pk = rfft_from(zk, cconj(zm), rt), pm = rfft_from(zm, cconj(zk), rtm);
        const float2 hk = make_float2(pk.y * s.scale, -pk.x * s.scale);        // -i * P
        const float2 hm = make_float2(pm.y * s.scale, -pm.x * s.scale);
        o[k] = irfft_pack(hk, cconj(hm), it);
        o[km] = irfft_pack(hm, cconj(hk), itm);
    }
};

// StoreOp of the Hilbert inverse FFT: element i carries (hhat[2i], hhat[2i+1]).
// pll.py:57-58 image(2) = Im(z^2)/|z^2| = 2 p hhat / (p^2 + hhat^2);
// wbfm.py:83 lmr = image(2) * mpx * 1.0175.
struct StoreLmrPacked {
    const float* pilot;   // [batch][n]
    const float* mpx;     // [batch][n]
    float* lmr;           // [batch][n]
    long long n;
    RC_HD static float one(float p, float hh, float m) {
#ifdef __CUDA_ARCH__
        const float s2 = __fdividef(2.0f * p * hh, p * p + hh * hh);     // 0/0 -> NaN like the reference
#else
        const float s2 = (2.0f * p * hh) / (p * p + hh * hh);
#endif
        return s2 * m * 1.0175f;
    }
    RC_HD void operator()(int b, long long i, float2 v) const {
        const long long o = b * n + 2 * i;
        const float2 p = *(const float2*)(pilot + o);
        const float2 m = *(const float2*)(mpx + o);
        *(float2*)(lmr + o) = make_float2(one(p.x, v.x, m.x), one(p.y, v.y, m.y));
    }
    // packed elements i and i+1 (i even): four consecutive samples
    RC_HD void pair(int b, long long i, float2 v, float2 w) const {
        const long long o = b * n + 2 * i;
        const float4 p = *(const float4*)(pilot + o);
        const float4 m = *(const float4*)(mpx + o);
        *(float4*)(lmr + o) = make_float4(one(p.x, v.x, m.x), one(p.y, v.y, m.y), one(p.z, w.x, m.z), one(p.w, w.y, m.w));
    }
};

// StoreOp: analytic signal z[n] = p[n] + i*hhat[n] (PLL.step standalone).
struct StoreAnalyticPacked {
    const float* sig;
    float2* z;
    long long n;
    RC_HD void operator()(int b, long long i, float2 v) const {
        const long long o = b * n + 2 * i;
        z[o] = make_float2(sig[o], v.x);
        z[o + 1] = make_float2(sig[o + 1], v.y);
    }
};

// PLL.real / PLL.image (pll.py:36-58): Re or Im of z^mult / |z^mult|.
struct PllEvalEw {
    const float2* z;
    float* out;
    long long n;
    float mult;
    int imag;
    RC_HD void operator()(int b, long long i) const {
        const float2 v = z[b * n + i];
        float r;
        if (mult == 2.0f) {
            const float d = v.x * v.x + v.y * v.y;
            r = imag ? (2.0f * v.x * v.y) / d : (v.x * v.x - v.y * v.y) / d;
        } else {
            const float a = mult * atan2f(v.y, v.x);
            r = (v.x == 0.f && v.y == 0.f) ? nanf("") : (imag ? sinf(a) : cosf(a));
        }
        out[b * n + i] = r;
    }
};

// ---------------------------------------------------------------------------
// Odd-length fallback of the real resampler (decimate.py:48 on real input when
// a size is odd): full-length complex FFT of (x, 0), Hermitian re-expansion of
// the kept bins, full-length inverse FFT, real part.
// ---------------------------------------------------------------------------
struct LoadRealAsComplex {
    const float* x;
    long long batch_stride;
    RC_HD float2 operator()(int b, long long i) const { return make_float2(ldg(x + b * batch_stride + i), 0.f); }
};

struct LoadHermitianResample {
    const float2* X;          // [batch][n_x] full spectrum of the real input
    long long n_x, num, m, m2;
    float a0, a1c;            // folded window a0 + a1c*cos(2 pi k / n_x)
    double two_over_n;
    float scale, nyq;
    RC_HD float2 operator()(int b, long long j) const {
        const bool neg = 2 * j > num;
        const long long k = neg ? num - j : j;
        if (k >= m2) return make_float2(0.f, 0.f);
        float2 v = ldg(X + b * n_x + k);
        const float th = (float)((double)k * two_over_n);
#ifdef __CUDA_ARCH__
        float w = a0 + a1c * cospif(th);
#else
        float w = a0 + a1c * (float)cos(kPi * (double)th);
#endif
        w *= scale;
        if ((m & 1) == 0 && k == m / 2) w *= nyq;
        v = make_float2(v.x * w, v.y * w);
        if (k == 0 || 2 * k == num) v.y = 0.f;       // c2r ignores Im of DC / Nyquist
        if (neg) v.y = -v.y;
        return v;
    }
};

struct StoreRealPart {
    float* out;
    long long batch_stride;
    RC_HD void operator()(int b, long long i, float2 v) const { out[b * batch_stride + i] = v.x; }
};

// ---------------------------------------------------------------------------
// Sub-band combine: the last, radix-G step of an N-point forward FFT whose first steps were G
// independent M-point FFTs (M = N / G) of the commutated input x_g[m] = x[G m + g], one per GPU
// (Tuner.load, tuner.py:137-138, sharded over G ranks -- SURVEY.md 7.3-1 / 8e):
//   X[k0 + M k1] = sum_g W_G^{g k1} * W_N^{g k0} * F_g[k0],     k0 in [0, M), k1 in [0, G).
// A rank holds F_g[k0] of ALL g for its piece k0 in [k0_base, k0_base + P) (after the first
// exchange) and produces the G bins k0 + M k1 of every k0 of the piece: Y[k1][j], j = k0 - k0_base.
// Twiddle W_N^{k0} from an fp64 sincospi, its powers by fp64 recurrence, rounded to fp32 once --
// the same accuracy as the inter-pass twiddles of the FFT engine.
// ---------------------------------------------------------------------------
// (A two-bins-per-thread variant with 16-byte loads and stores was measured in round 2: the combine
// alone got 7 % faster, but the overlapped step 6-9 % SLOWER at 2 and 8 GPUs -- one bin per thread kept.)
template <int G>
struct SubbandCombineEw {
    const float2* F;      // [G][P]
    float2* Y;            // [G][P]
    long long P, k0_base;
    double minus_two_over_n;
    RC_HD void operator()(int, long long j) const {
        float2 v[G];
        double sn, cs;
        const double th = (double)(k0_base + j) * minus_two_over_n;
#ifdef __CUDA_ARCH__
        sincospi(th, &sn, &cs);
#else
        sn = sin(kPi * th); cs = cos(kPi * th);
#endif
        const double2 w1 = make_double2(cs, sn);
        double2 w = w1;
        v[0] = ldg(F + j);
#pragma unroll
        for (int g = 1; g < G; g++) {
            v[g] = cmul(ldg(F + (long long)g * P + j), make_float2((float)w.x, (float)w.y));
            if (g + 1 < G) w = cmul64(w, w1);
        }
        Dft<G, -1>::run(v);
#pragma unroll
        for (int k1 = 0; k1 < G; k1++) Y[(long long)k1 * P + j] = v[k1];
    }
};

// The same combine with the second exchange fused into its stores: bin k0 + M k1 is written
// straight into the sub-band buffer of every rank whose channels read it -- this GPU's or an NVLink
// peer's (mapped symmetric memory) -- instead of into a local array that is copied afterwards.
// tab: per k1 up to kScatterSegs segments [lo, hi) of j with the address of the segment's first bin.
constexpr int kScatterSegs = 4;
struct ScatterTable {
    int n[kMaxRanks];
    long long lo[kMaxRanks][kScatterSegs], hi[kMaxRanks][kScatterSegs];
    float2* dst[kMaxRanks][kScatterSegs];
};
template <int G>
struct SubbandCombineScatterEw {
    const float2* F;      // [G][P]
    long long P, k0_base;
    double minus_two_over_n;
    ScatterTable tab;
    RC_HD void operator()(int, long long j) const {
        float2 v[G];
        double sn, cs;
        const double th = (double)(k0_base + j) * minus_two_over_n;
#ifdef __CUDA_ARCH__
        sincospi(th, &sn, &cs);
#else
        sn = sin(kPi * th); cs = cos(kPi * th);
#endif
        const double2 w1 = make_double2(cs, sn);
        double2 w = w1;
        v[0] = ldg(F + j);
#pragma unroll
        for (int g = 1; g < G; g++) {
            v[g] = cmul(ldg(F + (long long)g * P + j), make_float2((float)w.x, (float)w.y));
            if (g + 1 < G) w = cmul64(w, w1);
        }
        Dft<G, -1>::run(v);
#pragma unroll
        for (int k1 = 0; k1 < G; k1++) {
#pragma unroll
            for (int sgm = 0; sgm < kScatterSegs; sgm++)
                if (sgm < tab.n[k1] && j >= tab.lo[k1][sgm] && j < tab.hi[k1][sgm]) tab.dst[k1][sgm][j - tab.lo[k1][sgm]] = v[k1];
        }
    }
};

// ---------------------------------------------------------------------------
// Zero-phase FIR (bandpass.py:72 filtfilt(b, 1, x), padtype 'odd').
// With an FIR the lfilter_zi start-up terms only touch the first len(b)-1
// samples of the 3*len(b) extension, which filtfilt crops, so the result is
// exactly  out[n] = sum_j g[j] * xe[n + j - K],  g = b (*) reversed(b),
// K = len(b)-1, xe = x extended by odd reflection about both end samples.
// ---------------------------------------------------------------------------
// Fast interior path for the 81-tap pilot filter of WBFM (K = 40): g is symmetric
// (g[K-d] = g[K+d]), so
//   out[n] = g[K] x[n] + sum_{d=1..40} g[K+d] (x[n-d] + x[n+d]).
// The pair sums and the products run in fp32 (FADD + FFMA, five partial sums of eight distances
// each), the partial sums are combined in fp64: 10 fp64-pipe operations per sample instead of 81
// DFMA (the all-fp64 kernel ran at 0.15 of the HBM roofline, bound by the fp64 pipe).  Error of
// the pilot against the exact sum: ~5e-7 of its amplitude -- the level of the fp32 FFTs around it.
// Chunks (kFirChunk outputs) that touch the odd-extended block ends keep the exact fp64 path.
constexpr int kFoldK = 40;
constexpr int kFoldGroup = 8;
struct FoldTaps {
    float g[kFoldK + 1];      // g[d] = (float) gtaps[K + d], d = 1..40 (g[0] unused)
    double gc;                // centre tap gtaps[K]
    int on;
};
RC_HD float fold_mul_add(float g, float a, float b, float s) {
#ifdef __CUDA_ARCH__
    return __fmaf_rn(g, __fadd_rn(a, b), s);
#else
    return fmaf(g, a + b, s);
#endif
}

struct FiltFiltEw {
    const float* x;
    float* out;
    const double* g;     // 2K+1 autocorrelation taps
    long long n;
    int K;
    FoldTaps fold;       // fold.on: interior chunks use the folded fp32 path (K == kFoldK only)
    RC_HD double xe(const float* xb, long long i) const {
        if (i < 0) return 2.0 * (double)xb[0] - (double)xb[-i];
        if (i >= n) return 2.0 * (double)xb[n - 1] - (double)xb[2 * (n - 1) - i];
        return (double)xb[i];
    }
    // does the kFirChunk-output chunk holding sample i stay clear of both block ends?
    RC_HD bool chunk_interior(long long i) const {
        const long long c0 = i / 1024 * 1024;
        return fold.on && c0 >= K && c0 + 1024 + K <= n;
    }
    RC_HD float folded(const float* xb, long long i) const {
        double acc = fold.gc * (double)xb[i];
        for (int d0 = 1; d0 <= kFoldK; d0 += kFoldGroup) {
            float sgrp = 0.f;
            for (int d = d0; d < d0 + kFoldGroup; d++) sgrp = fold_mul_add(fold.g[d], xb[i - d], xb[i + d], sgrp);
            acc += (double)sgrp;
        }
        return (float)acc;
    }
    RC_HD void operator()(int b, long long i) const {
        const float* xb = x + b * n;
        if (chunk_interior(i)) { out[b * n + i] = folded(xb, i); return; }
        double acc = 0.0;
        if (i >= K && i + K < n) {
            for (int j = 0; j <= 2 * K; j++) acc += g[j] * (double)xb[i + j - K];
        } else {
            for (int j = 0; j <= 2 * K; j++) acc += g[j] * xe(xb, i + j - K);
        }
        out[b * n + i] = (float)acc;
    }
};

// ---------------------------------------------------------------------------
// Audio epilogue: stateful FIR de-emphasis (deemphasis.py:64 lfilter with
// carried zi), block mean removal and clip (mfm.py:64-65, wbfm.py:97-100).
// Phase functions are per-thread; the CTA owns one channel (nch audio channels).
// ---------------------------------------------------------------------------
struct EpilogueParams {
    const float* in;      // [batch][nch][A]   resampled audio
    float* out;           // [batch][A][nch]   interleaved, final
    double* zi;           // [batch][nch][K]   carried filter state (K = ntaps-1)
    double* zi_next;      // scratch of the same shape
    const float* taps;    // ntaps FIR taps (float32 values, as the reference stores them)
    long long A;
    int nch, ntaps;
    int deemph;           // 0: no filter (plain FM)
    int dc_clip;          // 1: subtract block mean and clip to +-0.999
    // dc_clip: bin 0 of the packed inverse-FFT input of every audio channel ([batch*nch] complex64 at
    // stride dc_stride), from which the block sum of `in` follows without reading the block:
    // sum_n in[n] = dc_scale * (re + im)   (packed real transform: sum_i z[i] = hp * Z'[0])
    const float2* dc;
    long long dc_stride;
    double dc_scale;
};

// Block mean WITHOUT a pass over the filtered block (mfm.py:64, wbfm.py:97: `x - mean(x)` over the whole
// block, both stereo channels together).  With y = lfilter(b, 1, a, zi):
//   sum_n y[n] = sum_k b[k] * (S - T_k) + sum_{n<K} zi[n],   S = sum(a),  T_k = sum of the last k samples of a
//             = S * c[0] - sum_{m=1..K} a[A-m] * c[m] + sum zi,          c[m] = sum_{k>=m} b[k]
// -- S comes from bin 0 of the spectrum the audio was synthesised from, the rest is K = 50 samples and the
// carried state, so every CTA of the FIR kernel can subtract the mean and clip in the same pass: the
// fp64 staging array and the second kernel of round 1 are gone.  (The mean so obtained differs from
// the mean of the rounded samples by ~1e-10 of full scale.)
struct FirSuffixParam { double c[56]; };       // c[m] = sum_{k >= m} taps[k], m <= K

RC_HD double epi_channel_sum_lane(const EpilogueParams& p, const FirSuffixParam& sp, long long bc, int lane, int nlanes) {
    const float* a = p.in + bc * p.A;
    const int K = p.deemph ? p.ntaps - 1 : 0;
    double acc = 0.0;
    for (int m = 1 + lane; m <= K && m <= p.A; m += nlanes) acc -= (double)a[p.A - m] * sp.c[m];
    for (int n = lane; n < K && n < p.A; n += nlanes) acc += p.zi[bc * K + n];
    if (lane == 0) {
        const float2 z = ldg(p.dc + bc * p.dc_stride);
        acc += p.dc_scale * ((double)z.x + (double)z.y) * sp.c[0];
    }
    return acc;
}

RC_HD double epi_fir(const EpilogueParams& p, int b, int ch, long long n) {
    const float* a = p.in + ((long long)b * p.nch + ch) * p.A;
    if (!p.deemph) return (double)a[n];
    const int K = p.ntaps - 1;
    double acc = 0.0;
    const int kmax = (n < K) ? (int)n : K;
    for (int k = 0; k <= kmax; k++) acc += (double)ldg(p.taps + k) * (double)a[n - k];
    if (n < K) acc += p.zi[((long long)b * p.nch + ch) * K + n];
    return acc;
}

// zf[i] = sum_{k>i} b[k] a[A-(k-i)]  (+ zi[i+A] when the block is shorter than the filter)
RC_HD double epi_next_state(const EpilogueParams& p, int b, int ch, int i) {
    const float* a = p.in + ((long long)b * p.nch + ch) * p.A;
    const int K = p.ntaps - 1;
    double acc = 0.0;
    for (int k = i + 1; k <= K; k++) {
        const long long idx = p.A - (k - i);
        if (idx >= 0) acc += (double)ldg(p.taps + k) * (double)a[idx];
    }
    if (i + p.A < K) acc += p.zi[((long long)b * p.nch + ch) * K + i + p.A];
    return acc;
}

RC_HD float epi_finish(const EpilogueParams& p, double v, double mean) {
    if (!p.dc_clip) return (float)v;
    v -= mean;
    v = v < -0.999 ? -0.999 : (v > 0.999 ? 0.999 : v);
    return (float)v;
}

// ---- shared FIR machinery of the epilogue and of the zero-phase filter ----------
// A CTA of kFirThreads threads produces kFirChunk consecutive outputs of
//   acc[n] = sum_{j < ntaps} taps[j] * xs[n + j]          (correlation form)
// from a shared-memory window xs (fp64, converted once on load).  Each thread owns
// 8 consecutive outputs and slides a 16-value register window over the taps, so
// one shared-memory read feeds eight fp64 FMAs.
constexpr int kFirThreads = 128;
constexpr int kFirPer = 8;
constexpr int kFirChunk = kFirThreads * kFirPer;     // outputs per CTA
constexpr int kFirMaxTaps = 136;                     // padded tap count supported (multiple of 8)

// window slot of logical sample i: one pad slot every 8 samples, so that threads reading
// 8-sample groups 64 bytes apart land on different banks
RC_HD int fir_slot(int i) { return i + (i >> 3); }
constexpr int kFirSlots = kFirChunk + kFirMaxTaps + 8 + (kFirChunk + kFirMaxTaps + 8) / 8 + 1;

#if defined(__CUDACC__) && !defined(RC_EMULATE)
// NT8 > 0: tap count known at compile time (fully unrolled: the sliding window lives in renamed
// registers, no moves); NT8 == 0: runtime count.
template <int NT8>
__device__ __forceinline__ void fir_window8_t(const double* xs, const double* taps, int ntaps8, int o, double acc[kFirPer]) {
    double w[16];
    const double* x0 = xs + fir_slot(o);           // o is a multiple of 8: groups are contiguous
#pragma unroll
    for (int i = 0; i < 8; i++) w[i] = x0[i];
    const int n8 = NT8 > 0 ? NT8 : ntaps8;
#pragma unroll
    for (int j0 = 0; j0 < n8; j0 += 8) {
        const double* x1 = xs + fir_slot(o + j0 + 8);
#pragma unroll
        for (int i = 0; i < 8; i++) w[8 + i] = x1[i];
#pragma unroll
        for (int jj = 0; jj < 8; jj++) {
            const double t = taps[j0 + jj];
#pragma unroll
            for (int r = 0; r < kFirPer; r++) acc[r] = fma(t, w[r + jj], acc[r]);
        }
#pragma unroll
        for (int i = 0; i < 8; i++) w[i] = w[8 + i];
    }
}
// Taps handed to the kernel by value: after full unrolling every tap is a constant-bank operand of
// its DFMA, so the inner loop is 64 DFMA per 8 shared-memory loads (the shared-memory tap table of
// the generic path costs one more load per 8 DFMA and makes the loop LSU-bound).
struct FirTapsParam {
    double t[88];
    int n8;                    // 0: not provided (use the device tap table)
};
template <int NT8>
__device__ __forceinline__ void fir_window8_c(const double* xs, const FirTapsParam& tp, int o, double acc[kFirPer]) {
    double w[16];
    const double* x0 = xs + fir_slot(o);
#pragma unroll
    for (int i = 0; i < 8; i++) w[i] = x0[i];
#pragma unroll
    for (int j0 = 0; j0 < NT8; j0 += 8) {
        const double* x1 = xs + fir_slot(o + j0 + 8);
#pragma unroll
        for (int i = 0; i < 8; i++) w[8 + i] = x1[i];
#pragma unroll
        for (int jj = 0; jj < 8; jj++) {
#pragma unroll
            for (int r = 0; r < kFirPer; r++) acc[r] = fma(tp.t[j0 + jj], w[r + jj], acc[r]);
        }
#pragma unroll
        for (int i = 0; i < 8; i++) w[i] = w[8 + i];
    }
}

__device__ __forceinline__ void fir_window8(const double* xs, const double* taps, int ntaps8, int o, double acc[kFirPer]) {
    if (ntaps8 == 88) fir_window8_t<88>(xs, taps, ntaps8, o, acc);          // 81-tap pilot filter (41 (*) 41)
    else if (ntaps8 == 56) fir_window8_t<56>(xs, taps, ntaps8, o, acc);     // 51-tap de-emphasis
    else fir_window8_t<0>(xs, taps, ntaps8, o, acc);
}

// The audio epilogue in ONE kernel: de-emphasis FIR of one chunk, block mean (analytic), clip, interleave.
static __global__ void __launch_bounds__(kFirThreads) epi_fir_kernel(const EpilogueParams p, int nchunks,
                                                                     const __grid_constant__ FirTapsParam ctaps,
                                                                     const __grid_constant__ FirSuffixParam suffix) {
    __shared__ double xs[kFirSlots];
    __shared__ double tp[kFirMaxTaps];
    __shared__ double mean_s;
    const int bc = blockIdx.y, chunk = blockIdx.x, tid = threadIdx.x;
    const int K = p.deemph ? p.ntaps - 1 : 0;
    const int ntaps8 = p.deemph ? (p.ntaps + 7) / 8 * 8 : 8;
    const long long n0 = (long long)chunk * kFirChunk;
    const float* a = p.in + (long long)bc * p.A;
    // xs[i] = x[n0 - K + i] (0 before the block: the carried state zi covers that part)
    for (int i = tid; i < kFirChunk + ntaps8 + 8; i += kFirThreads) {
        const long long idx = n0 - K + i;
        xs[fir_slot(i)] = (idx >= 0 && idx < p.A) ? (double)a[idx] : 0.0;
    }
    for (int i = tid; i < ntaps8; i += kFirThreads) {
        double t = 0.0;
        if (p.deemph) { if (i <= K) t = (double)p.taps[K - i]; }       // reversed taps: y[n] = sum_k b[k] x[n-k]
        else if (i == 0) t = 1.0;
        tp[i] = t;
    }
    const int bq = bc / p.nch, chq = bc - bq * p.nch;
    if (p.dc_clip && tid < 32) {
        // warp 0: the block mean of this channel (all nch audio channels together), analytically
        double t = 0.0;
        for (int ch = 0; ch < p.nch; ch++) t += epi_channel_sum_lane(p, suffix, (long long)bq * p.nch + ch, tid, 32);
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) t += __shfl_xor_sync(0xffffffffu, t, s);
        if (tid == 0) mean_s = t / (double)(p.A * p.nch);
    }
    __syncthreads();
    double acc[kFirPer];
#pragma unroll
    for (int r = 0; r < kFirPer; r++) acc[r] = 0.0;
    const int o = tid * kFirPer;
    if (ctaps.n8 == 56 && ntaps8 == 56) fir_window8_c<56>(xs, ctaps, o, acc);
    else fir_window8(xs, tp, ntaps8, o, acc);
    const double mean = p.dc_clip ? mean_s : 0.0;
#pragma unroll
    for (int r = 0; r < kFirPer; r++) {
        const long long n = n0 + o + r;
        if (n < p.A) {
            double v = acc[r];
            if (p.deemph && n < K) v += p.zi[(long long)bc * K + n];
            if (p.dc_clip) p.out[((long long)bq * p.A + n) * p.nch + chq] = epi_finish(p, v, mean);    // interleaved [A][nch]
            else p.out[(long long)bc * p.A + n] = (float)v;             // nch == 1 without mean removal
        }
    }
    if (p.deemph && chunk == nchunks - 1 && tid < K) {
        const int b = bc / p.nch, ch = bc - b * p.nch;
        p.zi_next[(long long)bc * K + tid] = epi_next_state(p, b, ch, tid);
    }
}

// Zero-phase FIR (FiltFiltEw) with shared-memory staging of the odd-extended input.
static __global__ void __launch_bounds__(kFirThreads) filtfilt_kernel(const FiltFiltEw f, const __grid_constant__ FirTapsParam ctaps) {
    __shared__ double xs[kFirSlots];
    __shared__ double tp[kFirMaxTaps];
    const int b = blockIdx.y, tid = threadIdx.x;
    const int ntaps = 2 * f.K + 1, ntaps8 = (ntaps + 7) / 8 * 8;
    const long long n0 = (long long)blockIdx.x * kFirChunk;
    const float* xb = f.x + (long long)b * f.n;
    for (int i = tid; i < kFirChunk + ntaps8 + 8; i += kFirThreads) {
        const long long idx = n0 - f.K + i;
        xs[fir_slot(i)] = (idx < f.n + f.K) ? f.xe(xb, idx) : 0.0;
    }
    for (int i = tid; i < ntaps8; i += kFirThreads) tp[i] = i < ntaps ? f.g[i] : 0.0;
    __syncthreads();
    double acc[kFirPer];
#pragma unroll
    for (int r = 0; r < kFirPer; r++) acc[r] = 0.0;
    const int o = tid * kFirPer;
    if (ctaps.n8 == 88 && ntaps8 == 88) fir_window8_c<88>(xs, ctaps, o, acc);
    else fir_window8(xs, tp, ntaps8, o, acc);
#pragma unroll
    for (int r = 0; r < kFirPer; r++) {
        const long long n = n0 + o + r;
        if (n < f.n) f.out[(long long)b * f.n + n] = (float)acc[r];
    }
}

// Folded pilot filter (FiltFiltEw::folded) for interior chunks: fp32 window in shared memory
// (one 4-float pad per 32 so that the 16-byte reads of 8 threads, 32 bytes apart, cover all banks),
// two register windows per thread sliding outwards from the centre, 8 outputs per thread.
RC_HD int fold_slot(int i) { return i + 4 * (i >> 5); }
constexpr int kFoldWin = kFirChunk + 2 * kFoldK + 16;
static __global__ void __launch_bounds__(kFirThreads) filtfilt_fold_kernel(const FiltFiltEw f, const __grid_constant__ FirTapsParam ctaps) {
    __shared__ __align__(16) float xf[kFoldWin + 4 * (kFoldWin / 32) + 8];
    __shared__ double xs[kFirSlots];
    __shared__ double tp[kFirMaxTaps];
    const int b = blockIdx.y, tid = threadIdx.x;
    const long long n0 = (long long)blockIdx.x * kFirChunk;
    const float* xb = f.x + (long long)b * f.n;
    if (!(n0 >= f.K && n0 + kFirChunk + f.K <= f.n)) {
        // a block end is in reach: exact fp64 path with the odd extension (as filtfilt_kernel)
        const int ntaps = 2 * f.K + 1, ntaps8 = (ntaps + 7) / 8 * 8;
        for (int i = tid; i < kFirChunk + ntaps8 + 8; i += kFirThreads) {
            const long long idx = n0 - f.K + i;
            xs[fir_slot(i)] = (idx < f.n + f.K) ? f.xe(xb, idx) : 0.0;
        }
        for (int i = tid; i < ntaps8; i += kFirThreads) tp[i] = i < ntaps ? f.g[i] : 0.0;
        __syncthreads();
        double acc[kFirPer];
#pragma unroll
        for (int r = 0; r < kFirPer; r++) acc[r] = 0.0;
        const int o = tid * kFirPer;
        if (ctaps.n8 == 88 && ntaps8 == 88) fir_window8_c<88>(xs, ctaps, o, acc);
        else fir_window8(xs, tp, ntaps8, o, acc);
#pragma unroll
        for (int r = 0; r < kFirPer; r++) {
            const long long n = n0 + o + r;
            if (n < f.n) f.out[(long long)b * f.n + n] = (float)acc[r];
        }
        return;
    }
    // window[i] = x[n0 - 40 + i], i in [0, 1024 + 80): 16-byte loads where the source allows
    const float* src = xb + (n0 - kFoldK);
    if ((((size_t)src) & 15) == 0) {
        for (int i = tid; i < (kFirChunk + 2 * kFoldK) / 4; i += kFirThreads)
            *(float4*)(xf + fold_slot(4 * i)) = __ldg((const float4*)src + i);
    } else {
        for (int i = tid; i < kFirChunk + 2 * kFoldK; i += kFirThreads) xf[fold_slot(i)] = __ldg(src + i);
    }
    __syncthreads();
    const int c = tid * kFirPer + kFoldK;          // window index of this thread's first output (multiple of 8)
    float lw[16], rw[16];
    {
        const float4 a = *(const float4*)(xf + fold_slot(c)), d = *(const float4*)(xf + fold_slot(c + 4));
        lw[8] = rw[0] = a.x; lw[9] = rw[1] = a.y; lw[10] = rw[2] = a.z; lw[11] = rw[3] = a.w;
        lw[12] = rw[4] = d.x; lw[13] = rw[5] = d.y; lw[14] = rw[6] = d.z; lw[15] = rw[7] = d.w;
    }
    double acc[kFirPer];
#pragma unroll
    for (int r = 0; r < kFirPer; r++) acc[r] = f.fold.gc * (double)rw[r];
#pragma unroll
    for (int m = 0; m < kFoldK / kFoldGroup; m++) {
        // left window: x[c - 8(m+1) .. +15], its upper half is the previous group's lower half;
        // right window: x[c + 8m .. +15], its lower half is the previous group's upper half
#pragma unroll
        for (int i = 0; i < 8; i++) { lw[8 + i] = m == 0 ? lw[8 + i] : lw[i]; rw[i] = m == 0 ? rw[i] : rw[8 + i]; }
        {
            const int lb = c - 8 * (m + 1), rb = c + 8 * m + 8;
            const float4 a = *(const float4*)(xf + fold_slot(lb)), d = *(const float4*)(xf + fold_slot(lb + 4));
            lw[0] = a.x; lw[1] = a.y; lw[2] = a.z; lw[3] = a.w; lw[4] = d.x; lw[5] = d.y; lw[6] = d.z; lw[7] = d.w;
            const float4 e = *(const float4*)(xf + fold_slot(rb)), h = *(const float4*)(xf + fold_slot(rb + 4));
            rw[8] = e.x; rw[9] = e.y; rw[10] = e.z; rw[11] = e.w; rw[12] = h.x; rw[13] = h.y; rw[14] = h.z; rw[15] = h.w;
        }
        float sg[kFirPer];
#pragma unroll
        for (int r = 0; r < kFirPer; r++) sg[r] = 0.f;
#pragma unroll
        for (int jj = 0; jj < kFoldGroup; jj++) {
            const float g = f.fold.g[8 * m + 1 + jj];
#pragma unroll
            for (int r = 0; r < kFirPer; r++) sg[r] = fold_mul_add(g, lw[7 - jj + r], rw[1 + jj + r], sg[r]);
        }
#pragma unroll
        for (int r = 0; r < kFirPer; r++) acc[r] += (double)sg[r];
    }
    float* o = f.out + (long long)b * f.n + n0 + tid * kFirPer;
    if ((((size_t)o) & 15) == 0) {
        *(float4*)o = make_float4((float)acc[0], (float)acc[1], (float)acc[2], (float)acc[3]);
        *(float4*)(o + 4) = make_float4((float)acc[4], (float)acc[5], (float)acc[6], (float)acc[7]);
    } else {
#pragma unroll
        for (int r = 0; r < kFirPer; r++) o[r] = (float)acc[r];
    }
}
#endif

inline int epi_chunks(long long A) { return (int)((A + kFirChunk - 1) / kFirChunk); }

// taps_host: the same ntaps float taps on the host (optional; enables the constant-operand FIR loop)
inline FirSuffixParam fir_suffix_sums(const float* taps_host, int ntaps, bool deemph) {
    FirSuffixParam sp;
    memset(&sp, 0, sizeof(sp));
    if (!deemph) { sp.c[0] = 1.0; return sp; }          // identity "filter": y = a
    double acc = 0.0;
    for (int m = ntaps - 1; m >= 0; m--) { acc += (double)taps_host[m]; sp.c[m] = acc; }
    return sp;
}

// taps_host: the same ntaps float taps on the host (required with dc_clip; enables the constant-operand FIR loop)
inline cudaError_t launch_epilogue(const EpilogueParams& p, int batch, cudaStream_t stream, const float* taps_host = nullptr) {
    if (p.dc_clip && (!taps_host || !p.dc || p.ntaps > 56)) return cudaErrorInvalidValue;
    const FirSuffixParam sp = taps_host ? fir_suffix_sums(taps_host, p.ntaps, p.deemph != 0) : FirSuffixParam{};
#ifdef RC_EMULATE
    const long long total = p.A * p.nch;
    for (int b = 0; b < batch; b++) {
        double mean = 0.0;
        if (p.dc_clip) {
            double sum = 0.0;
            for (int ch = 0; ch < p.nch; ch++)
                for (int lane = 0; lane < 32; lane++) sum += epi_channel_sum_lane(p, sp, (long long)b * p.nch + ch, lane, 32);
            mean = sum / (double)total;
        }
        if (p.deemph) {
            const int K = p.ntaps - 1;
            for (int e = 0; e < K * p.nch; e++)
                p.zi_next[(long long)b * p.nch * K + e] = epi_next_state(p, b, e / K, e % K);
        }
        for (long long e = 0; e < total; e++)
            p.out[(long long)b * total + e] = epi_finish(p, epi_fir(p, b, (int)(e % p.nch), e / p.nch), mean);
    }
    (void)stream;
    return cudaSuccess;
#else
    if (batch <= 0) return cudaSuccess;
    if (p.ntaps > kFirMaxTaps - 8 || (!p.dc_clip && p.nch != 1)) return cudaErrorInvalidValue;
    const int nchunks = epi_chunks(p.A);
    ProfileScope scope(p.dc_clip ? "demod.deemph_mean_clip" : "demod.deemph_fir", 8.0 * (double)p.A * p.nch * batch, stream);
    FirTapsParam ct;
    memset(&ct, 0, sizeof(ct));
    if (taps_host && p.deemph && p.ntaps > 48 && p.ntaps <= 56) {
        const int K = p.ntaps - 1;
        for (int i = 0; i <= K; i++) ct.t[i] = (double)taps_host[K - i];      // reversed, as in the kernel's table
        ct.n8 = 56;
    }
    epi_fir_kernel<<<dim3((unsigned)nchunks, (unsigned)(batch * p.nch)), kFirThreads, 0, stream>>>(p, nchunks, ct, sp);
    return cudaGetLastError();
#endif
}

// g_host: the 2K+1 taps on the host (optional; enables the constant-operand FIR loop)
inline cudaError_t launch_filtfilt(const FiltFiltEw& f, int batch, cudaStream_t stream, const char* tag = "filtfilt",
                                   const double* g_host = nullptr) {
#ifdef RC_EMULATE
    return launch_ew(f.n, batch, f, stream);
#else
    if (batch <= 0 || f.n <= 0) return cudaSuccess;
    if (2 * f.K + 1 > kFirMaxTaps - 8) return launch_ew(f.n, batch, f, stream, tag, 8.0 * (double)f.n * batch);
    ProfileScope scope(tag, 8.0 * (double)f.n * batch, stream);
    FirTapsParam ct;
    memset(&ct, 0, sizeof(ct));
    const int ntaps = 2 * f.K + 1;
    if (g_host && ntaps > 80 && ntaps <= 88) {
        for (int i = 0; i < ntaps; i++) ct.t[i] = g_host[i];
        ct.n8 = 88;
    }
    const dim3 grid((unsigned)((f.n + kFirChunk - 1) / kFirChunk), (unsigned)batch);
    if (f.fold.on && f.K == kFoldK) filtfilt_fold_kernel<<<grid, kFirThreads, 0, stream>>>(f, ct);
    else filtfilt_kernel<<<grid, kFirThreads, 0, stream>>>(f, ct);
    return cudaGetLastError();
#endif
}

}  // namespace rc
