// rc_fft.cuh -- batched mixed-radix complex FFT for sm_100a (sizes 2^a 3^b 5^c).
//
// Why it exists: every stage of the reference's receive chain is a Fourier
// operation on block sizes that are never powers of two -- Tuner.load is one
// N-point FFT (reference radiocore/tools/tuner.py:137-138), Tuner.run an
// inverse B-point FFT of gathered bins (tuner.py:159-161), Decimate an
// rfft/irfft pair (radiocore/analog/decimate.py:48), PLL.step a Hilbert
// transform (radiocore/analog/pll.py:34).  This header is the one FFT engine
// all of them are built on.
//
// Design (B200-first, no tensor cores -- the work is butterflies on fp32 pairs):
//   * A transform of length n is split into 1..3 "passes" n = R1*R2*R3.  A pass
//     is a Stockham step with a large radix R (up to ~1500): input element
//     (j, t) lives at j + t*(n/R), output element (j, K) at
//     expand(j, Ns, R) + K*Ns, Ns = product of earlier R's -- so the data comes
//     out in natural order without a transpose kernel.
//   * One CTA owns a tile of T adjacent columns j (T = 16 -> every global access
//     is a full 128-byte line) and all R rows; the T independent R-point FFTs
//     run in shared memory (pitch T+1 -> conflict-free both along c and along K)
//     as in-place decimation-in-time stages with radices {2,3,4,5,8,10,16,25}
//     evaluated in registers.
//   * Inter-pass twiddles W_{Ns*R}^{t*k} are generated per thread by an fp64
//     recurrence seeded from two small fp64 tables (exact to ~1e-16), then
//     rounded to fp32 once -- no sincos in the inner loop, no accuracy loss.
//   * The first pass reads through a LoadOp functor and the last pass writes
//     through a StoreOp functor, so gathers/windows/scales fuse into the FFT.
//
// Every per-thread phase is a __host__ __device__ function so the exact index
// arithmetic can be replayed on the CPU (tests/native/emulate_fft.cu) -- the
// build container has no GPU.
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <functional>
#include <map>
#include <string>
#include <vector>

#ifndef RC_HD
#define RC_HD __host__ __device__ __forceinline__
#endif

namespace rc {

// ------------------------------------------------------------------ complex
// On sm_100a a float2 lives in an aligned register pair and FADD2 / FMUL2 / FFMA2
// work on both halves at once, with per-half negate and half-swap operand
// modifiers: a complex add is one instruction, a complex multiply two, and a
// multiplication by +-i folds into the operand modifiers of its consumer.
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000) && !defined(RC_NO_PACKED)
#define RC_PACKED 1
#endif
RC_HD float2 cadd(float2 a, float2 b) {
#ifdef RC_PACKED
    return __fadd2_rn(a, b);
#else
    return make_float2(a.x + b.x, a.y + b.y);
#endif
}
RC_HD float2 csub(float2 a, float2 b) {
#ifdef RC_PACKED
    return __fadd2_rn(a, make_float2(-b.x, -b.y));
#else
    return make_float2(a.x - b.x, a.y - b.y);
#endif
}
RC_HD float2 cmul(float2 a, float2 b) {
#ifdef RC_PACKED
    const float2 p = __fmul2_rn(a, make_float2(b.x, b.x));
    return __ffma2_rn(make_float2(a.y, a.x), make_float2(-b.y, b.y), p);
#else
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
#endif
}
RC_HD float2 cscale(float2 a, float s) {
#ifdef RC_PACKED
    return __fmul2_rn(a, make_float2(s, s));
#else
    return make_float2(a.x * s, a.y * s);
#endif
}
// a + s*b with a real scale (both halves)
RC_HD float2 caxpy(float2 a, float s, float2 b) {
#ifdef RC_PACKED
    return __ffma2_rn(b, make_float2(s, s), a);
#else
    return make_float2(a.x + s * b.x, a.y + s * b.y);
#endif
}
RC_HD float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
RC_HD double2 cmul64(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// multiply by SIGN*i (forward transform SIGN = -1  ->  times -i)
template <int SIGN> RC_HD float2 mul_si(float2 a) {
    return SIGN < 0 ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x);
}

template <typename T> RC_HD T ldg(const T* p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}

// ------------------------------------------------- register-level butterflies
template <int N> struct SmallTw;
#include "rc_small_tw.inc"

// v *= W_N^k (forward sign) or its conjugate; k, N compile-time after unrolling.
template <int N, int SIGN> RC_HD float2 small_twiddle(float2 v, int k) {
    k %= N;
    if (k == 0) return v;
    if (N % 4 == 0 && k == N / 4) return mul_si<SIGN>(v);
    if (N % 2 == 0 && k == N / 2) return make_float2(-v.x, -v.y);
    if (N % 4 == 0 && k == 3 * N / 4) return mul_si<-SIGN>(v);
    return cmul(v, make_float2(SmallTw<N>::c(k), SIGN * SmallTw<N>::s(k)));
}

template <int R, int SIGN> struct Dft;

template <int SIGN> struct Dft<2, SIGN> {
    static RC_HD void run(float2* v) {
        float2 a = v[0], b = v[1];
        v[0] = cadd(a, b);
        v[1] = csub(a, b);
    }
};

template <int SIGN> struct Dft<3, SIGN> {
    static RC_HD void run(float2* v) {
        const float h = 0.86602540378443865f;   // sin(2 pi / 3)
        float2 t1 = cadd(v[1], v[2]);
        float2 t2 = csub(v[1], v[2]);
        float2 m = caxpy(v[0], -0.5f, t1);
        float2 r = mul_si<SIGN>(cscale(t2, h));  // SIGN*i*h*(v1-v2)
        v[0] = cadd(v[0], t1);
        v[1] = cadd(m, r);
        v[2] = csub(m, r);
    }
};

template <int SIGN> struct Dft<4, SIGN> {
    static RC_HD void run(float2* v) {
        float2 a = cadd(v[0], v[2]), b = csub(v[0], v[2]);
        float2 c = cadd(v[1], v[3]), d = mul_si<SIGN>(csub(v[1], v[3]));
        v[0] = cadd(a, c);
        v[1] = cadd(b, d);
        v[2] = csub(a, c);
        v[3] = csub(b, d);
    }
};

template <int SIGN> struct Dft<5, SIGN> {
    static RC_HD void run(float2* v) {
        const float c1 = 0.30901699437494742f, c2 = -0.80901699437494742f;
        const float s1 = 0.95105651629515357f, s2 = 0.58778525229247313f;
        float2 a1 = cadd(v[1], v[4]), b1 = csub(v[1], v[4]);
        float2 a2 = cadd(v[2], v[3]), b2 = csub(v[2], v[3]);
        float2 x0 = v[0];
        v[0] = cadd(x0, cadd(a1, a2));
        float2 p1 = caxpy(caxpy(x0, c1, a1), c2, a2);
        float2 p2 = caxpy(caxpy(x0, c2, a1), c1, a2);
        float2 q1 = mul_si<SIGN>(caxpy(cscale(b1, s1), s2, b2));
        float2 q2 = mul_si<SIGN>(caxpy(cscale(b1, s2), -s1, b2));
        v[1] = cadd(p1, q1);
        v[4] = csub(p1, q1);
        v[2] = cadd(p2, q2);
        v[3] = csub(p2, q2);
    }
};

// Cooley-Tukey composite in registers: n = R2*n1 + n2, k = k1 + R1*k2.
template <int R1, int R2, int SIGN> RC_HD void dft_composite(float2* v) {
    constexpr int N = R1 * R2;
    float2 t[N];
#pragma unroll
    for (int n2 = 0; n2 < R2; n2++) {
        float2 a[R1];
#pragma unroll
        for (int n1 = 0; n1 < R1; n1++) a[n1] = v[R2 * n1 + n2];
        Dft<R1, SIGN>::run(a);
#pragma unroll
        for (int k1 = 0; k1 < R1; k1++) t[k1 * R2 + n2] = small_twiddle<N, SIGN>(a[k1], n2 * k1);
    }
#pragma unroll
    for (int k1 = 0; k1 < R1; k1++) {
        float2 b[R2];
#pragma unroll
        for (int n2 = 0; n2 < R2; n2++) b[n2] = t[k1 * R2 + n2];
        Dft<R2, SIGN>::run(b);
#pragma unroll
        for (int k2 = 0; k2 < R2; k2++) v[k1 + R1 * k2] = b[k2];
    }
}
template <int SIGN> struct Dft<8, SIGN> { static RC_HD void run(float2* v) { dft_composite<2, 4, SIGN>(v); } };
template <int SIGN> struct Dft<10, SIGN> { static RC_HD void run(float2* v) { dft_composite<2, 5, SIGN>(v); } };
template <int SIGN> struct Dft<16, SIGN> { static RC_HD void run(float2* v) { dft_composite<4, 4, SIGN>(v); } };
template <int SIGN> struct Dft<25, SIGN> { static RC_HD void run(float2* v) { dft_composite<5, 5, SIGN>(v); } };
template <int SIGN> struct Dft<6, SIGN> { static RC_HD void run(float2* v) { dft_composite<2, 3, SIGN>(v); } };
template <int SIGN> struct Dft<12, SIGN> { static RC_HD void run(float2* v) { dft_composite<4, 3, SIGN>(v); } };
template <int SIGN> struct Dft<15, SIGN> { static RC_HD void run(float2* v) { dft_composite<3, 5, SIGN>(v); } };
template <int SIGN> struct Dft<20, SIGN> { static RC_HD void run(float2* v) { dft_composite<4, 5, SIGN>(v); } };

// ---------------------------------------------------------------------------
// Optional per-kernel timing (rc_profile_*): CUDA events recorded on the launch
// stream around every kernel, aggregated by kernel tag.  `bytes` is the
// kernel's compulsory HBM traffic (its own reads + writes, each element once).
// ---------------------------------------------------------------------------
struct ProfileRecord { std::string tag; double bytes; cudaEvent_t a, b; };
struct Profiler {
    bool on = false;
    std::vector<ProfileRecord> recs;
    long long launches = 0;          // counted even when timing is off
};
inline Profiler& profiler() { static Profiler p; return p; }

struct ProfileScope {
    cudaStream_t st;
    bool live = false;
    ProfileScope(const char* tag, double bytes, cudaStream_t stream) : st(stream) {
        Profiler& p = profiler();
        p.launches++;
#ifndef RC_EMULATE
        if (p.on) {
            ProfileRecord r;
            r.tag = tag; r.bytes = bytes;
            cudaEventCreate(&r.a); cudaEventCreate(&r.b);
            cudaEventRecord(r.a, st);
            p.recs.push_back(r);
            live = true;
        }
#else
        (void)tag; (void)bytes;
#endif
    }
    ~ProfileScope() {
#ifndef RC_EMULATE
        if (live) cudaEventRecord(profiler().recs.back().b, st);
#endif
    }
};

// --------------------------------------------------------------- pass record
constexpr int kMaxStages = 16;
constexpr int kMaxPasses = 4;
constexpr int kSmemBudgetElems = 25600;      // float2 elements per CTA (200 KiB)

struct FftPass {
    int R;                 // transform length handled in shared memory
    int T, logT;           // columns per CTA tile (power of two)
    int nstage;
    int radix[kMaxStages]; // product = R, DIT order (stage 0 has no twiddles)
    long long n;           // full transform length
    long long Ns;          // product of the R's of earlier passes
    long long stride;      // n / R : input row stride == number of columns
    unsigned long long M;  // Ns * R : modulus of the inter-pass twiddle
    int tw_shift;          // W_M^q = lo[q & mask] * hi[q >> shift]
    unsigned tw_mask;
    const double2* tw_lo;
    const double2* tw_hi;
    const float2* twR;     // W_R^m forward, m in [0, R)
    const int* pos;        // time index t -> shared-memory slot (digit reversal)
    int threads;
    int smem_elems;
    int fast_id;           // >= 0: register-radix schedule of rc_fft3.cuh, -1: generic kernel
    int pair_ok;           // outputs of this pass may be written as aligned element pairs (rc_fft3.cuh)
};

RC_HD int fft_phys(const FftPass& P, int p, int c) {
    return P.T > 1 ? p * (P.T + 1) + c : p + (p >> 4);
}

RC_HD double2 fft_tw64(const FftPass& P, unsigned long long q) {
    if (q >= P.M) q %= P.M;
    double2 a = ldg(P.tw_lo + (unsigned)(q & P.tw_mask));
    double2 b = ldg(P.tw_hi + (unsigned)(q >> P.tw_shift));
    return cmul64(a, b);
}

// ------------------------------------------------------------ generic I/O ops
// 16-byte global accesses (two adjacent complex64); the caller guarantees alignment
RC_HD float4 ldg4(const float2* p) {
#ifdef __CUDA_ARCH__
    return __ldg((const float4*)p);
#else
    return make_float4(p[0].x, p[0].y, p[1].x, p[1].y);
#endif
}
RC_HD void stg4(float2* p, float2 a, float2 b) {
#ifdef __CUDA_ARCH__
    *(float4*)p = make_float4(a.x, a.y, b.x, b.y);
#else
    p[0] = a; p[1] = b;
#endif
}

struct LoadC64 {          // contiguous complex64 batches
    const float2* p;
    long long batch_stride;
    RC_HD float2 operator()(int b, long long i) const { return ldg(p + b * batch_stride + i); }
    // elements i and i+1 (the second only when has_b); 16-byte load when the address allows it
    struct Ctx { const float2* base; };
    RC_HD Ctx prepare(int b) const { return Ctx{p + b * batch_stride}; }
    RC_HD float4 load2(const Ctx& c, long long i, bool has_b) const {
        const float2* q = c.base + i;
        if (has_b && (((size_t)q) & 15) == 0) return ldg4(q);
        const float2 a = ldg(q);
        const float2 d = has_b ? ldg(q + 1) : make_float2(0.f, 0.f);
        return make_float4(a.x, a.y, d.x, d.y);
    }
};
struct StoreC64 {
    float2* p;
    long long batch_stride;
    float scale;
    RC_HD void operator()(int b, long long i, float2 v) const {
        p[b * batch_stride + i] = cscale(v, scale);
    }
    // elements i and i+1, i even and the batch base 16-byte aligned (FftPass::pair_ok)
    RC_HD void pair(int b, long long i, float2 v, float2 w) const {
        stg4(p + b * batch_stride + i, cscale(v, scale), cscale(w, scale));
    }
};

// complex64 store that keeps only the elements outside [lo, hi) of each batch entry: the
// half-length spectrum of a real signal that is about to be truncated to its lowest bins
// (Decimate, decimate.py:48) is only ever read near both ends.  lo and hi are even.
struct StoreC64Win {
    float2* p;
    long long batch_stride;
    long long lo, hi;
    RC_HD void operator()(int b, long long i, float2 v) const {
        if (i < lo || i >= hi) p[b * batch_stride + i] = v;
    }
    RC_HD void pair(int b, long long i, float2 v, float2 w) const {
        if (i < lo || i >= hi) stg4(p + b * batch_stride + i, v, w);
    }
};

// complex64 store that SCATTERS the output over up to kMaxRanks destination buffers: element i goes
// to base[i / P][i % P].  The bases are device pointers of this GPU *or of its NVLink peers* (mapped
// symmetric memory): the last pass of the local transform of the sharded Tuner.load writes every
// piece straight into the rank that combines it -- the first exchange of radiocore/tools/sharding.py
// is the store of the FFT itself, not a separate copy.  P is even (pairs never straddle two pieces).
constexpr int kMaxRanks = 16;
struct StoreScatterC64 {
    float2* base[kMaxRanks];
    unsigned P;
    RC_HD void operator()(int, long long i, float2 v) const {
        const unsigned u = (unsigned)i, p = u / P;
        base[p][u - p * P] = v;
    }
    RC_HD void pair(int, long long i, float2 v, float2 w) const {
        const unsigned u = (unsigned)i, p = u / P;
        stg4(base[p] + (u - p * P), v, w);
    }
};

// ------------------------------------------------------------- pass phases
template <class LoadOp, int SIGN>
RC_HD void fft_pass_load(float2* sm, const FftPass& P, const LoadOp& ld, int batch,
                         long long j0, int tid, int nthreads) {
    const int c = tid & (P.T - 1);
    const int t0 = tid >> P.logT;
    const int dt = nthreads >> P.logT;
    const long long j = j0 + c;
    const bool active = j < P.stride;
    const bool use_tw = P.Ns > 1;
    double2 w = make_double2(1.0, 0.0), ws = w;
    if (use_tw && active) {
        unsigned long long k = (unsigned long long)(j % P.Ns);
        w = fft_tw64(P, (unsigned long long)t0 * k);
        ws = fft_tw64(P, (unsigned long long)dt * k);
    }
    for (int t = t0; t < P.R; t += dt) {
        float2 v = make_float2(0.f, 0.f);
        if (active) {
            v = ld(batch, j + (long long)t * P.stride);
            if (use_tw) {
                v = cmul(v, make_float2((float)w.x, (float)(SIGN < 0 ? w.y : -w.y)));
                w = cmul64(w, ws);
            }
        }
        sm[fft_phys(P, ldg(P.pos + t), c)] = v;
    }
}

template <int R, int SIGN>
RC_HD void fft_stage_item(float2* sm, const FftPass& P, int Lprev, int item) {
    const int c = item & (P.T - 1);
    const int bf = item >> P.logT;
    const int blk = bf / Lprev;
    const int u = bf - blk * Lprev;
    const int base = blk * (Lprev * R) + u;
    const int twstep = (P.R / (Lprev * R)) * u;   // W_{L_s}^{u} expressed on the W_R table
    float2 v[R];
#pragma unroll
    for (int m = 0; m < R; m++) v[m] = sm[fft_phys(P, base + m * Lprev, c)];
    if (u > 0) {
#pragma unroll
        for (int m = 1; m < R; m++) {
            float2 w = ldg(P.twR + twstep * m);
            if (SIGN > 0) w.y = -w.y;
            v[m] = cmul(v[m], w);
        }
    }
    Dft<R, SIGN>::run(v);
#pragma unroll
    for (int k = 0; k < R; k++) sm[fft_phys(P, base + k * Lprev, c)] = v[k];
}

template <int R, int SIGN>
RC_HD void fft_stage(float2* sm, const FftPass& P, int Lprev, int tid, int nthreads) {
    const int nitems = (P.R / R) << P.logT;
    for (int item = tid; item < nitems; item += nthreads) fft_stage_item<R, SIGN>(sm, P, Lprev, item);
}

template <int SIGN>
RC_HD void fft_stage_dispatch(float2* sm, const FftPass& P, int radix, int Lprev, int tid, int nthreads) {
    switch (radix) {
        case 2: fft_stage<2, SIGN>(sm, P, Lprev, tid, nthreads); break;
        case 3: fft_stage<3, SIGN>(sm, P, Lprev, tid, nthreads); break;
        case 4: fft_stage<4, SIGN>(sm, P, Lprev, tid, nthreads); break;
        case 5: fft_stage<5, SIGN>(sm, P, Lprev, tid, nthreads); break;
        case 8: fft_stage<8, SIGN>(sm, P, Lprev, tid, nthreads); break;
        case 10: fft_stage<10, SIGN>(sm, P, Lprev, tid, nthreads); break;
        case 16: fft_stage<16, SIGN>(sm, P, Lprev, tid, nthreads); break;
        case 25: fft_stage<25, SIGN>(sm, P, Lprev, tid, nthreads); break;
        default: break;
    }
}

template <class StoreOp>
RC_HD void fft_pass_store(const float2* sm, const FftPass& P, const StoreOp& st, int batch,
                          long long j0, int tid, int nthreads) {
    if (P.Ns == 1) {
        // First pass: the tile's output is the contiguous run [j0*R, (j0+T)*R), K fastest.
        for (int c = 0; c < P.T; c++) {
            const long long j = j0 + c;
            if (j >= P.stride) break;
            for (int K = tid; K < P.R; K += nthreads) st(batch, j * P.R + K, sm[fft_phys(P, K, c)]);
        }
    } else {
        const int c = tid & (P.T - 1);
        const long long j = j0 + c;
        if (j >= P.stride) return;
        const long long q = j / P.Ns;
        const long long base = q * P.Ns * P.R + (j - q * P.Ns);
        for (int K = tid >> P.logT; K < P.R; K += nthreads >> P.logT)
            st(batch, base + (long long)K * P.Ns, sm[fft_phys(P, K, c)]);
    }
}

#if defined(__CUDACC__) && !defined(RC_EMULATE)
template <class LoadOp, class StoreOp, int SIGN>
__global__ void __launch_bounds__(512) fft_pass_kernel(const FftPass P, const LoadOp ld, const StoreOp st) {
    extern __shared__ float2 rc_fft_smem[];
    float2* sm = rc_fft_smem;
    const int batch = blockIdx.y + blockIdx.z * gridDim.y;
    const long long j0 = (long long)blockIdx.x * P.T;
    fft_pass_load<LoadOp, SIGN>(sm, P, ld, batch, j0, threadIdx.x, blockDim.x);
    __syncthreads();
    int Lprev = 1;
    for (int s = 0; s < P.nstage; s++) {
        fft_stage_dispatch<SIGN>(sm, P, P.radix[s], Lprev, threadIdx.x, blockDim.x);
        __syncthreads();
        Lprev *= P.radix[s];
    }
    fft_pass_store<StoreOp>(sm, P, st, batch, j0, threadIdx.x, blockDim.x);
}
#endif

// ------------------------------------------------------------------- planning
struct TableStore;
struct FftPlan {
    long long n = 0;
    int npass = 0;                 // generic shared-memory passes (any 2^a 3^b 5^c size)
    FftPass pass[kMaxPasses];
    int nfast = 0;                 // register-radix passes (rc_fft3.cuh) when n splits into curated lengths
    FftPass fast[kMaxPasses];
    int max_passes() const { return npass > nfast ? npass : nfast; }
    struct TableStore* store = nullptr;
};

// Memory source for the plan tables: device (product) or host (CPU emulation).
struct TableStore {
    bool on_device;
    std::vector<void*> owned;
    std::map<std::string, const void*> cache;
    explicit TableStore(bool dev) : on_device(dev) {}
    ~TableStore() { release(); }
    void release() {
        for (void* p : owned) {
            if (on_device) cudaFree(p); else free(p);
        }
        owned.clear();
        cache.clear();
    }
    const void* put(const std::string& key, const void* host, size_t bytes, cudaError_t* err) {
        auto it = cache.find(key);
        if (it != cache.end()) return it->second;
        void* p = nullptr;
        if (on_device) {
            cudaError_t e = cudaMalloc(&p, bytes);
            if (e == cudaSuccess) e = cudaMemcpy(p, host, bytes, cudaMemcpyHostToDevice);
            if (e != cudaSuccess) { if (err) *err = e; return nullptr; }
        } else {
            p = malloc(bytes);
            memcpy(p, host, bytes);
        }
        owned.push_back(p);
        cache[key] = p;
        return p;
    }
    bool has(const std::string& key) const { return cache.count(key) != 0; }
    // uninitialised scratch owned by the store (zero-filled)
    void* alloc(size_t bytes, cudaError_t* err) {
        void* p = nullptr;
        if (on_device) {
            cudaError_t e = cudaMalloc(&p, bytes);
            if (e == cudaSuccess) e = cudaMemset(p, 0, bytes);
            if (e != cudaSuccess) { if (err) *err = e; return nullptr; }
        } else {
            p = calloc(1, bytes);
        }
        owned.push_back(p);
        return p;
    }
};

inline bool fft_size_supported(long long n) {
    if (n < 1) return false;
    for (int p : {2, 3, 5}) while (n % p == 0) n /= p;
    return n == 1;
}

// Radix schedule for one in-shared-memory transform of length R.
inline std::vector<int> fft_radix_schedule(int R) {
    int e2 = 0, e3 = 0, e5 = 0;
    while (R % 2 == 0) { R /= 2; e2++; }
    while (R % 3 == 0) { R /= 3; e3++; }
    while (R % 5 == 0) { R /= 5; e5++; }
    std::vector<int> r;
    while (e5 >= 2) { r.push_back(25); e5 -= 2; }
    if (e5 == 1) { if (e2 >= 1) { r.push_back(10); e2--; } else r.push_back(5); }
    while (e2 >= 4) { r.push_back(16); e2 -= 4; }
    if (e2 == 3) r.push_back(8);
    else if (e2 == 2) r.push_back(4);
    else if (e2 == 1) r.push_back(2);
    while (e3 > 0) { r.push_back(3); e3--; }
    if (r.empty()) r.push_back(1);
    return r;
}

inline int fft_max_R(int T) { return T > 1 ? kSmemBudgetElems / (T + 1) : (kSmemBudgetElems * 16) / 17 - 1; }

// Split n into the fewest pass lengths that fit shared memory; balanced factors.
inline bool fft_choose_passes(long long n, std::vector<int>& Rs, int& T) {
    Rs.clear();
    if (n <= fft_max_R(1)) { Rs.push_back((int)n); T = 1; return true; }
    std::vector<long long> divs;
    for (long long d = 1; d * d <= n; d++)
        if (n % d == 0) { divs.push_back(d); if (d != n / d) divs.push_back(n / d); }
    for (int Tc : {16, 8}) {
        const long long lim = fft_max_R(Tc);
        // two passes
        long long best = 0;
        for (long long d : divs) {
            long long e = n / d;
            if (d <= lim && e <= lim && d >= e) { if (best == 0 || d < best) best = d; }
        }
        if (best) { Rs = {(int)best, (int)(n / best)}; T = Tc; return true; }
    }
    for (int Tc : {16, 8}) {
        const long long lim = fft_max_R(Tc);
        long long bestmax = 0; long long ba = 0, bb = 0, bc = 0;
        for (long long a : divs) {
            if (a > lim) continue;
            long long rest = n / a;
            for (long long b : divs) {
                if (b > a || b > lim || rest % b) continue;
                long long c = rest / b;
                if (c > b || c > lim) continue;
                if (bestmax == 0 || a < bestmax) { bestmax = a; ba = a; bb = b; bc = c; }
            }
        }
        if (bestmax) { Rs = {(int)ba, (int)bb, (int)bc}; T = Tc; return true; }
    }
    return false;
}

// Fill one pass record (tables are shared through the store).  fast_id >= 0 marks a
// register-radix pass (rc_fft2.cuh): no digit-reversal table, T = 16.
inline cudaError_t fft_fill_pass(FftPass& P, long long n, int R, int T, long long Ns, TableStore& store, int fast_id) {
    cudaError_t err = cudaSuccess;
    memset(&P, 0, sizeof(P));
    P.fast_id = fast_id;
    P.R = R;
    P.T = T;
    P.logT = 0;
    while ((1 << P.logT) < T) P.logT++;
    std::vector<int> rad = fft_radix_schedule(P.R);
    if (rad.size() == 1 && rad[0] == 1) rad.clear();
    P.nstage = (int)rad.size();
    for (int s = 0; s < P.nstage; s++) P.radix[s] = rad[s];
    P.n = n;
    P.Ns = Ns;
    P.stride = n / P.R;
    P.M = (unsigned long long)Ns * P.R;
    // W_R table (forward)
    {
        std::string key = "twR:" + std::to_string(P.R);
        if (!store.has(key)) {
            std::vector<float2> t(P.R);
            for (int m = 0; m < P.R; m++) {
                long double a = -2.0L * 3.14159265358979323846264338327950288L * m / P.R;
                t[m] = make_float2((float)cosl(a), (float)sinl(a));
            }
            store.put(key, t.data(), t.size() * sizeof(float2), &err);
        }
        P.twR = (const float2*)store.put(key, nullptr, 0, &err);
    }
    // digit-reversal table (generic kernel only)
    if (fast_id < 0) {
        std::string key = "pos:" + std::to_string(P.R);
        if (!store.has(key)) {
            std::vector<int> pos(P.R);
            for (int t = 0; t < P.R; t++) {
                int tmp = t, p = 0;
                std::vector<int> L(P.nstage + 1, 1);
                for (int s = 0; s < P.nstage; s++) L[s + 1] = L[s] * P.radix[s];
                for (int s = P.nstage - 1; s >= 0; s--) {
                    int m = tmp % P.radix[s];
                    tmp /= P.radix[s];
                    p += m * L[s];
                }
                pos[t] = p;
            }
            store.put(key, pos.data(), pos.size() * sizeof(int), &err);
        }
        P.pos = (const int*)store.put(key, nullptr, 0, &err);
    }
    // inter-pass twiddle tables W_M, M = Ns*R (fp64, forward)
    if (Ns > 1) {
        int bits = 0;
        while ((1ULL << bits) < P.M) bits++;
        P.tw_shift = (bits + 1) / 2;
        P.tw_mask = (1u << P.tw_shift) - 1u;
        std::string key = "twM:" + std::to_string(P.M);
        if (!store.has(key + ":lo")) {
            size_t nlo = (size_t)1 << P.tw_shift;
            size_t nhi = (size_t)((P.M >> P.tw_shift) + 1);
            std::vector<double2> lo(nlo), hi(nhi);
            const long double tau = 2.0L * 3.14159265358979323846264338327950288L;
            for (size_t q = 0; q < nlo; q++) {
                long double a = -tau * (long double)q / (long double)P.M;
                lo[q] = make_double2((double)cosl(a), (double)sinl(a));
            }
            for (size_t q = 0; q < nhi; q++) {
                unsigned long long qq = ((unsigned long long)q << P.tw_shift) % P.M;
                long double a = -tau * (long double)qq / (long double)P.M;
                hi[q] = make_double2((double)cosl(a), (double)sinl(a));
            }
            store.put(key + ":lo", lo.data(), nlo * sizeof(double2), &err);
            store.put(key + ":hi", hi.data(), nhi * sizeof(double2), &err);
        }
        P.tw_lo = (const double2*)store.put(key + ":lo", nullptr, 0, &err);
        P.tw_hi = (const double2*)store.put(key + ":hi", nullptr, 0, &err);
    }
    P.pair_ok = (Ns == 1) ? (P.R % 2 == 0) : (Ns % 2 == 0);
    P.smem_elems = T > 1 ? P.R * (T + 1) : P.R + (P.R >> 4) + 1;
    long long work = (long long)P.R * T;
    int th = (int)((work / 8 + 31) / 32 * 32);
    if (th < 64) th = 64;
    if (th > 512) th = 512;
    P.threads = th;
    return err;
}

// ---- register-radix schedules (rc_fft3.cuh):
//      X(id, R0, R1, R2, threads, min CTAs/SM, column pairs per tile, role) ----
// R = R0*R1*R2; R1 == 1 marks a two-stage schedule.  A thread owns two columns, so a stage has
// (R / radix) * CP butterfly pairs per tile and threads / CP of them run at once.
// role: 0 = any pass, 1 = first pass only, 2 = later passes only.  Short passes exist twice:
// 16-column tiles for the first pass of a plan (its re-ordered column runs and fused loaders
// measured faster narrow) and 32-column tiles (256-byte rows) for the later passes.
// Kernels are instantiated in rc_fft3_g*.cu, one group per translation unit (id % 4) so they
// compile in parallel.
#define RC_V3_GROUP0(X) \
    X(0, 10, 1, 10, 80, 10, 8, 1) X(4, 10, 1, 16, 128, 4, 8, 0) X(8, 5, 6, 10, 240, 3, 8, 0) X(12, 8, 8, 8, 256, 2, 8, 0) \
    X(16, 8, 10, 10, 320, 2, 8, 0) X(20, 8, 1, 8, 64, 10, 8, 1) X(24, 10, 1, 10, 160, 5, 16, 2) X(32, 8, 1, 8, 128, 8, 16, 2)
#define RC_V3_GROUP1(X) \
    X(1, 8, 1, 16, 128, 4, 8, 0) X(5, 4, 5, 10, 200, 4, 8, 0) X(9, 4, 8, 10, 320, 2, 8, 0) X(13, 6, 10, 10, 160, 2, 8, 0) \
    X(17, 10, 10, 10, 400, 1, 8, 0) X(21, 8, 1, 10, 80, 10, 8, 1) X(33, 8, 1, 10, 160, 5, 16, 2)
#define RC_V3_GROUP2(X) \
    X(2, 5, 5, 5, 200, 4, 8, 0) X(6, 5, 5, 10, 200, 4, 8, 0) X(10, 5, 8, 10, 320, 2, 8, 0) X(14, 5, 5, 25, 200, 2, 8, 0) \
    X(18, 5, 1, 8, 64, 10, 8, 1) X(30, 5, 1, 8, 128, 8, 16, 2) X(34, 5, 1, 10, 160, 5, 32, 2)
#define RC_V3_GROUP3(X) \
    X(3, 10, 1, 15, 128, 4, 8, 0) X(7, 4, 8, 8, 256, 4, 8, 0) X(11, 5, 10, 10, 200, 3, 8, 0) X(15, 8, 8, 10, 256, 2, 8, 0) \
    X(19, 5, 1, 10, 80, 10, 8, 1) X(27, 5, 1, 10, 160, 5, 16, 2)
#define RC_V3_ALL(X) RC_V3_GROUP0(X) RC_V3_GROUP1(X) RC_V3_GROUP2(X) RC_V3_GROUP3(X)
constexpr int kV3Groups = 4;

struct V3Entry { int id, R0, R1, R2, threads, cp, role; int R() const { return R0 * R1 * R2; } };
inline const std::vector<V3Entry>& v3_table() {
    static const std::vector<V3Entry> t = {
#define RC_V3_ROW(id, r0, r1, r2, nt, mb, cp, role) {id, r0, r1, r2, nt, cp, role},
        RC_V3_ALL(RC_V3_ROW)
#undef RC_V3_ROW
    };
    return t;
}
// schedule of length R usable as the first (first = true) or as a later pass of a plan
// 64-column variants (cp = 32: 512-byte rows, one warp per row) are preferred for later passes
inline bool v3_wide64() { return getenv("RC_NO_WIDE64") == nullptr; }
inline const V3Entry* v3_find(int R, bool first) {
    const V3Entry* hit = nullptr;
    for (const V3Entry& e : v3_table()) {
        if (e.R() != R || !(e.role == 0 || e.role == (first ? 1 : 2))) continue;
        if (e.cp == 32) {
            if (v3_wide64()) return &e;
            continue;
        }
        if (!hit) hit = &e;
    }
    return hit;
}
inline bool v3_has(int R) { return v3_find(R, true) && v3_find(R, false); }

// Relative cost of one pass of length R (1.0 = a pass running at the best measured rate;
// B200, 256 M-point transforms, tools/membench.cu + bench per-kernel timings): small tiles keep
// many CTAs per SM and overlap their load / compute / store phases, long ones do not.
inline double fft_pass_cost(int R, bool first) {
    double c;
    if (first) {                      // 16-column tiles; re-ordered column runs
        if (R <= 50) c = 1.5;
        else if (R <= 100) c = 1.25;
        else if (R <= 256) c = 1.05;
        else if (R <= 400) c = 1.4;
        else if (R <= 640) c = 1.6;
        else c = 2.2;
    } else {
        if (R <= 50) c = 1.15;        // 32-column tiles for R <= 100
        else if (R <= 100) c = 1.0;
        else if (R <= 160) c = 1.15;
        else if (R <= 256) c = 1.2;
        else if (R <= 400) c = 1.35;
        else if (R <= 800) c = 1.3;
        else c = 2.0;                 // one CTA per SM
        if (R == 625) c += 0.15;      // radix-25 stage
    }
    return c;
}

// Splits measured best on B200 for the sizes of the BASELINE configurations (bench.py per-kernel
// timings, profiles/); everything else goes through the cost model.
inline bool fft_tuned_split(long long n, std::vector<int>& Rs) {
    struct Row { long long n; int r[4]; };
    static const Row rows[] = {
        {256000000LL, {800, 640, 500, 0}},
        {1000000LL, {200, 50, 100, 0}},
        {500000LL, {200, 50, 50, 0}},
        {250000LL, {500, 500, 0, 0}},
        // round 2 sweep (profiles/r02_s12_split_sweep_small.txt): configs 2 / 4, short blocks, sharded load
        {10000000LL, {160, 250, 250, 0}},
        {16000000LL, {256, 250, 250, 0}},
        {8000000LL, {320, 250, 100, 0}},
        {31250LL, {250, 125, 0, 0}},
        {125000LL, {250, 500, 0, 0}},
        {128000000LL, {200, 800, 800, 0}},
        {32000000LL, {160, 800, 250, 0}},
    };
    for (const Row& row : rows)
        if (row.n == n) {
            Rs.clear();
            for (int i = 0; i < 4 && row.r[i]; i++) Rs.push_back(row.r[i]);
            return true;
        }
    return false;
}

// Split n into 2..4 curated pass lengths of least estimated cost; false when n has no such split.
// RC_FFT_SPLIT="n:r0xr1x..;n:..." forces a split (kernel experiments).
inline bool fft_choose_fast(long long n, std::vector<int>& Rs) {
    std::vector<int> cur;
    int max_r = 1 << 30;
    if (const char* env = getenv("RC_FFT_MAXR")) max_r = atoi(env);      // experiments: cap the pass length
    if (const char* env = getenv("RC_FFT_SPLIT")) {
        std::string e(env);
        size_t pos = 0;
        while (pos < e.size()) {
            size_t end = e.find(';', pos);
            if (end == std::string::npos) end = e.size();
            std::string item = e.substr(pos, end - pos);
            pos = end + 1;
            size_t colon = item.find(':');
            if (colon == std::string::npos || atoll(item.substr(0, colon).c_str()) != n) continue;
            std::vector<int> r;
            long long prod = 1;
            std::string rest = item.substr(colon + 1);
            size_t q = 0;
            while (q < rest.size()) {
                size_t x = rest.find('x', q);
                if (x == std::string::npos) x = rest.size();
                int v = atoi(rest.substr(q, x - q).c_str());
                if (v <= 0 || !v3_has(v)) { r.clear(); break; }
                r.push_back(v);
                prod *= v;
                q = x + 1;
            }
            if (r.size() >= 2 && r.size() <= (size_t)kMaxPasses && prod == n) { Rs = r; return true; }
        }
    }
    if (max_r == (1 << 30) && fft_tuned_split(n, Rs)) return true;
    for (const V3Entry& e : v3_table())
        if (n % e.R() == 0 && e.R() <= max_r && e.role != 2 && v3_has(e.R())) cur.push_back(e.R());
    double best = 1e30;
    std::vector<int> pick;
    auto score = [&](const std::vector<int>& r) {
        double s = 0.0;
        long long Ns = 1;
        for (size_t i = 0; i < r.size(); i++) {
            double c = fft_pass_cost(r[i], i == 0);
            if ((n / r[i]) % 16) c += 0.05;            // input rows not 128-byte aligned
            if (i > 0 && (Ns % 16)) c += 0.05;         // output tiles straddle Ns blocks
            s += c;
            Ns *= r[i];
        }
        return s;
    };
    std::vector<int> r;
    std::function<void(long long, int)> rec = [&](long long rest, int depth) {
        if (rest == 1 && r.size() >= 2) {
            double sc = score(r);
            if (sc < best - 1e-9) { best = sc; pick = r; }
            return;
        }
        if (depth == kMaxPasses) return;
        for (int c : cur) {
            if (rest % c) continue;
            r.push_back(c);
            rec(rest / c, depth + 1);
            r.pop_back();
        }
    };
    rec(n, 0);
    Rs = pick;
    return !pick.empty();
}

inline cudaError_t fft_plan_build(FftPlan& plan, long long n, TableStore& store) {
    if (!fft_size_supported(n)) return cudaErrorInvalidValue;
    std::vector<int> Rs;
    int T = 1;
    if (!fft_choose_passes(n, Rs, T)) return cudaErrorInvalidValue;
    plan.n = n;
    plan.npass = (int)Rs.size();
    long long Ns = 1;
    for (int i = 0; i < plan.npass; i++) {
        cudaError_t err = fft_fill_pass(plan.pass[i], n, Rs[i], T, Ns, store, -1);
        if (err != cudaSuccess) return err;
        Ns *= Rs[i];
    }
    plan.nfast = 0;
    std::vector<int> Fs;
    if (fft_choose_fast(n, Fs)) {
        plan.nfast = (int)Fs.size();
        Ns = 1;
        for (int i = 0; i < plan.nfast; i++) {
            const V3Entry* ent = v3_find(Fs[i], i == 0);
            cudaError_t err = fft_fill_pass(plan.fast[i], n, Fs[i], 2 * ent->cp, Ns, store, ent->id);
            if (err != cudaSuccess) return err;
            Ns *= Fs[i];
        }
        plan.store = &store;
    }
    return cudaSuccess;
}

}  // namespace rc
