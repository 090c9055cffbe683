// rc_engine.cu -- host side of libradiocore_b200: plans, buffers, kernel
// sequencing and the extern "C" boundary declared in include/radiocore_b200.h.
//
// Data layout in HBM (all row-major, one row per channel / batch entry):
//   X      complex64[N]            wideband spectrum of the current block
//   y      complex64[C][B]         channelised IQ (Tuner.run output)
//   Z*     complex64[C][B/2]       half-length FFTs of packed real signals
//   audio  float32  [C][A][nch]    final audio, interleaved L/R for WBFM
// Carried state: de-emphasis zi, double[C][nch][50].
#include <math.h>
#include <stdint.h>

#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/radiocore_b200.h"
#include "rc_exec.cuh"

#ifndef RC_EMULATE
#include <nvtx3/nvToolsExt.h>     // header-only NVTX v3: ranges around the block-level entry points
#endif

namespace rc {

static thread_local std::string g_last_error;

static int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}
static int cuda_fail(cudaError_t e, const char* where) {
    return fail(RC_ERR_CUDA, std::string(where) + ": " + cudaGetErrorString(e));
}
#define RC_API_CUDA(expr, where)                          \
    do {                                                  \
        cudaError_t _e = (expr);                          \
        if (_e != cudaSuccess) return cuda_fail(_e, where); \
    } while (0)

// NVTX range for the lifetime of an API call (visible in Nsight Systems / ncu --nvtx): the
// reference has no profiler hooks; these mark Tuner.load and the channel loop.
struct NvtxRange {
#ifndef RC_EMULATE
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
#else
    explicit NvtxRange(const char*) {}
#endif
};

struct DeviceGuard {
    int prev = -1;
    bool active = false;
    explicit DeviceGuard(int dev) {
#ifndef RC_EMULATE
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) { cudaSetDevice(dev); active = true; }
#else
        (void)dev;
#endif
    }
    ~DeviceGuard() {
#ifndef RC_EMULATE
        if (active) cudaSetDevice(prev);
#endif
    }
};

// Owns device allocations of one handle.
struct Arena {
    std::vector<void*> ptrs;
    size_t bytes = 0;
    ~Arena() { for (void* p : ptrs) dev_free(p); }
    template <typename T> cudaError_t alloc(T** out, size_t count) {
        void* p = nullptr;
        cudaError_t e = dev_malloc(&p, count * sizeof(T));
        if (e != cudaSuccess) return e;
        ptrs.push_back(p);
        bytes += count * sizeof(T);
        *out = (T*)p;
        return cudaSuccess;
    }
    template <typename T> cudaError_t upload(T** out, const std::vector<T>& host) {
        RC_CHECK(alloc(out, host.size()));
        return dev_copy(*out, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice, 0);
    }
};

// ------------------------------------------------------------------ host math
// Restatements of the reference's constructor-time tap design (host, double).

// deemphasis.py:37-49: 51-tap FIR of the one-pole IIR via dlti/dimpulse, cast to
// float32; zi = lfilter_zi(taps, 1) evaluated in float32 (SciPy 1.18.1).
static void deemphasis_design(double tau, long long size, float taps[51], float zi[50]) {
    const double x = exp(-1.0 / ((double)size * tau));
    double p = 1.0;
    taps[0] = 0.0f;
    for (int n = 1; n <= 50; n++) {
        taps[n] = (float)((1.0 - x) * p);
        p = x * p;
    }
    float sum = 0.0f;                       // y_inf = sum(b) in float32, sequential like numpy for n < 8? (pairwise
    for (int n = 0; n <= 50; n++) sum += taps[n];   // summation only starts above 128 elements)
    float acc = 0.0f;                       // flip(cumsum(flip(b - y_inf*a)))[1:], a = [1,0,...]
    for (int i = 50; i >= 1; i--) {
        acc += taps[i];
        zi[i - 1] = acc;
    }
    (void)sum;                              // b[0] - y_inf only lands in the dropped element
}

static double sinc_pi(double x) { return x == 0.0 ? 1.0 : sin(kPi * x) / (kPi * x); }

// bandpass.py:50-54: firwin(num_taps, [lo, hi], pass_zero=False, window=...), float32.
static int firwin_bandpass(int num_taps, double lo, double hi, const std::string& window,
                           std::vector<float>& taps) {
    if (num_taps < 1 || !(lo > 0.0 && lo < hi && hi < 1.0)) return RC_ERR_INVALID;
    std::vector<double> h(num_taps);
    const double alpha = 0.5 * (num_taps - 1);
    for (int i = 0; i < num_taps; i++) {
        const double m = i - alpha;
        double w;
        const double ph = num_taps > 1 ? 2.0 * kPi * i / (num_taps - 1) : 0.0;
        if (window == "hamm" || window == "hamming") w = 0.54 - 0.46 * cos(ph);
        else if (window == "hann" || window == "hanning") w = 0.5 - 0.5 * cos(ph);
        else if (window == "boxcar" || window == "rect") w = 1.0;
        else if (window == "blackman") w = 0.42 - 0.5 * cos(ph) + 0.08 * cos(2.0 * ph);
        else return RC_ERR_INVALID;
        if (num_taps == 1) w = 1.0;
        h[i] = (hi * sinc_pi(hi * m) - lo * sinc_pi(lo * m)) * w;
    }
    const double fc = 0.5 * (lo + hi);
    double s = 0.0;
    for (int i = 0; i < num_taps; i++) s += h[i] * cos(kPi * (i - alpha) * fc);
    taps.resize(num_taps);
    for (int i = 0; i < num_taps; i++) taps[i] = (float)(h[i] / s);
    return RC_OK;
}

// g = b (*) reversed(b): the zero-phase kernel filtfilt applies in the interior.
static std::vector<double> autocorr_taps(const std::vector<float>& b) {
    const int K = (int)b.size() - 1;
    std::vector<double> g(2 * K + 1, 0.0);
    for (int i = 0; i <= K; i++)
        for (int j = 0; j <= K; j++) g[i - j + K] += (double)b[i] * (double)b[j];
    return g;
}

// FiltFiltEw for taps g (2K+1 values); the folded fp32 interior path is enabled for the 81-tap
// pilot filter (K == kFoldK) unless RC_NO_FOLD is set (kernel experiments).
static FiltFiltEw make_filtfilt(const float* x, float* out, const double* d_g, const std::vector<double>& g_host,
                                long long n, int K) {
    FiltFiltEw f{};
    f.x = x; f.out = out; f.g = d_g; f.n = n; f.K = K;
    f.fold.on = 0;
    if (K == kFoldK && (int)g_host.size() == 2 * K + 1 && getenv("RC_NO_FOLD") == nullptr) {
        f.fold.on = 1;
        f.fold.gc = g_host[K];
        f.fold.g[0] = 0.f;
        for (int d = 1; d <= K; d++) f.fold.g[d] = (float)g_host[K + d];
    }
    return f;
}

static std::vector<float2> unit_circle_table(long long n, long long count, int sign) {
    std::vector<float2> t((size_t)count);
    for (long long k = 0; k < count; k++) {
        const long double a = sign * 2.0L * 3.14159265358979323846264338327950288L * (long double)k / (long double)n;
        t[(size_t)k] = make_float2((float)cosl(a), (float)sinl(a));
    }
    return t;
}

// Hann weights of Tuner.run's gathered bins (tuner.py:155-161) for channel size B out of N:
// entry i is W(kc) / N with kc = i (i <= B/2) or i - B, W the fftshifted periodic Hann window
// of length N; *w_neg_half = W(-B/2) / N, the weight of the bin merged into i = B/2.
static std::vector<float> tuner_window_table(long long N, long long B, float* w_neg_half) {
    std::vector<float> t((size_t)B);
    const long double phi = (N % 2 == 0) ? 0.0L : 3.14159265358979323846264338327950288L / (long double)N;
    auto w = [&](long long kc) {
        const long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double)kc / (long double)N + phi;
        return (0.5L + 0.5L * cosl(a)) / (long double)N;
    };
    for (long long i = 0; i < B; i++) t[(size_t)i] = (float)w(i <= B / 2 ? i : i - B);
    *w_neg_half = (float)w(-(B / 2));
    return t;
}

// ------------------------------------------------------- real resampler setup
static cudaError_t make_real_spec(RealResampleSpec& s, long long n_x, long long num, bool hamming,
                                  bool full_table, Arena& arena) {
    s.n_x = n_x; s.num = num; s.h = n_x / 2; s.hp = num / 2;
    s.m = n_x < num ? n_x : num;
    s.m2 = s.m / 2 + 1;
    const double phi = (n_x % 2 == 0) ? 0.0 : kPi / (double)n_x;
    s.a0 = hamming ? 0.54f : 1.0f;
    s.a1c = hamming ? (float)(0.46 * cos(phi)) : 0.0f;
    s.scale = (float)(((double)num / (double)n_x) / (double)s.hp);
    s.nyq = (num == n_x) ? 1.0f : (num < n_x ? 2.0f : 0.5f);
    const long long nr = full_table ? s.h + 1 : (s.m2 < s.h + 1 ? s.m2 : s.h + 1);
    float2* rtw = nullptr; float2* itw = nullptr;
    RC_CHECK(arena.upload(&rtw, unit_circle_table(n_x, nr, -1)));
    RC_CHECK(arena.upload(&itw, unit_circle_table(num, s.hp, +1)));
    s.rtw = rtw; s.itw = itw;
    return cudaSuccess;
}

// ---------------------------------------------------------------- demod bank
// `batch` channels of identical (mode, B, A, tau): everything after Tuner.run.
struct DemodBank {
    int mode = 0, batch = 0, nch = 1;
    long long B = 0, A = 0, h = 0, hp = 0;
    double tau = 75e-6;
    FftPlan planBh, planAh;
    RealResampleSpec specBA{}, specBB{}, specH{};
    float taps[51];
    float zi0[50];
    float* d_taps = nullptr;
    double* d_zi = nullptr; double* d_zi_next = nullptr;
    double* d_g = nullptr; int gK = 0;
    std::vector<double> g_host;
    float2 *Z1 = nullptr, *Z2 = nullptr, *ZpB = nullptr, *ZpA = nullptr, *w0 = nullptr, *w1 = nullptr;
    float *mpx = nullptr, *pilot = nullptr, *lmr = nullptr, *audio_tmp = nullptr;

    int init(int mode_, long long B_, long long A_, double tau_, int batch_, TableStore& store, Arena& arena) {
        mode = mode_; B = B_; A = A_; tau = tau_; batch = batch_;
        nch = mode == RC_MODE_WBFM ? 2 : 1;
        if (B < 2 || A < 2 || batch < 1) return fail(RC_ERR_INVALID, "demod: sizes must be >= 2 and batch >= 1");
        if ((B & 1) || (A & 1)) return fail(RC_ERR_UNSUPPORTED, "demod: input_size and output_size must be even");
        h = B / 2; hp = A / 2;
        if (!fft_size_supported(h) || !fft_size_supported(hp))
            return fail(RC_ERR_UNSUPPORTED, "demod: sizes must factor into 2^a 3^b 5^c");
        if (mode == RC_MODE_WBFM && B <= 3 * 41)
            return fail(RC_ERR_INVALID, "wbfm: input_size must exceed the filtfilt pad length (123)");
        RC_API_CUDA(fft_plan_build(planBh, h, store), "plan B/2");
        RC_API_CUDA(fft_plan_build(planAh, hp, store), "plan A/2");
        const bool wb = mode == RC_MODE_WBFM;
        RC_API_CUDA(make_real_spec(specBA, B, A, true, false, arena), "spec B->A");
        RC_API_CUDA(arena.alloc(&Z1, (size_t)batch * h), "alloc Z1");
        RC_API_CUDA(arena.alloc(&ZpA, (size_t)batch * nch * hp), "alloc ZpA");
        size_t wlen = (size_t)batch * (h > nch * hp ? h : nch * hp);
        RC_API_CUDA(arena.alloc(&w0, wlen), "alloc w0");
        RC_API_CUDA(arena.alloc(&w1, wlen), "alloc w1");
        RC_API_CUDA(arena.alloc(&audio_tmp, (size_t)batch * nch * A), "alloc audio_tmp");
        if (mode != RC_MODE_FM) {
            deemphasis_design(tau, A, taps, zi0);
            std::vector<float> t(taps, taps + 51);
            RC_API_CUDA(arena.upload(&d_taps, t), "taps");
            RC_API_CUDA(arena.alloc(&d_zi, (size_t)batch * nch * 50), "zi");
            RC_API_CUDA(arena.alloc(&d_zi_next, (size_t)batch * nch * 50), "zi_next");
            int rc = reset_state();
            if (rc) return rc;
        }
        if (wb) {
            RC_API_CUDA(make_real_spec(specBB, B, B, true, true, arena), "spec B->B");
            RC_API_CUDA(make_real_spec(specH, B, B, false, true, arena), "spec hilbert");
            specH.scale = (float)(1.0 / (double)h);
            std::vector<float> b;
            const double nyq = 0.5 * (double)B;
            int rc = firwin_bandpass(41, (19e3 - 50) / nyq, (19e3 + 50) / nyq, "hamm", b);
            if (rc) return fail(rc, "wbfm: 19 kHz pilot filter needs input_size > 38100");
            std::vector<double> g = autocorr_taps(b);
            gK = 40;
            g_host = g;
            RC_API_CUDA(arena.upload(&d_g, g), "pilot taps");
            RC_API_CUDA(arena.alloc(&Z2, (size_t)batch * h), "alloc Z2");
            RC_API_CUDA(arena.alloc(&ZpB, (size_t)batch * h), "alloc ZpB");
            RC_API_CUDA(arena.alloc(&mpx, (size_t)batch * B), "alloc mpx");
            RC_API_CUDA(arena.alloc(&pilot, (size_t)batch * B), "alloc pilot");
            RC_API_CUDA(arena.alloc(&lmr, (size_t)batch * B), "alloc lmr");
        }
        return RC_OK;
    }

    int reset_state() {
        if (mode == RC_MODE_FM) return RC_OK;
        std::vector<double> z((size_t)batch * nch * 50);
        for (size_t i = 0; i < z.size(); i++) z[i] = (double)zi0[i % 50];
        RC_API_CUDA(dev_copy(d_zi, z.data(), z.size() * sizeof(double), cudaMemcpyHostToDevice, 0), "zi reset");
        RC_API_CUDA(dev_sync(0), "zi reset sync");
        return RC_OK;
    }

    // y: [batch][B] complex64;  out: [batch][A][nch] float32
    int run(const float2* y, float* out, cudaStream_t st) {
        return run_from(LoadDiscriminatorPacked{y, B}, 8.0 * B * batch, out, st);
    }
    // ang: [batch][B] angle(y)/pi as stored by the engine's channel IFFT (StoreAngle)
    int run_angle(const float* ang, float* out, cudaStream_t st) {
        return run_from(LoadAnglePacked{ang, B}, 4.0 * B * batch, out, st);
    }

    template <class DiscLoad>
    int run_from(const DiscLoad& disc, double disc_bytes, float* out, cudaStream_t st) {
        if (mode != RC_MODE_WBFM) {
            // Decimate keeps bins k < m2 of the discriminator's spectrum; through the packed-real
            // algebra they depend on Z1[k] and Z1[h-k] only: skip the stores in between.
            long long lo = (specBA.m2 + 2) / 2 * 2, hi = (h - specBA.m2 - 1) / 2 * 2;
            if (hi < lo) hi = lo;
            RC_API_CUDA((fft_exec<-1>(planBh, batch, disc, StoreC64Win{Z1, h, lo, hi}, w0, w1, st,
                                  "demod.rfft_disc", disc_bytes, 8.0 * (double)(h - (hi - lo)) * batch)), "fft discriminator");
        } else {
            RC_API_CUDA((fft_exec<-1>(planBh, batch, disc, StoreC64{Z1, h, 1.0f}, w0, w1, st,
                                  "demod.rfft_disc", disc_bytes, 0.0)), "fft discriminator");
        }
        if (mode != RC_MODE_WBFM) {
            RC_API_CUDA(launch_ew(hp, batch, SpecResampleEw{specBA, Z1, ZpA}, st, "demod.spec_resample",
                                  24.0 * hp * batch), "spec B->A");
            float* dst = mode == RC_MODE_FM ? out : audio_tmp;
            RC_API_CUDA((fft_exec<+1>(planAh, batch, LoadC64{ZpA, hp}, StoreC64{(float2*)dst, hp, 1.0f}, w0, w1, st,
                                  "demod.irfft_audio")), "ifft audio");
            if (mode == RC_MODE_FM) return RC_OK;
        } else {
            // mpx = FM(B, B): same-size resample = folded Hamming taper (wbfm.py:42-43,77)
            RC_API_CUDA(launch_ew(h / 2 + 1, batch, SpecTaperPairEw{specBB, Z1, ZpB}, st, "wbfm.spec_taper", 16.0 * h * batch),
                        "spec B->B");
            RC_API_CUDA((fft_exec<+1>(planBh, batch, LoadC64{ZpB, h}, StoreC64{(float2*)mpx, h, 1.0f}, w0, w1, st,
                                  "wbfm.irfft_mpx")), "ifft mpx");
            // pilot = Bandpass(19 kHz +- 50, 41 taps).run(mpx)   (wbfm.py:45-46,80)
            RC_API_CUDA(launch_filtfilt(make_filtfilt(mpx, pilot, d_g, g_host, B, gK), batch, st, "wbfm.pilot_filtfilt", g_host.data()),
                        "pilot filtfilt");
            // PLL.step (hilbert) + image(2) * mpx * 1.0175      (wbfm.py:80-83)
            RC_API_CUDA((fft_exec<-1>(planBh, batch, LoadC64{(const float2*)pilot, h}, StoreC64{Z2, h, 1.0f}, w0, w1, st,
                                  "wbfm.rfft_pilot")), "fft pilot");
            RC_API_CUDA(launch_ew(h / 2 + 1, batch, SpecHilbertPairEw{specH, Z2, ZpB}, st, "wbfm.spec_hilbert", 16.0 * h * batch),
                        "spec hilbert");
            RC_API_CUDA((fft_exec<+1>(planBh, batch, LoadC64{ZpB, h}, StoreLmrPacked{pilot, mpx, lmr, B}, w0, w1, st,
                                  "wbfm.irfft_hilbert_lmr", 0.0, 12.0 * B * batch)), "ifft hilbert");
            // L, R = Decimate(mpx +- lmr)                        (wbfm.py:86-87)
            RC_API_CUDA((fft_exec<-1>(planBh, batch, LoadC64{(const float2*)lmr, h}, StoreC64{Z2, h, 1.0f}, w0, w1, st,
                                  "wbfm.rfft_lmr")), "fft lmr");
            RC_API_CUDA(launch_ew(hp, batch, SpecStereoEw{specBA, Z1, Z2, ZpA}, st, "wbfm.spec_stereo",
                                  48.0 * hp * batch), "spec stereo");
            RC_API_CUDA((fft_exec<+1>(planAh, batch * 2, LoadC64{ZpA, hp}, StoreC64{(float2*)audio_tmp, hp, 1.0f}, w0, w1, st,
                                  "wbfm.irfft_audio")), "ifft audio LR");
        }
        EpilogueParams p;
        p.in = audio_tmp; p.out = out; p.zi = d_zi; p.zi_next = d_zi_next; p.taps = d_taps;
        p.A = A; p.nch = nch; p.ntaps = 51; p.deemph = 1; p.dc_clip = 1;
        p.dc = ZpA; p.dc_stride = hp; p.dc_scale = (double)hp;      // bin 0 of every audio channel's packed spectrum
        RC_API_CUDA(launch_epilogue(p, batch, st, taps), "epilogue");
        // carried state: copied back rather than swapping the two pointers, so that a captured
        // CUDA graph of the block (fixed kernel arguments) carries it correctly when replayed
        RC_API_CUDA(dev_copy(d_zi, d_zi_next, (size_t)batch * nch * 50 * sizeof(double), cudaMemcpyDeviceToDevice, st), "zi carry");
        return RC_OK;
    }
};

}  // namespace rc

using namespace rc;

// ------------------------------------------------------------------- engine
struct rc_engine {
    int device = 0;
    long long N = 0;
    bool committed = false, loaded = false;
    TableStore store{kOnDevice};
    Arena arena;
    FftPlan planN;
    float2 *X = nullptr, *wN0 = nullptr, *wN1 = nullptr;
    struct Chan { long long roll, B, A; int mode; double tau; int bank, slot; };
    std::vector<Chan> chans;
    std::map<long long, FftPlan> planB;     // inverse channel FFT plans by bandwidth
    std::map<long long, const float*> wtab; // Hann weights of the gathered bins by bandwidth (LoadTunerGather)
    std::map<long long, float> wneg;
    struct Bank {
        DemodBank demod;
        const FftPlan* planB = nullptr;
        std::vector<int> members;
        long long* d_roll = nullptr;
        float* ang = nullptr;                   // angle(y)/pi of every channel sample (StoreAngle)
        float2 *w0 = nullptr, *w1 = nullptr;
        long long audio_offset = 0;
    };
    std::vector<std::unique_ptr<Bank>> banks;
    long long* d_roll_all = nullptr;
    float2* y_single = nullptr; size_t y_single_len = 0;
    long long audio_total = 0;
    // sub-band mode (rc_engine_set_subband): X is handed in per block and holds bins
    // [x_lo, x_lo + x_len) of the N-bin spectrum; rolls are shifted by x_lo
    bool subband = false;
    long long x_lo = 0, x_len = 0;
    long long shifted_roll(long long roll) const {
        long long r = ((roll % N) + N) % N;
        if (subband) r = (r + x_lo) % N;
        return r;
    }
};

extern "C" {

const char* rc_last_error(void) { return g_last_error.c_str(); }
int rc_version(void) { return 100; }
int rc_size_supported(int64_t n) { return fft_size_supported(n) ? 1 : 0; }

int rc_engine_create(int device, int64_t n_input, rc_engine** out) {
    if (!out || n_input < 2) return fail(RC_ERR_INVALID, "engine: n_input must be >= 2");
    if (!fft_size_supported(n_input)) return fail(RC_ERR_UNSUPPORTED, "engine: n_input must factor into 2^a 3^b 5^c");
    rc_engine* e = new rc_engine();
    e->device = device;
    e->N = n_input;
    *out = e;
    return RC_OK;
}

int rc_engine_destroy(rc_engine* e) {
    if (!e) return RC_OK;
    DeviceGuard g(e->device);
    delete e;
    return RC_OK;
}

int rc_engine_add_channel(rc_engine* e, int64_t roll, int64_t B, int64_t A, int mode, double tau, int* out_index) {
    if (!e) return fail(RC_ERR_INVALID, "engine: null handle");
    if (e->committed) return fail(RC_ERR_STATE, "engine: add_channel after commit");
    if (mode < 0 || mode > RC_MODE_NONE) return fail(RC_ERR_INVALID, "engine: unknown mode");
    if (mode == RC_MODE_NONE) A = 2;
    // B == n_input (a single channel as wide as the block): resample with num == Nx keeps every bin
    // and merges nothing (SciPy _signaltools.py:3861-3875), i.e. window multiply + inverse FFT
    if (B < 2 || B > e->N) return fail(RC_ERR_INVALID, "engine: channel bandwidth must be in [2, n_input]");
    if (roll <= -e->N || roll >= e->N) return fail(RC_ERR_INVALID, "engine: |roll| must be < n_input");
    if ((B & 1) || (A & 1) || !fft_size_supported(B) || !fft_size_supported(A))
        return fail(RC_ERR_UNSUPPORTED, "engine: channel sizes must be even and factor into 2^a 3^b 5^c");
    e->chans.push_back({roll, B, A, mode, tau, -1, -1});
    if (out_index) *out_index = (int)e->chans.size() - 1;
    return RC_OK;
}

int rc_engine_commit(rc_engine* e) {
    if (!e) return fail(RC_ERR_INVALID, "engine: null handle");
    if (e->committed) return fail(RC_ERR_STATE, "engine: already committed");
    if (e->chans.empty()) return fail(RC_ERR_STATE, "engine: no channels");
    DeviceGuard g(e->device);
    if (e->subband) {
        // every bin a channel gathers -- (i - roll - x_lo) mod N for i in [-B/2, B/2] -- must lie
        // inside the sub-band the engine will be handed
        for (auto& c : e->chans) {
            const long long first = (((-(c.B / 2) - c.roll - e->x_lo) % e->N) + e->N) % e->N;
            if (first + c.B + 1 > e->x_len) return fail(RC_ERR_INVALID, "engine: a channel reads bins outside the sub-band");
        }
    } else {
        RC_API_CUDA(fft_plan_build(e->planN, e->N, e->store), "plan N");
        RC_API_CUDA(e->arena.alloc(&e->X, (size_t)e->N), "alloc X");
        if (e->planN.max_passes() >= 2) RC_API_CUDA(e->arena.alloc(&e->wN0, (size_t)e->N), "alloc wN0");
        if (e->planN.max_passes() >= 3) RC_API_CUDA(e->arena.alloc(&e->wN1, (size_t)e->N), "alloc wN1");
    }
    // group channels with identical demodulator configuration
    size_t maxB = 0;
    for (size_t i = 0; i < e->chans.size(); i++) {
        auto& c = e->chans[i];
        if (!e->planB.count(c.B)) {
            RC_API_CUDA(fft_plan_build(e->planB[c.B], c.B, e->store), "plan B");
            float wn = 0.f;
            float* d = nullptr;
            RC_API_CUDA(e->arena.upload(&d, tuner_window_table(e->N, c.B, &wn)), "window table");
            e->wtab[c.B] = d;
            e->wneg[c.B] = c.B == e->N ? 0.f : wn;      // num == Nx: bins +-num/2 coincide, nothing is merged
        }
        if ((size_t)c.B > maxB) maxB = (size_t)c.B;
        if (c.mode == RC_MODE_NONE) continue;
        int found = -1;
        for (size_t b = 0; b < e->banks.size(); b++) {
            auto& c0 = e->chans[e->banks[b]->members[0]];
            if (c0.B == c.B && c0.A == c.A && c0.mode == c.mode && c0.tau == c.tau) { found = (int)b; break; }
        }
        if (found < 0) { e->banks.emplace_back(new rc_engine::Bank()); found = (int)e->banks.size() - 1; }
        c.bank = found;
        c.slot = (int)e->banks[found]->members.size();
        e->banks[found]->members.push_back((int)i);
    }
    long long off = 0;
    for (auto& bp : e->banks) {
        auto& bk = *bp;
        auto& c0 = e->chans[bk.members[0]];
        const int batch = (int)bk.members.size();
        int rc = bk.demod.init(c0.mode, c0.B, c0.A, c0.tau, batch, e->store, e->arena);
        if (rc) return rc;
        bk.planB = &e->planB[c0.B];
        std::vector<long long> rolls;
        for (int m : bk.members) rolls.push_back(e->shifted_roll(e->chans[m].roll));
        RC_API_CUDA(e->arena.upload(&bk.d_roll, rolls), "rolls");
        RC_API_CUDA(e->arena.alloc(&bk.ang, (size_t)batch * c0.B), "alloc angle");
        if (bk.planB->max_passes() >= 2) RC_API_CUDA(e->arena.alloc(&bk.w0, (size_t)batch * c0.B), "alloc yw0");
        if (bk.planB->max_passes() >= 3) RC_API_CUDA(e->arena.alloc(&bk.w1, (size_t)batch * c0.B), "alloc yw1");
        bk.audio_offset = off;
        off += (long long)batch * c0.A * bk.demod.nch;
    }
    e->audio_total = off;
    std::vector<long long> all;
    for (auto& c : e->chans) all.push_back(e->shifted_roll(c.roll));
    RC_API_CUDA(e->arena.upload(&e->d_roll_all, all), "rolls all");
    e->y_single_len = maxB;
    RC_API_CUDA(e->arena.alloc(&e->y_single, maxB * 2), "alloc y_single");
    RC_API_CUDA(dev_sync(0), "commit sync");
    e->committed = true;
    return RC_OK;
}

int rc_engine_audio_floats(rc_engine* e, int64_t* total) {
    if (!e || !e->committed) return fail(RC_ERR_STATE, "engine: not committed");
    *total = e->audio_total;
    return RC_OK;
}

int rc_engine_channel_layout(rc_engine* e, int index, int64_t* offset, int64_t* audio_size, int* nch) {
    if (!e || !e->committed) return fail(RC_ERR_STATE, "engine: not committed");
    if (index < 0 || index >= (int)e->chans.size()) return fail(RC_ERR_INVALID, "engine: channel index out of range");
    auto& c = e->chans[index];
    if (c.bank < 0) {                       // IQ-only channel: no audio
        if (offset) *offset = 0;
        if (audio_size) *audio_size = 0;
        if (nch) *nch = 0;
        return RC_OK;
    }
    auto& bk = *e->banks[c.bank];
    if (offset) *offset = bk.audio_offset + (long long)c.slot * c.A * bk.demod.nch;
    if (audio_size) *audio_size = c.A;
    if (nch) *nch = bk.demod.nch;
    return RC_OK;
}

int rc_engine_workspace_bytes(rc_engine* e, int64_t* bytes) {
    if (!e) return fail(RC_ERR_INVALID, "engine: null handle");
    *bytes = (int64_t)e->arena.bytes;
    return RC_OK;
}

// Tuner.load (tuner.py:137-138): X = fft(iq), complex64 in / complex64 out.
int rc_engine_load(rc_engine* e, const void* iq_dev, void* stream) {
    if (!e || !e->committed) return fail(RC_ERR_STATE, "engine: load before commit");
    if (e->subband) return fail(RC_ERR_STATE, "engine: sub-band mode takes rc_engine_load_subband");
    if (!iq_dev) return fail(RC_ERR_INVALID, "engine: null input");
    DeviceGuard g(e->device);
    NvtxRange range("radiocore.Tuner.load");
    cudaStream_t st = (cudaStream_t)stream;
    RC_API_CUDA((fft_exec<-1>(e->planN, 1, LoadC64{(const float2*)iq_dev, e->N}, StoreC64{e->X, e->N, 1.0f},
                              e->wN0, e->wN1, st, "tuner.load_fft")), "tuner load fft");
    e->loaded = true;
    return RC_OK;
}

int rc_engine_set_subband(rc_engine* e, int64_t x_lo, int64_t x_len) {
    if (!e) return fail(RC_ERR_INVALID, "engine: null handle");
    if (e->committed) return fail(RC_ERR_STATE, "engine: set_subband after commit");
    if (x_len < 2 || x_len > e->N || x_lo < 0 || x_lo >= e->N) return fail(RC_ERR_INVALID, "engine: sub-band outside the spectrum");
    e->subband = true;
    e->x_lo = x_lo;
    e->x_len = x_len;
    return RC_OK;
}

// Tuner.load for a sub-band produced elsewhere (sharded load): adopt the pointer, no copy.
int rc_engine_load_subband(rc_engine* e, const void* spectrum_dev) {
    if (!e || !e->committed || !e->subband) return fail(RC_ERR_STATE, "engine: load_subband needs a committed sub-band engine");
    if (!spectrum_dev) return fail(RC_ERR_INVALID, "engine: null spectrum");
    e->X = (float2*)spectrum_dev;
    e->loaded = true;
    return RC_OK;
}

static LoadTunerGather tuner_gather(const rc_engine* e, const long long* d_roll, long long B) {
    LoadTunerGather ld;
    ld.X = e->X; ld.roll = d_roll; ld.wtab = e->wtab.at(B);
    ld.n_x = e->N; ld.num = B; ld.half = B / 2;
    ld.w_neg_half = e->wneg.at(B);
    return ld;
}

// For every channel: Tuner.run (tuner.py:151-161) then demodulator.run.
int rc_engine_run(rc_engine* e, float* audio_dev, void* stream) {
    if (!e || !e->committed || !e->loaded) return fail(RC_ERR_STATE, "engine: run before load");
    if (!audio_dev && e->audio_total > 0) return fail(RC_ERR_INVALID, "engine: null output");
    DeviceGuard g(e->device);
    NvtxRange range("radiocore.Tuner.run+demodulators");
    cudaStream_t st = (cudaStream_t)stream;
    for (auto& bp : e->banks) {
        auto& bk = *bp;
        const long long B = bk.demod.B;
        const int batch = bk.demod.batch;
        RC_API_CUDA((fft_exec<+1>(*bk.planB, batch, tuner_gather(e, bk.d_roll, B), StoreAngle{bk.ang, B},
                                  bk.w0, bk.w1, st, "tuner.channel_ifft", 8.0 * B * batch, 4.0 * B * batch)),   // DRAM bytes: the Hann table stays in L2
                    "tuner channel ifft");
        int rc = bk.demod.run_angle(bk.ang, audio_dev + bk.audio_offset, st);
        if (rc) return rc;
    }
    return RC_OK;
}

int rc_engine_channel_iq(rc_engine* e, int index, void* out, void* stream) {
    if (!e || !e->committed || !e->loaded) return fail(RC_ERR_STATE, "engine: channel_iq before load");
    if (index < 0 || index >= (int)e->chans.size()) return fail(RC_ERR_INVALID, "engine: channel index out of range");
    DeviceGuard g(e->device);
    cudaStream_t st = (cudaStream_t)stream;
    auto& c = e->chans[index];
    RC_API_CUDA((fft_exec<+1>(e->planB[c.B], 1, tuner_gather(e, e->d_roll_all + index, c.B),
                              StoreC64{(float2*)out, c.B, 1.0f}, e->y_single, e->y_single + e->y_single_len, st)),
                "tuner single ifft");
    return RC_OK;
}

int rc_engine_spectrum(rc_engine* e, void* out, void* stream) {
    if (!e || !e->committed || !e->loaded) return fail(RC_ERR_STATE, "engine: spectrum before load");
    DeviceGuard g(e->device);
    RC_API_CUDA(dev_copy(out, e->X, (size_t)(e->subband ? e->x_len : e->N) * sizeof(float2), cudaMemcpyDeviceToDevice,
                         (cudaStream_t)stream), "spectrum copy");
    return RC_OK;
}

int rc_engine_reset_state(rc_engine* e) {
    if (!e || !e->committed) return fail(RC_ERR_STATE, "engine: not committed");
    DeviceGuard g(e->device);
    for (auto& bp : e->banks) { int rc = bp->demod.reset_state(); if (rc) return rc; }
    return RC_OK;
}

// ---------------------------------------------------------- standalone demod
struct rc_demod {
    int device = 0;
    TableStore store{kOnDevice};
    Arena arena;
    DemodBank bank;
};

int rc_demod_create(int device, int mode, int64_t B, int64_t A, double tau, int batch, rc_demod** out) {
    if (!out) return fail(RC_ERR_INVALID, "demod: null out");
    if (mode < 0 || mode > 2) return fail(RC_ERR_INVALID, "demod: unknown mode");
    DeviceGuard g(device);
    std::unique_ptr<rc_demod> d(new rc_demod());
    d->device = device;
    int rc = d->bank.init(mode, B, A, tau, batch, d->store, d->arena);
    if (rc) return rc;
    RC_API_CUDA(dev_sync(0), "demod create sync");
    *out = d.release();
    return RC_OK;
}
int rc_demod_destroy(rc_demod* d) {
    if (!d) return RC_OK;
    DeviceGuard g(d->device);
    delete d;
    return RC_OK;
}
int rc_demod_run(rc_demod* d, const void* iq, float* audio, void* stream) {
    if (!d || !iq || !audio) return fail(RC_ERR_INVALID, "demod: null argument");
    DeviceGuard g(d->device);
    return d->bank.run((const float2*)iq, audio, (cudaStream_t)stream);
}
int rc_demod_reset_state(rc_demod* d) {
    if (!d) return fail(RC_ERR_INVALID, "demod: null handle");
    DeviceGuard g(d->device);
    return d->bank.reset_state();
}

// ------------------------------------------------------------------ decimate
struct rc_decimate {
    int device = 0;
    long long n_in = 0, n_out = 0;
    TableStore store{kOnDevice};
    Arena arena;
    bool real_ready = false, cplx_ready = false;
    FftPlan planIh, planOh, planI, planO;
    RealResampleSpec spec{};
    float2 *Z = nullptr, *Zp = nullptr, *w0 = nullptr, *w1 = nullptr, *Xc = nullptr, *cw0 = nullptr, *cw1 = nullptr;
    bool odd_ready = false;
    float2 *Xr = nullptr, *ow0 = nullptr, *ow1 = nullptr;
};

int rc_decimate_create(int device, int64_t n_in, int64_t n_out, rc_decimate** out) {
    if (!out || n_in < 1 || n_out < 1) return fail(RC_ERR_INVALID, "decimate: sizes must be positive");
    if (!fft_size_supported(n_in) || !fft_size_supported(n_out))
        return fail(RC_ERR_UNSUPPORTED, "decimate: sizes must factor into 2^a 3^b 5^c");
    rc_decimate* d = new rc_decimate();
    d->device = device; d->n_in = n_in; d->n_out = n_out;
    *out = d;
    return RC_OK;
}
int rc_decimate_destroy(rc_decimate* d) {
    if (!d) return RC_OK;
    DeviceGuard g(d->device);
    delete d;
    return RC_OK;
}
int rc_decimate_run_real(rc_decimate* d, const float* in, float* outp, void* stream) {
    if (!d || !in || !outp) return fail(RC_ERR_INVALID, "decimate: null argument");
    DeviceGuard g(d->device);
    cudaStream_t st = (cudaStream_t)stream;
    if ((d->n_in & 1) || (d->n_out & 1)) {
        // odd length: no half-length packing; full-length transforms
        if (!d->odd_ready) {
            if (!d->cplx_ready) {
                RC_API_CUDA(fft_plan_build(d->planI, d->n_in, d->store), "plan in");
                RC_API_CUDA(fft_plan_build(d->planO, d->n_out, d->store), "plan out");
            }
            RC_API_CUDA(d->arena.alloc(&d->Xr, (size_t)d->n_in), "alloc");
            size_t w = (size_t)(d->n_in > d->n_out ? d->n_in : d->n_out);
            RC_API_CUDA(d->arena.alloc(&d->ow0, w), "alloc");
            RC_API_CUDA(d->arena.alloc(&d->ow1, w), "alloc");
            RC_API_CUDA(dev_sync(0), "sync");
            d->odd_ready = true;
        }
        RC_API_CUDA((fft_exec<-1>(d->planI, 1, LoadRealAsComplex{in, d->n_in}, StoreC64{d->Xr, d->n_in, 1.0f},
                                  d->ow0, d->ow1, st)), "fft real");
        LoadHermitianResample ld;
        ld.X = d->Xr; ld.n_x = d->n_in; ld.num = d->n_out;
        ld.m = d->n_in < d->n_out ? d->n_in : d->n_out; ld.m2 = ld.m / 2 + 1;
        const double phi = (d->n_in % 2 == 0) ? 0.0 : kPi / (double)d->n_in;
        ld.a0 = 0.54f; ld.a1c = (float)(0.46 * cos(phi));
        ld.two_over_n = 2.0 / (double)d->n_in;
        ld.scale = (float)(1.0 / (double)d->n_in);
        ld.nyq = (d->n_out == d->n_in) ? 1.0f : (d->n_out < d->n_in ? 2.0f : 0.5f);
        RC_API_CUDA((fft_exec<+1>(d->planO, 1, ld, StoreRealPart{outp, d->n_out}, d->ow0, d->ow1, st)), "ifft real");
        return RC_OK;
    }
    const long long h = d->n_in / 2, hp = d->n_out / 2;
    if (!d->real_ready) {
        RC_API_CUDA(fft_plan_build(d->planIh, h, d->store), "plan in/2");
        RC_API_CUDA(fft_plan_build(d->planOh, hp, d->store), "plan out/2");
        RC_API_CUDA(make_real_spec(d->spec, d->n_in, d->n_out, true, false, d->arena), "spec");
        RC_API_CUDA(d->arena.alloc(&d->Z, (size_t)h), "alloc");
        RC_API_CUDA(d->arena.alloc(&d->Zp, (size_t)hp), "alloc");
        size_t w = (size_t)(h > hp ? h : hp);
        RC_API_CUDA(d->arena.alloc(&d->w0, w), "alloc");
        RC_API_CUDA(d->arena.alloc(&d->w1, w), "alloc");
        RC_API_CUDA(dev_sync(0), "sync");
        d->real_ready = true;
    }
    RC_API_CUDA((fft_exec<-1>(d->planIh, 1, LoadC64{(const float2*)in, h}, StoreC64{d->Z, h, 1.0f}, d->w0, d->w1, st)), "rfft");
    RC_API_CUDA(launch_ew(hp, 1, SpecResampleEw{d->spec, d->Z, d->Zp}, st), "spec");
    RC_API_CUDA((fft_exec<+1>(d->planOh, 1, LoadC64{d->Zp, hp}, StoreC64{(float2*)outp, hp, 1.0f}, d->w0, d->w1, st)), "irfft");
    return RC_OK;
}
int rc_decimate_run_complex(rc_decimate* d, const void* in, void* outp, void* stream) {
    if (!d || !in || !outp) return fail(RC_ERR_INVALID, "decimate: null argument");
    DeviceGuard g(d->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (!d->cplx_ready) {
        if (!d->odd_ready) {
            RC_API_CUDA(fft_plan_build(d->planI, d->n_in, d->store), "plan in");
            RC_API_CUDA(fft_plan_build(d->planO, d->n_out, d->store), "plan out");
        }
        RC_API_CUDA(d->arena.alloc(&d->Xc, (size_t)d->n_in), "alloc");
        size_t w = (size_t)(d->n_in > d->n_out ? d->n_in : d->n_out);
        RC_API_CUDA(d->arena.alloc(&d->cw0, w), "alloc");
        RC_API_CUDA(d->arena.alloc(&d->cw1, w), "alloc");
        RC_API_CUDA(dev_sync(0), "sync");
        d->cplx_ready = true;
    }
    RC_API_CUDA((fft_exec<-1>(d->planI, 1, LoadC64{(const float2*)in, d->n_in}, StoreC64{d->Xc, d->n_in, 1.0f}, d->cw0, d->cw1, st)), "fft");
    LoadResampleGather ld;
    ld.X = d->Xc; ld.x_batch_stride = 0; ld.roll = nullptr;
    ld.n_x = d->n_in; ld.num = d->n_out; ld.m = d->n_in < d->n_out ? d->n_in : d->n_out; ld.m2 = ld.m / 2 + 1;
    ld.win = make_window(false, d->n_in);
    ld.scale = (float)(1.0 / (double)d->n_in);   // resample scale num/n_x times the inverse FFT's 1/num
    RC_API_CUDA((fft_exec<+1>(d->planO, 1, ld, StoreC64{(float2*)outp, d->n_out, 1.0f}, d->cw0, d->cw1, st)), "ifft");
    return RC_OK;
}

// ---------------------------------------------------------------- deemphasis
struct rc_deemph {
    int device = 0;
    long long size = 0;
    Arena arena;
    float taps[51], zi0[50];
    float* d_taps = nullptr;
    double *d_zi = nullptr, *d_zi_next = nullptr;
};
int rc_deemph_taps(double tau, int64_t size, float* taps51, float* zi50) {
    if (!taps51 || !zi50 || size < 1 || !(tau > 0)) return fail(RC_ERR_INVALID, "deemph: bad argument");
    deemphasis_design(tau, size, taps51, zi50);
    return RC_OK;
}
int rc_deemph_reset_state(rc_deemph* d) {
    if (!d) return fail(RC_ERR_INVALID, "deemph: null handle");
    DeviceGuard g(d->device);
    std::vector<double> z(50);
    for (int i = 0; i < 50; i++) z[i] = d->zi0[i];
    RC_API_CUDA(dev_copy(d->d_zi, z.data(), 50 * sizeof(double), cudaMemcpyHostToDevice, 0), "zi");
    RC_API_CUDA(dev_sync(0), "sync");
    return RC_OK;
}
int rc_deemph_create(int device, int64_t size, double tau, rc_deemph** out) {
    if (!out || size < 1 || !(tau > 0)) return fail(RC_ERR_INVALID, "deemph: bad argument");
    DeviceGuard g(device);
    std::unique_ptr<rc_deemph> d(new rc_deemph());
    d->device = device; d->size = size;
    deemphasis_design(tau, size, d->taps, d->zi0);
    std::vector<float> t(d->taps, d->taps + 51);
    RC_API_CUDA(d->arena.upload(&d->d_taps, t), "taps");
    RC_API_CUDA(d->arena.alloc(&d->d_zi, 50), "zi");
    RC_API_CUDA(d->arena.alloc(&d->d_zi_next, 50), "zi");
    int rc = rc_deemph_reset_state(d.get());
    if (rc) return rc;
    *out = d.release();
    return RC_OK;
}
int rc_deemph_destroy(rc_deemph* d) {
    if (!d) return RC_OK;
    DeviceGuard g(d->device);
    delete d;
    return RC_OK;
}
int rc_deemph_run(rc_deemph* d, const float* in, float* outp, void* stream) {
    if (!d || !in || !outp) return fail(RC_ERR_INVALID, "deemph: null argument");
    DeviceGuard g(d->device);
    EpilogueParams p;
    p.in = in; p.out = outp; p.zi = d->d_zi; p.zi_next = d->d_zi_next; p.taps = d->d_taps;
    p.A = d->size; p.nch = 1; p.ntaps = 51; p.deemph = 1; p.dc_clip = 0;
    p.dc = nullptr; p.dc_stride = 0; p.dc_scale = 0.0;
    RC_API_CUDA(launch_epilogue(p, 1, (cudaStream_t)stream, d->taps), "deemph");
    RC_API_CUDA(dev_copy(d->d_zi, d->d_zi_next, 50 * sizeof(double), cudaMemcpyDeviceToDevice, (cudaStream_t)stream), "zi carry");
    return RC_OK;
}

// ------------------------------------------------------------------ bandpass
struct rc_bandpass {
    int device = 0;
    long long size = 0;
    Arena arena;
    std::vector<float> taps;
    std::vector<double> g_host;
    double* d_g = nullptr;
    int K = 0;
};
int rc_bandpass_create(int device, int64_t size, double lo_hz, double hi_hz, int num_taps, const char* window,
                       rc_bandpass** out) {
    if (!out || size < 1) return fail(RC_ERR_INVALID, "bandpass: bad argument");
    DeviceGuard g(device);
    std::unique_ptr<rc_bandpass> b(new rc_bandpass());
    b->device = device; b->size = size;
    const double nyq = 0.5 * (double)size;
    int rc = firwin_bandpass(num_taps, lo_hz / nyq, hi_hz / nyq, window ? window : "hamm", b->taps);
    if (rc) return fail(rc, "bandpass: invalid cutoff frequencies, tap count or window");
    std::vector<double> gt = autocorr_taps(b->taps);
    b->g_host = gt;
    b->K = num_taps - 1;
    RC_API_CUDA(b->arena.upload(&b->d_g, gt), "taps");
    RC_API_CUDA(dev_sync(0), "sync");
    *out = b.release();
    return RC_OK;
}
int rc_bandpass_destroy(rc_bandpass* b) {
    if (!b) return RC_OK;
    DeviceGuard g(b->device);
    delete b;
    return RC_OK;
}
int rc_bandpass_taps(rc_bandpass* b, float* host, int capacity) {
    if (!b || !host || capacity < (int)b->taps.size()) return fail(RC_ERR_INVALID, "bandpass: bad argument");
    for (size_t i = 0; i < b->taps.size(); i++) host[i] = b->taps[i];
    return (int)b->taps.size();
}
int rc_bandpass_run(rc_bandpass* b, const float* in, float* outp, void* stream) {
    if (!b || !in || !outp) return fail(RC_ERR_INVALID, "bandpass: null argument");
    if (b->size <= 3 * (long long)b->taps.size())
        return fail(RC_ERR_INVALID, "The length of the input vector x must be greater than padlen");
    DeviceGuard g(b->device);
    RC_API_CUDA(launch_filtfilt(make_filtfilt(in, outp, b->d_g, b->g_host, b->size, b->K), 1, (cudaStream_t)stream, "filtfilt", b->g_host.data()),
                "filtfilt");
    return RC_OK;
}

// ----------------------------------------------------------------------- PLL
struct rc_pll {
    int device = 0;
    long long n = 0, h = 0;
    TableStore store{kOnDevice};
    Arena arena;
    FftPlan plan;
    RealResampleSpec spec{};
    float2 *Z = nullptr, *Zp = nullptr, *w0 = nullptr, *w1 = nullptr, *z = nullptr;
    bool stepped = false;
};
int rc_pll_create(int device, int64_t size, rc_pll** out) {
    if (!out || size < 2) return fail(RC_ERR_INVALID, "pll: bad argument");
    if ((size & 1) || !fft_size_supported(size)) return fail(RC_ERR_UNSUPPORTED, "pll: size must be even and factor into 2^a 3^b 5^c");
    DeviceGuard g(device);
    std::unique_ptr<rc_pll> p(new rc_pll());
    p->device = device; p->n = size; p->h = size / 2;
    RC_API_CUDA(fft_plan_build(p->plan, p->h, p->store), "plan");
    RC_API_CUDA(make_real_spec(p->spec, size, size, false, true, p->arena), "spec");
    p->spec.scale = (float)(1.0 / (double)p->h);
    RC_API_CUDA(p->arena.alloc(&p->Z, (size_t)p->h), "alloc");
    RC_API_CUDA(p->arena.alloc(&p->Zp, (size_t)p->h), "alloc");
    RC_API_CUDA(p->arena.alloc(&p->w0, (size_t)p->h), "alloc");
    RC_API_CUDA(p->arena.alloc(&p->w1, (size_t)p->h), "alloc");
    RC_API_CUDA(p->arena.alloc(&p->z, (size_t)size), "alloc");
    RC_API_CUDA(dev_sync(0), "sync");
    *out = p.release();
    return RC_OK;
}
int rc_pll_destroy(rc_pll* p) {
    if (!p) return RC_OK;
    DeviceGuard g(p->device);
    delete p;
    return RC_OK;
}
int rc_pll_step(rc_pll* p, const float* in, void* stream) {
    if (!p || !in) return fail(RC_ERR_INVALID, "pll: null argument");
    DeviceGuard g(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    RC_API_CUDA((fft_exec<-1>(p->plan, 1, LoadC64{(const float2*)in, p->h}, StoreC64{p->Z, p->h, 1.0f}, p->w0, p->w1, st)), "fft");
    RC_API_CUDA(launch_ew(p->h / 2 + 1, 1, SpecHilbertPairEw{p->spec, p->Z, p->Zp}, st), "hilbert");
    RC_API_CUDA((fft_exec<+1>(p->plan, 1, LoadC64{p->Zp, p->h}, StoreAnalyticPacked{in, p->z, p->n}, p->w0, p->w1, st)), "ifft");
    p->stepped = true;
    return RC_OK;
}
int rc_pll_eval(rc_pll* p, double mult, int imag, float* outp, void* stream) {
    if (!p || !outp) return fail(RC_ERR_INVALID, "pll: null argument");
    if (!p->stepped) return fail(RC_ERR_STATE, "pll: eval before step");
    DeviceGuard g(p->device);
    RC_API_CUDA(launch_ew(p->n, 1, PllEvalEw{p->z, outp, p->n, (float)mult, imag}, (cudaStream_t)stream), "pll eval");
    return RC_OK;
}

// ------------------------------------------------------------------ profiling
int rc_profile_enable(int on) {
    Profiler& p = profiler();
    p.on = on != 0;
    return RC_OK;
}
int64_t rc_profile_launches(void) { return (int64_t)profiler().launches; }
int rc_profile_reset(void) {
    Profiler& p = profiler();
#ifndef RC_EMULATE
    for (auto& r : p.recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
#endif
    p.recs.clear();
    p.launches = 0;
    return RC_OK;
}
// JSON: {"tag": {"count": n, "total_ms": t, "bytes_per_launch": b}, ...}; returns the length needed.
int rc_profile_report(char* buf, int capacity) {
    Profiler& p = profiler();
    std::map<std::string, std::pair<int, std::pair<double, double>>> agg;   // count, (ms, bytes)
    std::vector<std::string> order;
#ifndef RC_EMULATE
    for (auto& r : p.recs) {
        cudaEventSynchronize(r.b);
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) ms = 0.f;
        if (!agg.count(r.tag)) order.push_back(r.tag);
        auto& a = agg[r.tag];
        a.first += 1; a.second.first += ms; a.second.second = r.bytes;
    }
#endif
    std::string out = "{";
    for (size_t i = 0; i < order.size(); i++) {
        auto& a = agg[order[i]];
        char line[256];
        snprintf(line, sizeof(line), "%s\"%s\": {\"count\": %d, \"total_ms\": %.6f, \"bytes_per_launch\": %.1f}",
                 i ? ", " : "", order[i].c_str(), a.first, a.second.first, a.second.second);
        out += line;
    }
    out += "}";
    if (buf && capacity > 0) {
        strncpy(buf, out.c_str(), (size_t)capacity - 1);
        buf[capacity - 1] = 0;
    }
    return (int)out.size() + 1;
}

// ------------------------------------------------------- sharded-load pieces
int rc_subband_combine(int device, int n_ranks, int64_t piece_len, int64_t n_input, int64_t k0_base,
                       const void* pieces_dev, void* bins_dev, void* stream) {
    if (!pieces_dev || !bins_dev || piece_len < 2 || n_input < 1 || k0_base < 0)
        return fail(RC_ERR_INVALID, "subband_combine: bad argument");
    if ((piece_len & 1) || (((size_t)pieces_dev | (size_t)bins_dev) & 15))
        return fail(RC_ERR_UNSUPPORTED, "subband_combine: even piece length and 16-byte aligned buffers");
    DeviceGuard g(device);
    cudaStream_t st = (cudaStream_t)stream;
    const double bytes = 16.0 * (double)piece_len * n_ranks;
    const float2* F = (const float2*)pieces_dev;
    float2* Y = (float2*)bins_dev;
    const double m2n = -2.0 / (double)n_input;
    cudaError_t err;
    switch (n_ranks) {
        case 2: err = launch_ew(piece_len, 1, SubbandCombineEw<2>{F, Y, piece_len, k0_base, m2n}, st, "tuner.subband_combine", bytes); break;
        case 4: err = launch_ew(piece_len, 1, SubbandCombineEw<4>{F, Y, piece_len, k0_base, m2n}, st, "tuner.subband_combine", bytes); break;
        case 8: err = launch_ew(piece_len, 1, SubbandCombineEw<8>{F, Y, piece_len, k0_base, m2n}, st, "tuner.subband_combine", bytes); break;
        case 16: err = launch_ew(piece_len, 1, SubbandCombineEw<16>{F, Y, piece_len, k0_base, m2n}, st, "tuner.subband_combine", bytes); break;
        default: return fail(RC_ERR_UNSUPPORTED, "subband_combine: 2, 4, 8 or 16 ranks");
    }
    RC_API_CUDA(err, "subband combine");
    return RC_OK;
}

int rc_subband_combine_scatter(int device, int n_ranks, int64_t piece_len, int64_t n_input, int64_t k0_base,
                               const void* pieces_dev, const rc_scatter_seg* segs, int n_segs, void* stream) {
    if (!pieces_dev || !segs || n_segs < 1 || piece_len < 2 || n_input < 1 || k0_base < 0)
        return fail(RC_ERR_INVALID, "subband_combine_scatter: bad argument");
    if ((piece_len & 1) || (((size_t)pieces_dev) & 15))
        return fail(RC_ERR_UNSUPPORTED, "subband_combine_scatter: even piece length and 16-byte aligned pieces");
    ScatterTable tab;
    memset(&tab, 0, sizeof(tab));
    for (int i = 0; i < n_segs; i++) {
        const rc_scatter_seg& sg = segs[i];
        if (sg.k1 < 0 || sg.k1 >= n_ranks || sg.k1 >= kMaxRanks || !sg.dst || sg.j_lo < 0 || sg.j_hi > piece_len || sg.j_lo >= sg.j_hi)
            return fail(RC_ERR_INVALID, "subband_combine_scatter: bad segment");
        int& n = tab.n[sg.k1];
        if (n >= kScatterSegs) return fail(RC_ERR_UNSUPPORTED, "subband_combine_scatter: more than 4 segments for one k1");
        tab.lo[sg.k1][n] = sg.j_lo; tab.hi[sg.k1][n] = sg.j_hi; tab.dst[sg.k1][n] = (float2*)sg.dst;
        n++;
    }
    DeviceGuard g(device);
    cudaStream_t st = (cudaStream_t)stream;
    const double bytes = 16.0 * (double)piece_len * n_ranks;
    const float2* F = (const float2*)pieces_dev;
    const double m2n = -2.0 / (double)n_input;
    // NVLink-bound: a few resident CTAs per SM keep the links busy and leave the rest of the SM to the
    // channel kernels of the previous block running beside it (RC_SCATTER_CTAS: experiments)
    int cps = 2;
    if (const char* env = getenv("RC_SCATTER_CTAS")) cps = atoi(env);
    cudaError_t err;
    switch (n_ranks) {
        case 2: err = launch_ew(piece_len, 1, SubbandCombineScatterEw<2>{F, piece_len, k0_base, m2n, tab}, st, "tuner.subband_combine_scatter", bytes, cps); break;
        case 4: err = launch_ew(piece_len, 1, SubbandCombineScatterEw<4>{F, piece_len, k0_base, m2n, tab}, st, "tuner.subband_combine_scatter", bytes, cps); break;
        case 8: err = launch_ew(piece_len, 1, SubbandCombineScatterEw<8>{F, piece_len, k0_base, m2n, tab}, st, "tuner.subband_combine_scatter", bytes, cps); break;
        case 16: err = launch_ew(piece_len, 1, SubbandCombineScatterEw<16>{F, piece_len, k0_base, m2n, tab}, st, "tuner.subband_combine_scatter", bytes, cps); break;
        default: return fail(RC_ERR_UNSUPPORTED, "subband_combine_scatter: 2, 4, 8 or 16 ranks");
    }
    RC_API_CUDA(err, "subband combine scatter");
    return RC_OK;
}

struct rc_fft {
    int device = 0, batch = 1;
    long long n = 0;
    TableStore store{kOnDevice};
    Arena arena;
    FftPlan plan;
    float2 *w0 = nullptr, *w1 = nullptr;
};
int rc_fft_create(int device, int64_t n, int batch, rc_fft** out) {
    if (!out || n < 1 || batch < 1) return fail(RC_ERR_INVALID, "fft: bad argument");
    if (!fft_size_supported(n)) return fail(RC_ERR_UNSUPPORTED, "fft: size must factor into 2^a 3^b 5^c");
    DeviceGuard g(device);
    std::unique_ptr<rc_fft> f(new rc_fft());
    f->device = device; f->n = n; f->batch = batch;
    RC_API_CUDA(fft_plan_build(f->plan, n, f->store), "plan");
    if (f->plan.max_passes() >= 2) RC_API_CUDA(f->arena.alloc(&f->w0, (size_t)n * batch), "alloc");
    if (f->plan.max_passes() >= 3) RC_API_CUDA(f->arena.alloc(&f->w1, (size_t)n * batch), "alloc");
    RC_API_CUDA(dev_sync(0), "table sync");
    *out = f.release();
    return RC_OK;
}
int rc_fft_destroy(rc_fft* f) {
    if (!f) return RC_OK;
    DeviceGuard g(f->device);
    delete f;
    return RC_OK;
}
int rc_fft_exec(rc_fft* f, int sign, const void* in, void* outp, void* stream) {
    if (!f || !in || !outp || in == outp || (sign != 1 && sign != -1)) return fail(RC_ERR_INVALID, "fft: bad argument");
    DeviceGuard g(f->device);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e;
    if (sign < 0) e = fft_exec<-1>(f->plan, f->batch, LoadC64{(const float2*)in, f->n}, StoreC64{(float2*)outp, f->n, 1.0f}, f->w0, f->w1, st, "tuner.local_fft");
    else e = fft_exec<+1>(f->plan, f->batch, LoadC64{(const float2*)in, f->n}, StoreC64{(float2*)outp, f->n, 1.0f}, f->w0, f->w1, st, "tuner.local_fft");
    RC_API_CUDA(e, "fft exec");
    return RC_OK;
}

int rc_fft_exec_scatter(rc_fft* f, int sign, const void* in, void* const* piece_bases, int n_pieces,
                        int64_t piece_len, void* stream) {
    if (!f || !in || !piece_bases || (sign != 1 && sign != -1)) return fail(RC_ERR_INVALID, "fft: bad argument");
    if (f->batch != 1 || n_pieces < 1 || n_pieces > kMaxRanks || piece_len < 2 || (piece_len & 1) ||
        (long long)n_pieces * piece_len != f->n || f->n >= (1LL << 31))
        return fail(RC_ERR_UNSUPPORTED, "fft scatter: batch 1, <= 16 even pieces that tile the transform");
    StoreScatterC64 st_op;
    memset(&st_op, 0, sizeof(st_op));
    for (int i = 0; i < n_pieces; i++) {
        if (!piece_bases[i] || (((size_t)piece_bases[i]) & 15)) return fail(RC_ERR_INVALID, "fft scatter: piece bases must be 16-byte aligned");
        st_op.base[i] = (float2*)piece_bases[i];
    }
    st_op.P = (unsigned)piece_len;
    DeviceGuard g(f->device);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e;
    if (sign < 0) e = fft_exec<-1>(f->plan, 1, LoadC64{(const float2*)in, f->n}, st_op, f->w0, f->w1, st, "tuner.local_fft");
    else e = fft_exec<+1>(f->plan, 1, LoadC64{(const float2*)in, f->n}, st_op, f->w0, f->w1, st, "tuner.local_fft");
    RC_API_CUDA(e, "fft exec scatter");
    return RC_OK;
}

// ----------------------------------------------------------------- FFT hook
int rc_fft_c2c(int device, int64_t n, int batch, int sign, const void* in, void* outp, void* stream) {
    if (!in || !outp || n < 1 || batch < 1 || (sign != 1 && sign != -1)) return fail(RC_ERR_INVALID, "fft: bad argument");
    if (!fft_size_supported(n)) return fail(RC_ERR_UNSUPPORTED, "fft: size must factor into 2^a 3^b 5^c");
    DeviceGuard g(device);
    cudaStream_t st = (cudaStream_t)stream;
    TableStore store(kOnDevice);
    Arena arena;
    FftPlan plan;
    RC_API_CUDA(fft_plan_build(plan, n, store), "plan");
    float2 *w0 = nullptr, *w1 = nullptr;
    RC_API_CUDA(arena.alloc(&w0, (size_t)n * batch), "alloc");
    RC_API_CUDA(arena.alloc(&w1, (size_t)n * batch), "alloc");
    RC_API_CUDA(dev_sync(0), "table sync");     // twiddle tables were uploaded on the default stream
    cudaError_t e;
    if (sign < 0) e = fft_exec<-1>(plan, batch, LoadC64{(const float2*)in, n}, StoreC64{(float2*)outp, n, 1.0f}, w0, w1, st);
    else e = fft_exec<+1>(plan, batch, LoadC64{(const float2*)in, n}, StoreC64{(float2*)outp, n, 1.0f}, w0, w1, st);
    RC_API_CUDA(e, "fft exec");
    RC_API_CUDA(dev_sync(st), "fft sync");      // scratch and tables are freed on return
    return RC_OK;
}

}  // extern "C"
