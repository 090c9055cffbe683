// CPU replay of the FFT pass kernel's per-thread phases (no GPU needed).
// Build: nvcc -O2 -std=c++17 -I radio-core_b200/csrc tests/native/emulate_fft.cu -o /tmp/emulate_fft
// Checks random bins of every pass plan against a long-double direct DFT.
#include <cstdio>
#include <random>
#include <vector>
#include "rc_fft.cuh"

using namespace rc;

template <int SIGN>
static void emulate(const FftPlan& plan, const std::vector<float2>& in, std::vector<float2>& out, int batch) {
    std::vector<float2> w0((size_t)plan.n * batch), w1((size_t)plan.n * batch);
    out.assign((size_t)plan.n * batch, make_float2(0, 0));
    for (int i = 0; i < plan.npass; i++) {
        const FftPass& P = plan.pass[i];
        const float2* src = i == 0 ? in.data() : (i == 1 ? w0.data() : w1.data());
        float2* dst = i == plan.npass - 1 ? out.data() : (i == 0 ? w0.data() : w1.data());
        LoadC64 ld{src, plan.n};
        StoreC64 st{dst, plan.n, 1.0f};
        long long tiles = (P.stride + P.T - 1) / P.T;
        std::vector<float2> sm(P.smem_elems);
        for (int b = 0; b < batch; b++)
            for (long long tile = 0; tile < tiles; tile++) {
                long long j0 = tile * P.T;
                for (int tid = 0; tid < P.threads; tid++) fft_pass_load<LoadC64, SIGN>(sm.data(), P, ld, b, j0, tid, P.threads);
                int Lprev = 1;
                for (int s = 0; s < P.nstage; s++) {
                    for (int tid = 0; tid < P.threads; tid++) fft_stage_dispatch<SIGN>(sm.data(), P, P.radix[s], Lprev, tid, P.threads);
                    Lprev *= P.radix[s];
                }
                for (int tid = 0; tid < P.threads; tid++) fft_pass_store<StoreC64>(sm.data(), P, st, b, j0, tid, P.threads);
            }
    }
}

static double check(long long n, int batch, int sign, int nbins) {
    TableStore store(false);
    FftPlan plan;
    if (fft_plan_build(plan, n, store) != cudaSuccess) { printf("n=%lld plan failed\n", n); return 1e9; }
    std::mt19937_64 rng(n * 7 + sign);
    std::uniform_real_distribution<float> U(-1.f, 1.f);
    std::vector<float2> in((size_t)n * batch), out;
    for (auto& v : in) v = make_float2(U(rng), U(rng));
    if (sign < 0) emulate<-1>(plan, in, out, batch); else emulate<1>(plan, in, out, batch);
    double worst = 0, scale = sqrt((double)n);
    for (int t = 0; t < nbins; t++) {
        long long k = (t == 0) ? 0 : (t == 1 ? n - 1 : (long long)(rng() % n));
        int b = t % batch;
        long double re = 0, im = 0;
        for (long long j = 0; j < n; j++) {
            long double a = sign * 2.0L * 3.14159265358979323846264338327950288L * (long double)((j * k) % n) / n;
            long double c = cosl(a), s = sinl(a);
            re += in[(size_t)b * n + j].x * c - in[(size_t)b * n + j].y * s;
            im += in[(size_t)b * n + j].x * s + in[(size_t)b * n + j].y * c;
        }
        double e = hypot((double)(out[(size_t)b * n + k].x - re), (double)(out[(size_t)b * n + k].y - im)) / scale;
        if (e > worst) worst = e;
    }
    printf("n=%-9lld batch=%d sign=%+d passes=%d [", n, batch, sign, plan.npass);
    for (int i = 0; i < plan.npass; i++) {
        printf(" %d(T%d:", plan.pass[i].R, plan.pass[i].T);
        for (int s = 0; s < plan.pass[i].nstage; s++) printf("%s%d", s ? "," : "", plan.pass[i].radix[s]);
        printf(")");
    }
    printf(" ] rel.err=%.3g\n", worst);
    return worst;
}

int main(int argc, char** argv) {
    double worst = 0;
    long long sizes[] = {1, 2, 3, 4, 5, 6, 8, 10, 16, 25, 30, 48, 100, 125, 360, 1000, 640, 625, 4800, 24000,
                         25000, 31250, 48000, 250000, 240000, 1000000};
    for (long long n : sizes) {
        int nb = n > 100000 ? 6 : 12;
        worst = fmax(worst, check(n, n < 2000 ? 3 : 1, -1, nb));
        worst = fmax(worst, check(n, n < 2000 ? 2 : 1, +1, nb));
    }
    if (argc > 1) worst = fmax(worst, check(atoll(argv[1]), 1, -1, 4));
    printf("worst %.3g\n", worst);
    return worst < 3e-6 ? 0 : 1;
}
