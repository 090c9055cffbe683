// Print the pass plan the FFT engine picks for given sizes (host only).
//   nvcc -DRC_EMULATE -O1 -std=c++17 -I radio-core_b200/csrc -o tools/bin/plan_dump tools/plan_dump.cu
#include <cstdio>
#include <cstdlib>
#include "rc_fft.cuh"
using namespace rc;
int main(int argc, char** argv) {
    for (int a = 1; a < argc; a++) {
        long long n = atoll(argv[a]);
        std::vector<int> fs;
        bool ok = fft_choose_fast(n, fs);
        printf("%lld:", n);
        if (!ok) printf(" (no fast split)");
        double c = 0;
        for (size_t i = 0; i < fs.size(); i++) { printf(" %d", fs[i]); c += fft_pass_cost(fs[i], i == 0); }
        printf("   cost %.2f\n", c);
    }
    return 0;
}
