// rc_fuse_ad.cuh -- ONE kernel for the seam between Tuner.run and FM.run:
//
//   last pass of the channel inverse FFT  ->  angle(y)/pi  ->  wrapped phase difference
//   (fm.py:60-65)  ->  first pass of the packed real FFT of Decimate (decimate.py:48)
//
// Stand-alone, the last IFFT pass stores angle(y)/pi (4 bytes per sample) and the first pass of the
// real FFT reads it back: 2 x 4 bytes per channel sample of HBM traffic and one launch that carry
// no information the CTA did not already hold.  With R = 100 for both passes the tiles coincide:
// the last IFFT pass of a B-point transform (stride B/100 columns) produces, for 32 adjacent
// columns j and all 100 rows K, the samples n = j + K*B/100 -- exactly the 16 packed columns x 100
// rows (row stride B/200 packed elements) one tile of the first real-FFT pass consumes.  The only
// value a tile lacks is the sample BEFORE its first column (the discriminator differences
// neighbours); ten extra threads recompute that one column pair of the inverse FFT from a second,
// two-column TMA box (6 % more butterflies) instead of any exchange between CTAs.  For tile 0 the
// predecessor is the last column of the previous row.
//
// Thread roles (192 threads): 0..159 the 16 column pairs x 10 row groups of the IFFT pass
// (schedule 24), 160..169 the halo pair (same schedule with one column pair), 0..79 the real-FFT
// pass (schedule 0).  Every phase is the per-thread code of rc_fft3.cuh; only the sources and
// sinks are new, so the CPU replay (RC_EMULATE) runs the same functions.
#pragma once

#include "rc_exec.cuh"
#include "rc_fft3.cuh"

namespace rc {

typedef V3Sched<10, 1, 10, 160, 5, 16> FuseSA;      // last IFFT pass: R = 100, 32 columns        (schedule 24)
typedef V3Sched<10, 1, 10, 10, 1, 1> FuseSH;        // the same pass on the halo column pair
typedef V3Sched<10, 1, 10, 80, 10, 8> FuseSB;       // first real-FFT pass: R = 100, 16 packed columns (schedule 0)
constexpr int kFuseThreads = 192;
constexpr int kFuseR = 100;
constexpr int kFuseAngPitch = 34;                   // [0] unused, [1] the sample before the tile, [2..33] the 32 columns
constexpr int kFuseAngRows = kFuseR + 1;            // tile 0 stores its halo one row down
// shared memory: IFFT tile | halo pair | angle tile | two W_100 tables | mbarrier
constexpr int kFuseOffHalo = kFuseR * 16 * 16;                               // 25600
constexpr int kFuseOffAng = kFuseOffHalo + kFuseR * 16;                      // + 1600
constexpr int kFuseOffTw = kFuseOffAng + (kFuseAngRows * kFuseAngPitch * 4 + 15) / 16 * 16;
constexpr int kFuseOffBar = kFuseOffTw + 2 * kFuseR * 8;
constexpr int kFuseSmem = kFuseOffBar + 16;
static_assert(FuseSB::TILE_F4 * 16 <= kFuseOffHalo, "the real-FFT tile re-uses the IFFT tile");

// sink of the IFFT's last stage: angle(y)/pi into the shared angle tile (index = V3Out arithmetic)
struct StoreAngleSmem {
    float* ang;
    RC_HD void operator()(int, long long i, float2 v) const { ang[i] = atan2pi_fast(v.y, v.x); }
    RC_HD void pair(int, long long i, float2 v, float2 w) const {
        ang[i] = atan2pi_fast(v.y, v.x);
        ang[i + 1] = atan2pi_fast(w.y, w.x);
    }
};

// source of the real FFT's first stage: packed discriminator (LoadAnglePacked's arithmetic) from
// the shared angle tile; column pair cp = samples 4 cp .. 4 cp + 3 of the row
struct V3FromAngSmem {
    static constexpr bool kTile = true;
    const float* ang;
    struct Ctx {};
    RC_HD Ctx prepare(int) const { return Ctx{}; }
    RC_HD float4 get(const Ctx&, int row, int cp, long long, bool) const {
        const float* a = ang + row * kFuseAngPitch + 2 + 4 * cp;
        const float p = a[-1], x0 = a[0], x1 = a[1], x2 = a[2], x3 = a[3];
        return make_float4(wrap_half_turns_dev(x0 - p), wrap_half_turns_dev(x1 - x0), wrap_half_turns_dev(x2 - x1),
                           wrap_half_turns_dev(x3 - x2));
    }
};

struct FuseAdArgs {
    FftPass PA, PB;            // last pass of the inverse FFT (B points), first pass of the real FFT (B/2 points)
    StoreC64 stB;              // first intermediate of the real FFT
};

RC_HD V3Out fuse_out_main(const FftPass& PA, long long j0, int tid) {
    const int cp = tid & (FuseSA::CP - 1);
    const long long j = j0 + 2 * cp;
    V3Out o;
    o.act_a = j < PA.stride; o.act_b = j + 1 < PA.stride;
    o.ns = kFuseAngPitch; o.oa = 2 * cp + 2; o.ob = 2 * cp + 3; o.pair = false;
    return o;
}
RC_HD V3Out fuse_out_halo(long long tile) {
    V3Out o;
    o.act_a = o.act_b = true;
    o.ns = kFuseAngPitch;
    o.oa = tile == 0 ? kFuseAngPitch : 0;           // tile 0: column stride - 1 of row K precedes column 0 of row K + 1
    o.ob = o.oa + 1;
    o.pair = false;
    return o;
}

#if defined(__CUDACC__) && !defined(RC_EMULATE) && defined(RC_FUSE_AD_IMPL)      // defined once, in rc_fuse_ad.cu
__global__ void __launch_bounds__(kFuseThreads, 4)
v3_fuse_ad_kernel(const FuseAdArgs A, const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapH) {
    extern __shared__ __align__(128) unsigned char rc_fuse_smem[];
    float4* tileA = (float4*)rc_fuse_smem;
    float4* halo = (float4*)(rc_fuse_smem + kFuseOffHalo);
    float* ang = (float*)(rc_fuse_smem + kFuseOffAng);
    float2* twA = (float2*)(rc_fuse_smem + kFuseOffTw);
    float2* twB = twA + kFuseR;
    uint64_t* bar = (uint64_t*)(rc_fuse_smem + kFuseOffBar);
    const int batch = blockIdx.y + blockIdx.z * gridDim.y;
    const long long tile = blockIdx.x, j0 = tile * FuseSA::T;
    if (j0 >= A.PA.stride) return;                  // padding CTA of the last cluster
    const long long hj = tile > 0 ? j0 - 2 : A.PA.stride - 2;
    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, (uint32_t)(kFuseR * (FuseSA::T + 2) * sizeof(float2)));
        tma_load_3d(tileA, &tmapA, bar, (int)(j0 * 2), 0, batch);
        tma_load_3d(halo, &tmapH, bar, (int)(hj * 2), 0, batch);
    }
    v3_load_table<FuseSA, +1>(twA, A.PA, tid);
    v3_load_table<FuseSB, -1>(twB, A.PB, tid);
    V3Tw tws;
    if (tid < FuseSA::NT) tws = v3_twiddle_setup<FuseSA, true>(A.PA, j0, tid);
    else if (tid < FuseSA::NT + FuseSH::NT) tws = v3_twiddle_setup<FuseSH, true>(A.PA, hj, tid - FuseSA::NT);
    __syncthreads();
    mbar_wait(bar, 0);
    if (tid < FuseSA::NT) v3_stage0<FuseSA, +1, true>(tileA, twA, A.PA, V3FromTile<FuseSA::CP>{tileA}, batch, j0, tid, tws);
    else if (tid < FuseSA::NT + FuseSH::NT)
        v3_stage0<FuseSH, +1, true>(halo, twA, A.PA, V3FromTile<1>{halo}, batch, hj, tid - FuseSA::NT, tws);
    __syncthreads();
    if (tid < FuseSA::NT) v3_last_direct<FuseSA, +1>(tileA, StoreAngleSmem{ang}, 0, fuse_out_main(A.PA, j0, tid), tid);
    else if (tid < FuseSA::NT + FuseSH::NT)
        v3_last_direct<FuseSH, +1>(halo, StoreAngleSmem{ang}, 0, fuse_out_halo(tile), tid - FuseSA::NT);
    __syncthreads();
    if (tile == 0) {                                // d[0] = 0: the block's first sample has no predecessor (fm.py:63)
        if (tid == 0) ang[1] = ang[2];
        __syncthreads();
    }
    float4* tileB = tileA;
    const long long j0B = tile * FuseSB::T;
    if (tid < FuseSB::NT) {
        const V3Tw twsB = v3_twiddle_setup<FuseSB, false>(A.PB, j0B, tid);
        v3_stage0<FuseSB, -1, false>(tileB, twB, A.PB, V3FromAngSmem{ang}, batch, j0B, tid, twsB);
    }
    __syncthreads();
    float4 hold[FuseSB::HOLD];
    if (tid < FuseSB::NT) v3_last_first_a<FuseSB, -1>(tileB, hold, tid);
    __syncthreads();
    if (tid < FuseSB::NT) v3_last_first_b<FuseSB>((float2*)tileB, hold, tid);
    __syncthreads();
    if (tid < FuseSB::NT) v3_first_copy_out<FuseSB>((const float2*)tileB, A.PB, A.stB, batch, j0B, tid);
}
#endif

// srcA: the input of the inverse FFT's last pass ([batch][B] complex64, ping-pong buffer of the plan)
cudaError_t v3_run_fuse_ad(const FuseAdArgs& A, const LoadC64& srcA, int batch, cudaStream_t stream);
bool v3_fuse_ad_possible(const FftPass& PA, const FftPass& PB, const LoadC64& srcA, int batch);

}  // namespace rc
