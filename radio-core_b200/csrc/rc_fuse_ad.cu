// Launcher (and CPU replay) of the fused "last IFFT pass -> discriminator -> first real-FFT pass" kernel.
#define RC_FUSE_AD_IMPL
#include "rc_fuse_ad.cuh"

#include "rc_fft3_inst.cuh"

namespace rc {

bool v3_fuse_ad_possible(const FftPass& PA, const FftPass& PB, const LoadC64& srcA, int batch) {
    if (getenv("RC_NO_FUSE_AD")) return false;
    if (PA.fast_id != 24 || PB.fast_id != 0 || PA.R != kFuseR || PB.R != kFuseR) return false;   // schedules 24 / 0
    if (PA.Ns != PA.stride || PB.Ns != 1 || PA.stride != 2 * PB.stride || PA.stride < 4 || PA.n >= (1LL << 31)) return false;
#ifdef RC_EMULATE
    (void)srcA; (void)batch;
    return PA.stride % 2 == 0;
#else
    TileSource src{srcA.p, PA.stride, srcA.batch_stride, PA.stride, PA.R, batch};
    return tma_source_ok(src) && tma_box_rows(PA.R) == PA.R;
#endif
}

cudaError_t v3_run_fuse_ad(const FuseAdArgs& A, const LoadC64& srcA, int batch, cudaStream_t stream) {
    const long long tiles = (A.PA.stride + FuseSA::T - 1) / FuseSA::T;
#ifdef RC_EMULATE
    (void)stream;
    std::vector<float4> tileA((size_t)kFuseR * FuseSA::CP + 64), halo(kFuseR), hold((size_t)FuseSB::NT * FuseSB::HOLD);
    std::vector<float> ang((size_t)kFuseAngRows * kFuseAngPitch);
    std::vector<float2> twA(kFuseR), twB(kFuseR);
    for (int tid = 0; tid < FuseSA::NT; tid++) v3_load_table<FuseSA, +1>(twA.data(), A.PA, tid);
    for (int tid = 0; tid < FuseSB::NT; tid++) v3_load_table<FuseSB, -1>(twB.data(), A.PB, tid);
    for (int b = 0; b < batch; b++)
        for (long long tile = 0; tile < tiles; tile++) {
            const long long j0 = tile * FuseSA::T, hj = tile > 0 ? j0 - 2 : A.PA.stride - 2;
            v3_emulate_tma<FuseSA>(tileA.data(), srcA, A.PA, b, j0);
            v3_emulate_tma<FuseSH>(halo.data(), srcA, A.PA, b, hj);
            for (int tid = 0; tid < FuseSA::NT; tid++)
                v3_stage0<FuseSA, +1, true>(tileA.data(), twA.data(), A.PA, V3FromTile<FuseSA::CP>{tileA.data()}, b, j0, tid,
                                            v3_twiddle_setup<FuseSA, true>(A.PA, j0, tid));
            for (int tid = 0; tid < FuseSH::NT; tid++)
                v3_stage0<FuseSH, +1, true>(halo.data(), twA.data(), A.PA, V3FromTile<1>{halo.data()}, b, hj, tid,
                                            v3_twiddle_setup<FuseSH, true>(A.PA, hj, tid));
            for (int tid = 0; tid < FuseSA::NT; tid++)
                v3_last_direct<FuseSA, +1>(tileA.data(), StoreAngleSmem{ang.data()}, 0, fuse_out_main(A.PA, j0, tid), tid);
            for (int tid = 0; tid < FuseSH::NT; tid++)
                v3_last_direct<FuseSH, +1>(halo.data(), StoreAngleSmem{ang.data()}, 0, fuse_out_halo(tile), tid);
            if (tile == 0) ang[1] = ang[2];
            float4* tileB = tileA.data();
            const long long j0B = tile * FuseSB::T;
            for (int tid = 0; tid < FuseSB::NT; tid++)
                v3_stage0<FuseSB, -1, false>(tileB, twB.data(), A.PB, V3FromAngSmem{ang.data()}, b, j0B, tid,
                                             v3_twiddle_setup<FuseSB, false>(A.PB, j0B, tid));
            for (int tid = 0; tid < FuseSB::NT; tid++) v3_last_first_a<FuseSB, -1>(tileB, hold.data() + (size_t)tid * FuseSB::HOLD, tid);
            for (int tid = 0; tid < FuseSB::NT; tid++) v3_last_first_b<FuseSB>((float2*)tileB, hold.data() + (size_t)tid * FuseSB::HOLD, tid);
            for (int tid = 0; tid < FuseSB::NT; tid++) v3_first_copy_out<FuseSB>((const float2*)tileB, A.PB, A.stB, b, j0B, tid);
        }
    return cudaSuccess;
#else
    TileSource src{srcA.p, A.PA.stride, srcA.batch_stride, A.PA.stride, A.PA.R, batch};
    CUtensorMap tmapA, tmapH;
    if (!tma_encode_tile_map(&tmapA, src, kFuseR, FuseSA::T) || !tma_encode_tile_map(&tmapH, src, kFuseR, 2))
        return cudaErrorInvalidValue;
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(v3_fuse_ad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFuseSmem);
        if (e != cudaSuccess) return e;
        configured[dev & 63] = true;
    }
    return v3_launch(v3_fuse_ad_kernel, tiles, batch, kFuseThreads, (size_t)kFuseSmem, stream, A, tmapA, tmapH);
#endif
}

}  // namespace rc
