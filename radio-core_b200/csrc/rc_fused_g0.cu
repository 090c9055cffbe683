// Fused last-two-passes kernels (rc_fused.cuh) for the schedule pairs of RC_FUSED_LIST (rc_fft.cuh).
#include "rc_fused.cuh"
namespace rc {
template <int ID> struct FusedSched { typedef typename V3ById<ID>::type type; };
template <> struct FusedSched<24> { typedef V3Sched<10, 1, 10, 80, 6, 16> type; };     // R = 100, 32 columns
template <> struct FusedSched<27> { typedef V3Sched<5, 1, 10, 80, 6, 16> type; };      // R = 50, 32 columns
cudaError_t v3_fused_dispatch(int id_a, int id_b, int sign, const FusedPair& f, const LoadAny& ld_a, const StoreAny& st_b,
                              cudaStream_t stream) {
// Inside the fused kernel a tile is worked by half as many threads as in the stand-alone pass
// kernels (each thread takes two row groups): twice the CTAs per SM hide the dependency
// look-ups and the serial phases of a tile.
#define RC_FUSED_CASE(a, b)                                                                                           \
    if (id_a == a && id_b == b)                                                                                        \
        return sign < 0 ? v3_run_fused_ll<FusedSched<a>::type, FusedSched<b>::type, -1>(f, ld_a, st_b, stream)        \
                        : v3_run_fused_ll<FusedSched<a>::type, FusedSched<b>::type, +1>(f, ld_a, st_b, stream);
    RC_FUSED_LIST(RC_FUSED_CASE)
#undef RC_FUSED_CASE
    return cudaErrorInvalidValue;
}
}  // namespace rc
