// Fused last-two-passes kernels (rc_fused.cuh) for the schedule pairs of RC_FUSED_LIST (rc_fft.cuh).
#include "rc_fused.cuh"
namespace rc {
cudaError_t v3_fused_dispatch(int id_a, int id_b, int sign, const FusedPair& f, const LoadAny& ld_a, const StoreAny& st_b,
                              cudaStream_t stream) {
#define RC_FUSED_CASE(a, b)                                                                                           \
    if (id_a == a && id_b == b)                                                                                        \
        return sign < 0 ? v3_run_fused_ll<V3ById<a>::type, V3ById<b>::type, -1>(f, ld_a, st_b, stream)                \
                        : v3_run_fused_ll<V3ById<a>::type, V3ById<b>::type, +1>(f, ld_a, st_b, stream);
    RC_FUSED_LIST(RC_FUSED_CASE)
#undef RC_FUSED_CASE
    return cudaErrorInvalidValue;
}
}  // namespace rc
