// Register-radix FFT pass kernels, schedule group 3 (see rc_fft.cuh RC_V2_GROUP3).
#include "rc_fft2_inst.cuh"
namespace rc {
RC_V2_DEFINE_GROUP(3, RC_V2_GROUP3)
}
