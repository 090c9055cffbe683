// Register-radix FFT pass kernels, schedule group 0 (see rc_fft.cuh RC_V2_GROUP0).
#include "rc_fft2_inst.cuh"
namespace rc {
RC_V2_DEFINE_GROUP(0, RC_V2_GROUP0)
}
