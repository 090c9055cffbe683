// Register-radix FFT pass kernels, schedule group 2 (see rc_fft.cuh RC_V2_GROUP2).
#include "rc_fft2_inst.cuh"
namespace rc {
RC_V2_DEFINE_GROUP(2, RC_V2_GROUP2)
}
