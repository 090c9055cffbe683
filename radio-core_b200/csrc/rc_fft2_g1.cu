// Register-radix FFT pass kernels, schedule group 1 (see rc_fft.cuh RC_V2_GROUP1).
#include "rc_fft2_inst.cuh"
namespace rc {
RC_V2_DEFINE_GROUP(1, RC_V2_GROUP1)
}
