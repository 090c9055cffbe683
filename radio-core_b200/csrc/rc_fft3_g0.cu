// Register-radix FFT pass kernels, schedule group 0 (see rc_fft.cuh RC_V3_GROUP0).
#include "rc_fft3_inst.cuh"
namespace rc {
RC_V3_DEFINE_GROUP(0, RC_V3_GROUP0)
}
