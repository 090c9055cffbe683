// Register-radix FFT pass kernels, schedule group 2 (see rc_fft.cuh RC_V3_GROUP2).
#include "rc_fft3_inst.cuh"
namespace rc {
RC_V3_DEFINE_GROUP(2, RC_V3_GROUP2)
}
