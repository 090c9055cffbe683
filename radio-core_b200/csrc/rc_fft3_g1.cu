// Register-radix FFT pass kernels, schedule group 1 (see rc_fft.cuh RC_V3_GROUP1).
#include "rc_fft3_inst.cuh"
namespace rc {
RC_V3_DEFINE_GROUP(1, RC_V3_GROUP1)
}
