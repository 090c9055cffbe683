// Register-radix FFT pass kernels, schedule group 3 (see rc_fft.cuh RC_V3_GROUP3).
#include "rc_fft3_inst.cuh"
namespace rc {
RC_V3_DEFINE_GROUP(3, RC_V3_GROUP3)
}
