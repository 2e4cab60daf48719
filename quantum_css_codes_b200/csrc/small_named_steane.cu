// Static (compile-time H, L, truth tables) decode kernels for the steane descriptor.
#include "named_codes.inc"
#include "small_common.cuh"

namespace qcss {
cudaError_t launch_small_steane(const SmallLaunch& l, cudaStream_t stream) {
    return small::launch_named<named::Steane_X, named::Steane_Z>(l, stream);
}
}  // namespace qcss
