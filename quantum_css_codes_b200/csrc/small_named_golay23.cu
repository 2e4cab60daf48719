// Static (compile-time H, L, truth tables) decode kernels for the golay23 descriptor.
#include "named_codes.inc"
#include "small_common.cuh"

namespace qcss {
cudaError_t launch_small_golay23(const SmallLaunch& l, cudaStream_t stream) {
    return small::launch_named<named::Golay23_X, named::Golay23_Z>(l, stream);
}
}  // namespace qcss
