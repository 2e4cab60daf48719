// Static (compile-time H, L, truth tables) decode kernels for the qrm15 descriptor.
#include "named_codes.inc"
#include "small_common.cuh"

namespace qcss {
cudaError_t launch_small_qrm15(const SmallLaunch& l, cudaStream_t stream) {
    return small::launch_named<named::Qrm15_X, named::Qrm15_Z>(l, stream);
}
}  // namespace qcss
