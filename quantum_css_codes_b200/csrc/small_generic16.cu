// Generic (runtime H) decode kernels for codes with n <= 16 qubits.
#include "small_common.cuh"

namespace qcss {
cudaError_t launch_small_generic16(const SmallLaunch& l, int mb, cudaStream_t stream) {
    if (mb == kSlicedM) return small::launch_generic<16, kSlicedM, 4>(l, stream);
    if (mb == 8) return small::launch_generic<16, 8, 4>(l, stream);
    return small::launch_generic<16, 16, 4>(l, stream);
}
}  // namespace qcss
