// Generic (runtime H) decode kernels for codes with n <= 32 qubits.
#include "small_common.cuh"

namespace qcss {
cudaError_t launch_small_generic32(const SmallLaunch& l, int mb, cudaStream_t stream) {
    if (mb == kSlicedM) return small::launch_generic<32, kSlicedM, 2>(l, stream);
    if (mb == 8) return small::launch_generic<32, 8, 2>(l, stream);
    return small::launch_generic<32, 16, 2>(l, stream);
}
}  // namespace qcss
