// filter_kernel, ANY mode (containsAny: a flag, other CTAs stop at their next tile).
#include "am_filter_impl.cuh"
namespace am {
cudaError_t launch_filter_any(const DevAutomaton& A, const ScanArgs& a, cudaStream_t st) { return launch_filter_mode<MODE_ANY>(A, a, st); }
}  // namespace am
