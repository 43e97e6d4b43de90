// filter_kernel, EMIT mode (all matches: keys into per-segment slots).
#include "am_filter_impl.cuh"
namespace am {
cudaError_t launch_filter_emit(const DevAutomaton& A, const ScanArgs& a, cudaStream_t st) { return launch_filter_mode<MODE_EMIT>(A, a, st); }
}  // namespace am
