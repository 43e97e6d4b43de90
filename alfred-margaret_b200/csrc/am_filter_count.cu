// filter_kernel, COUNT mode (+ the mode dispatch of launch_filter).
#include "am_filter_impl.cuh"
namespace am {
cudaError_t launch_filter_any(const DevAutomaton& A, const ScanArgs& a, cudaStream_t st);
cudaError_t launch_filter_emit(const DevAutomaton& A, const ScanArgs& a, cudaStream_t st);
cudaError_t launch_filter(const DevAutomaton& A, const ScanArgs& a, int mode, cudaStream_t st) {
  if (mode == MODE_COUNT) return launch_filter_mode<MODE_COUNT>(A, a, st);
  if (mode == MODE_ANY) return launch_filter_any(A, a, st);
  return launch_filter_emit(A, a, st);
}
int filter_kernel_smem_bytes() { return (int)sizeof(FilterSmem); }
}  // namespace am
