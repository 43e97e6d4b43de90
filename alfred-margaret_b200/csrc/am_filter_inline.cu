// filter_kernel, inline form (IMODE = the scan mode): CaseSensitive automata with q <= 4 verify their survivors in the kernel.
#include "am_filter_impl.cuh"
namespace am {
cudaError_t launch_filter_list(const DevAutomaton& A, const ScanArgs& a, cudaStream_t st);
template <int IMODE>
static cudaError_t launch_inline_m(const DevAutomaton& A, const ScanArgs& a, cudaStream_t st) {
  const bool x = A.t2_exact != 0;
  if (A.q == 4) return x ? launch_filter_t<4, 1, false, IMODE>(A, a, st) : launch_filter_t<4, 0, false, IMODE>(A, a, st);
  return x ? launch_filter_t<0, 1, false, IMODE>(A, a, st) : launch_filter_t<0, 0, false, IMODE>(A, a, st);
}
// The filter scan.  List form: filter_kernel lists the survivors, verify_kernel turns them into matches.
cudaError_t launch_filter(const DevAutomaton& A, const ScanArgs& a, int mode, cudaStream_t st) {
  if (a.text_len <= a.report_begin) return cudaSuccess;
  if (!A.ignore_case && A.q >= 1 && A.q <= 4 && !a.force_list) {
    if (mode == MODE_COUNT) return launch_inline_m<MODE_COUNT>(A, a, st);
    if (mode == MODE_ANY) return launch_inline_m<MODE_ANY>(A, a, st);
    return launch_inline_m<MODE_EMIT>(A, a, st);
  }
  cudaError_t e = launch_filter_list(A, a, st);
  if (e != cudaSuccess) return e;
  return launch_verify(A, a, mode, st);
}
}  // namespace am
