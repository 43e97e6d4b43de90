// filter_kernel, list form (IMODE = -1): IgnoreCase automata and long q-grams; verify_kernel follows (am_verify.cu).
#include "am_filter_impl.cuh"
namespace am {
template <bool FOLD>
static cudaError_t launch_list_c(const DevAutomaton& A, const ScanArgs& a, cudaStream_t st) {
  const bool x = A.t2_exact != 0;
  switch (A.q) {
    case 8: return x ? cudaErrorInvalidValue : launch_filter_t<8, 2, FOLD, -1>(A, a, st);
    case 6: return x ? cudaErrorInvalidValue : launch_filter_t<6, 2, FOLD, -1>(A, a, st);
    case 4: return x ? launch_filter_t<4, 1, FOLD, -1>(A, a, st) : launch_filter_t<4, 0, FOLD, -1>(A, a, st);
    case 1: case 2: case 3: return x ? launch_filter_t<0, 1, FOLD, -1>(A, a, st) : launch_filter_t<0, 0, FOLD, -1>(A, a, st);
    default: return cudaErrorInvalidValue;
  }
}
cudaError_t launch_filter_list(const DevAutomaton& A, const ScanArgs& a, cudaStream_t st) {
  return A.ignore_case ? launch_list_c<true>(A, a, st) : launch_list_c<false>(A, a, st);
}
int filter_kernel_smem_bytes() { return (int)sizeof(FilterSmem); }
}  // namespace am
