// am_kernels.h -- host-callable launchers of the scan kernels (am_kernels.cu).
#pragma once

#include <cuda_runtime.h>

#include "am_device.cuh"

#include <atomic>

namespace am {

extern std::atomic<uint64_t> g_kernel_launches;   // kernels launched by this library (bench.py's gpu_launches)

// SM count of the current device (grids are sized as one persistent CTA, or two, per SM).
inline int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0; cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// Per-segment goto+failure walk (general path).  mode: ScanMode.
cudaError_t launch_walk(const DevAutomaton& A, const ScanArgs& a, int mode, cudaStream_t st);
// q-gram filter + goto verify (fast path; requires A.q > 0 and CaseSensitive).
cudaError_t launch_filter(const DevAutomaton& A, const ScanArgs& a, int mode, cudaStream_t st);

// IgnoreCase front end of the filter kernel: lowered copy of the text (+ count of length-changing code points).
cudaError_t launch_lower(const DevAutomaton& A, const uint8_t* text, uint64_t text_len, uint8_t* out, unsigned int* exceptions, bool keep, cudaStream_t st);
size_t sort_temp_bytes(uint64_t n, int end_bit);
cudaError_t sort_keys(void* temp, size_t temp_bytes, const uint64_t* in, uint64_t* out, uint64_t n, int end_bit, cudaStream_t st);
cudaError_t launch_unpack(const uint64_t* keys, uint64_t n, uint32_t rank_bits, const uint32_t* id_of_rank, am_match* out, cudaStream_t st);
int filter_kernel_smem_bytes();

// synthetic workload generator (am_synth.cu)
cudaError_t launch_synth_fill(uint8_t* buf, uint64_t len, uint64_t first, uint64_t seed, const uint8_t* d_alpha, uint32_t alpha_len, cudaStream_t st);
cudaError_t launch_synth_plant(uint8_t* buf, uint64_t len, uint64_t first, uint64_t seed, const uint8_t* d_needle_bytes,
                               const uint32_t* d_needle_off, uint32_t n, uint32_t block, cudaStream_t st);

}  // namespace am
