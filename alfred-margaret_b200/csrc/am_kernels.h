// am_kernels.h -- host-callable launchers of the scan kernels (am_kernels.cu).
#pragma once

#include <cuda_runtime.h>

#include "am_device.cuh"

#include <atomic>

namespace am {

extern std::atomic<uint64_t> g_kernel_launches;   // kernels launched by this library (bench.py's gpu_launches)

// SM count of the current device (grids are sized as one persistent CTA, or two, per SM).  Cached per device: one
// process may drive several GPUs (am_options.device).
inline int sm_count() {
  static std::atomic<int> cache[64];
  int dev = 0; cudaGetDevice(&dev);
  const int slot = dev & 63;
  int n = cache[slot].load(std::memory_order_relaxed);
  if (n <= 0) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    cache[slot].store(n, std::memory_order_relaxed);
  }
  return n;
}
// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE (per-context) attribute of a kernel: opt in once per
// (kernel, device).  `done` is one mask per kernel instantiation (a function-local static of the caller), bit = device.
template <class K>
inline cudaError_t ensure_dynamic_smem(K kernel, int bytes, std::atomic<uint64_t>& done) {
  int dev = 0; cudaGetDevice(&dev);
  const uint64_t bit = 1ull << (dev & 63);
  if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
  return e;
}

// Per-segment goto+failure walk (general path).  mode: ScanMode.
cudaError_t launch_walk(const DevAutomaton& A, const ScanArgs& a, int mode, cudaStream_t st);
// q-gram filter scan (fast path; requires A.q > 0): filter_kernel lists the survivors, verify_kernel verifies them.
cudaError_t launch_filter(const DevAutomaton& A, const ScanArgs& a, int mode, cudaStream_t st);
cudaError_t launch_verify(const DevAutomaton& A, const ScanArgs& a, int mode, cudaStream_t st);

// IgnoreCase front end of the filter kernel: lowered copy of the text (+ count of length-changing code points).
cudaError_t launch_lower(const DevAutomaton& A, const uint8_t* text, uint64_t text_len, uint8_t* out, unsigned int* exceptions, bool keep, cudaStream_t st);
// Segmented emission of the filter kernel (ScanArgs::seg_*): bases = exclusive prefix of min(count, seg_cap), bases[num_segs] = their sum;
// then either every segment is rank-sorted into its final place (no key overflowed), or everything is compacted for the radix sort.
constexpr uint32_t SEG_SHIFT = 17;            // 128 KiB of end positions per segment
constexpr uint32_t SEG_CAP_MAX = 256;         // slots per segment at most (the local sort is quadratic in the fill)
size_t seg_scan_temp_bytes(uint64_t num_segs);
cudaError_t launch_seg_scan(void* temp, size_t temp_bytes, const uint32_t* seg_counts /* num_segs + 1, last = 0 */, uint64_t num_segs, uint32_t seg_cap,
                            uint64_t* bases /* num_segs + 1, last = total */, cudaStream_t st);
cudaError_t launch_seg_sort(const uint64_t* seg_keys, const uint32_t* seg_counts, const uint64_t* bases, uint64_t num_segs, uint32_t seg_cap, uint64_t* out,
                            am_match* matches /* nullable: also unpack into am_match records */, uint64_t matches_cap, uint32_t rank_bits, const uint32_t* id_of_rank, cudaStream_t st);
cudaError_t launch_seg_compact(const uint64_t* seg_keys, const uint32_t* seg_counts, const uint64_t* bases, uint64_t num_segs, uint32_t seg_cap,
                               const uint64_t* ovf_keys, uint64_t n_ovf, uint64_t stored_total, uint64_t* out, cudaStream_t st);
size_t sort_temp_bytes(uint64_t n, int end_bit);
cudaError_t sort_keys(void* temp, size_t temp_bytes, const uint64_t* in, uint64_t* out, uint64_t n, int end_bit, cudaStream_t st);
cudaError_t launch_unpack(const uint64_t* keys, uint64_t n, uint32_t rank_bits, const uint32_t* id_of_rank, am_match* out, cudaStream_t st);
int filter_kernel_smem_bytes();
// *d_out = d_scalars[first] + ... + d_scalars[first + n - 1], bit 63 raised when the survivor list overflowed (one thread; joins the
// counters of a scan on the stream, ahead of the sharded calls' all-gather)
cudaError_t launch_shard_total(const unsigned long long* d_scalars, int first, int n, const unsigned long long* surv_count, unsigned long long surv_cap,
                               unsigned long long* d_out, cudaStream_t st);
// containsAll: OR the needle ranks of keys[0..n) into the bit set `seen`, then *d_missing = needles not seen yet
cudaError_t launch_mark_seen(const uint64_t* keys, uint64_t n, uint32_t rank_bits, uint32_t* seen, uint32_t num_needles, unsigned int* d_missing, cudaStream_t st);

// synthetic workload generator (am_synth.cu)
cudaError_t launch_synth_fill(uint8_t* buf, uint64_t len, uint64_t first, uint64_t seed, const uint8_t* d_alpha, uint32_t alpha_len, cudaStream_t st);
cudaError_t launch_synth_plant(uint8_t* buf, uint64_t len, uint64_t first, uint64_t seed, const uint8_t* d_needle_bytes,
                               const uint32_t* d_needle_off, uint32_t n, uint32_t block, cudaStream_t st);

}  // namespace am
