// am_verify.cu -- verification of the q-gram filter's survivors (second kernel of the filter scan).
//
// filter_kernel (am_filter_impl.cuh) only FILTERS: the positions that pass its two levels -- "survivors", a fraction of a
// per cent of the text -- are appended to a list in global memory.  This kernel verifies them, one survivor per thread:
// thousands of independent goto-trie walks in flight hide the latency of their dependent lookups (jump table -> tail
// bytes -> text), which the scan warps used to eat one survivor batch at a time; and the scan kernel's hot loop no
// longer carries the verification code (five inlined copies: > 64 KB of SASS) through its instruction cache.
#include <cstddef>

#include "am_device.cuh"
#include "am_kernels.h"
#include "am_verify.cuh"

namespace am {

template <int MODE, bool LOWER>
__global__ void __launch_bounds__(256) verify_kernel(const __grid_constant__ DevAutomaton A, const __grid_constant__ ScanArgs a) {
  __shared__ unsigned long long red[8];
  // Block b works on region b mod R -- the survivors filter_kernel's CTA (b mod R) listed -- together with the other blocks of
  // that region: no prefix over the regions, no search for "the region that holds survivor k".  The regions hold similar
  // numbers of survivors (every CTA of the scan reads the same share of the text).
  const uint32_t R = a.surv_regions;
  const uint32_t region = blockIdx.x % R, part = blockIdx.x / R, parts = (gridDim.x - region + R - 1) / R;   // blocks region, region + R, ... share it
  const unsigned long long listed = a.surv_counts[region];
  const unsigned long long n = listed < a.surv_cap_cta ? listed : a.surv_cap_cta;   // (a region that overflowed holds its first surv_cap_cta survivors)
  if (blockIdx.x == 0 && threadIdx.x < 32) {                   // the totals the host reads: see ScanArgs::surv_count
    unsigned long long mx = 0, total = 0;
    for (uint32_t r = threadIdx.x; r < R; r += 32) { const unsigned long long c = a.surv_counts[r]; total += c; mx = c > mx ? c : mx; }
    for (int o = 16; o > 0; o >>= 1) {
      total += __shfl_down_sync(0xFFFFFFFFu, total, o);
      const unsigned long long m2 = __shfl_down_sync(0xFFFFFFFFu, mx, o); mx = m2 > mx ? m2 : mx;
    }
    if (threadIdx.x == 0) { a.surv_count[0] = mx * R; atomicAdd(a.surv_count + 1, total); }
  }
  const uint32_t a0 = (uint32_t)(reinterpret_cast<uintptr_t>(a.text) & 15);
  const ulonglong2* list = a.surv + (uint64_t)region * a.surv_cap_cta;
  unsigned long long local_count = 0;
  if (fk_has_level3(A)) {
    // Images with a third filter level (C3): nine survivors in ten end at that check, and the rest -- scattered over the lanes --
    // would run the divergent verification three lanes to a warp.  So a batch is CHECKED one survivor per thread, those that
    // pass are queued in shared memory, and a full block's worth of them is VERIFIED one per thread (all lanes busy).
    __shared__ ulonglong2 q[2 * 256];
    __shared__ uint32_t q_n;
    const uint32_t tid = threadIdx.x;
    if (tid == 0) q_n = 0;
    __syncthreads();
    for (unsigned long long base = (unsigned long long)part * 256; base < n; base += (unsigned long long)parts * 256) {   // (uniform over the block)
      if (MODE == MODE_ANY && __syncthreads_or(*reinterpret_cast<volatile int*>(a.d_flag))) break;
      const unsigned long long k = base + tid;
      if (k < n) {
        const ulonglong2 e = list[k];                            // {virtual index = a0 + text index, eight text bytes}
        if (e.x >= a0 && fk_level3_pass(A, a, e.x - a0, (uint32_t)e.y, (uint32_t)(e.y >> 32))) q[atomicAdd(&q_n, 1u)] = e;
      }
      __syncthreads();
      const uint32_t waiting = q_n;
      __syncthreads();                                           // (everyone has read the count before the next batch adds to it)
      if (waiting >= 256) {
        const ulonglong2 e = q[tid];
        fk_deep_verify<MODE, LOWER, false>(A, a, e.x - a0, (uint32_t)e.y, (uint32_t)(e.y >> 32), local_count);
        const bool more = 256 + tid < waiting;
        ulonglong2 mv = make_ulonglong2(0, 0);
        if (more) mv = q[256 + tid];
        __syncthreads();
        if (more) q[tid] = mv;
        if (tid == 0) q_n = waiting - 256;
        __syncthreads();
      }
    }
    __syncthreads();
    if (tid < q_n) {
      const ulonglong2 e = q[tid];
      fk_deep_verify<MODE, LOWER, false>(A, a, e.x - a0, (uint32_t)e.y, (uint32_t)(e.y >> 32), local_count);
    }
  } else {
    for (unsigned long long k = (unsigned long long)part * blockDim.x + threadIdx.x; k < n; k += (unsigned long long)parts * blockDim.x) {
      if (MODE == MODE_ANY && *reinterpret_cast<volatile int*>(a.d_flag)) break;
      const ulonglong2 e = list[k];                              // {virtual index = a0 + text index, eight text bytes}
      if (e.x < a0) continue;                                    // bytes of the first granule that precede the text
      fk_deep_verify<MODE, LOWER, false>(A, a, e.x - a0, (uint32_t)e.y, (uint32_t)(e.y >> 32), local_count);
    }
  }
  if (MODE == MODE_COUNT) {
    for (int o = 16; o > 0; o >>= 1) local_count += __shfl_down_sync(0xFFFFFFFFu, local_count, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local_count;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long s = 0;
      for (int i = 0; i < 8; i++) s += red[i];
      if (s) atomicAdd(a.d_count, s);
    }
  }
}

template <int MODE>
static cudaError_t launch_verify_m(const DevAutomaton& A, const ScanArgs& a, cudaStream_t st) {
  const unsigned blocks = (unsigned)sm_count() * 8;            // 2048 threads per SM; the loop bound lives on the device
  g_kernel_launches++;
  if (A.ignore_case && a.ic_one_pass) verify_kernel<MODE, true><<<blocks, 256, 0, st>>>(A, a);
  else verify_kernel<MODE, false><<<blocks, 256, 0, st>>>(A, a);
  return cudaGetLastError();
}

cudaError_t launch_verify(const DevAutomaton& A, const ScanArgs& a, int mode, cudaStream_t st) {
  if (a.text_len <= a.report_begin) return cudaSuccess;
  if (mode == MODE_COUNT) return launch_verify_m<MODE_COUNT>(A, a, st);
  if (mode == MODE_ANY) return launch_verify_m<MODE_ANY>(A, a, st);
  return launch_verify_m<MODE_EMIT>(A, a, st);
}

}  // namespace am
