// am_device.cuh -- device image of the automaton and the device helpers shared by the kernels.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

#include "am_internal.h"

namespace am {

enum ScanMode { MODE_COUNT = 0, MODE_ANY = 1, MODE_EMIT = 2 };

// Plain-pointer view of the automaton in HBM, passed to kernels by value.
struct DevAutomaton {
  const uint32_t* dense;        // dense_states x 256 failure-resolved next states (tagged)
  const uint32_t* fail;         // per state
  const EdgeSlot* edges;        // hashed goto
  const JumpSlot* jump;         // q-gram -> depth-q state (+ the tail of a simple sub-trie)
  const uint8_t* tails;         // tail bytes of the simple jump slots
  const uint32_t* filter;       // FILTER_WORDS words, bank-replicated q-gram bitmap
  const uint32_t* filter2;      // second-level table (shared-memory form)
  const uint32_t* gbits;        // second-level bitmap of large needle sets (global memory, q > 4)
  const uint32_t* own_off;      // CSR of needles ending exactly at a state
  const uint32_t* own_rank;
  const uint32_t* first_out;    // output chain heads / links (walk kernel)
  const uint32_t* next_out;
  const uint32_t* chain_count;
  const uint32_t* id_of_rank;
  const uint32_t* len_of_rank;
  const uint32_t* cdfa;         // class-compressed failure-resolved rows (walk kernel)
  const uint8_t* cls;           // byte -> class (256 entries)
  const uint16_t* lower1;       // two-stage Char.toLower table
  const int32_t* lower2;
  uint32_t dense_states, edge_mask, jump_mask;
  uint32_t q, qmask, min_len, max_len, rank_bits, num_states, num_needles;
  uint32_t ignore_case, halo;
  uint32_t t2_exact, t2_empty_key, gbits_shift;
  uint32_t cdfa_states, cdfa_shift;
};

struct ScanArgs {
  const uint8_t* text;          // device text, bytes [0, text_len)
  uint64_t text_len;
  uint64_t report_begin;        // report matches with end_pos in (report_begin, text_len]
  uint64_t pos_base;            // added to reported positions
  unsigned long long* d_count;  // COUNT: matches; EMIT: keys produced (may exceed cap)
  uint64_t* d_keys;             // EMIT: (pos << rank_bits | rank)
  uint64_t cap;
  // filter kernel, EMIT: keys go into per-segment slots of d_keys (segment = 2^seg_shift bytes of end positions, seg_cap
  // slots each, counters in seg_counts); keys of a full segment go to d_keys[ovf_base ..) and are counted in d_count
  uint32_t* seg_counts; uint32_t seg_shift, seg_cap; uint64_t ovf_base, ovf_cap;
  // filter kernel, IgnoreCase: 1 = `text` is the ORIGINAL text (one pass: folded probe, survivors lowered on the fly);
  // 0 = `text` is a lowered copy
  uint32_t ic_one_pass;
  // filter scan, list form: the positions that pass both filter levels are listed here and verified by verify_kernel.  Every
  // CTA of filter_kernel owns a region of surv_cap_cta entries and its own counter (148 atomics on ONE address per flush
  // round serialise in L2); a counter may exceed the capacity (the region then holds the first surv_cap_cta entries).
  // verify_kernel sums up: surv_count[0] = regions * max counter (the capacity a complete list would have needed; the host
  // repeats the scan with a larger list when it exceeds regions * surv_cap_cta), surv_count[1] = survivors in total (the
  // inline form adds its own there: the host's survivor-rate monitor).
  ulonglong2* surv; unsigned long long* surv_counts; unsigned long long* surv_count; uint64_t surv_cap_cta; uint32_t surv_regions;   // entry: {text index, the eight text bytes at it}
  uint32_t any_mode;            // containsAny: the kernels poll *d_flag and stop early
  uint32_t force_list;          // development (AM_FILTER_LIST=1): take the list form where the inline form would run
  int* d_flag;                  // ANY
  uint32_t debug;               // development only (AM_DEBUG_FLAGS): 1 = probes only, 2 = no deep verify
  uint32_t krow;                // bytes per filter row (4 * copies) as a run-time value: keeps the address an IMAD (FMA pipe)
};

// ---- small device helpers ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint4 ld_stream_v4(const uint4* p) {  // streaming 128-bit load, do not pollute L1
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// SWAR toLowerAscii (Utf8.hs:131-135) on the ASCII bytes of a word; bytes >= 0x80 pass through.
__device__ __forceinline__ uint32_t lower_ascii_word(uint32_t q) {
  const uint32_t q7 = q & 0x7f7f7f7fu;                       // bit 7 of (c + 0x3f) & ~(c + 0x25) marks 'A'..'Z'
  return q | ((((q7 + 0x3f3f3f3fu) & ~(q7 + 0x25252525u)) & ~q & 0x80808080u) >> 2);
}

__device__ __forceinline__ uint32_t lower_cp(const DevAutomaton& A, uint32_t cp) {  // Utf8.hs:145-151
  if (cp < 128u) return cp + ((cp - 'A' < 26u) ? 0x20u : 0u);
  if (cp >= 0x110000u) return cp;
  uint32_t blk = __ldg(A.lower1 + (cp >> LOWER_BLOCK_SHIFT));
  if (blk == 0) return cp;
  return (uint32_t)((int32_t)cp + __ldg(A.lower2 + blk * 128u + (cp & 127u)));
}

// goto(s, b) through the hashed edge table; NONE if the trie has no such edge.  Result is tagged.
__device__ __forceinline__ uint32_t edge_lookup(const DevAutomaton& A, uint32_t s, uint32_t b) {
  const uint32_t klo = (s << 8) | b, khi = s >> 24;
  uint32_t i = edge_hash(s, b) & A.edge_mask;
  for (;;) {
    uint4 e = __ldg(reinterpret_cast<const uint4*>(A.edges) + i);
    if (e.z == NONE) return NONE;
    if (e.x == klo && e.y == khi) return e.z;
    i = (i + 1) & A.edge_mask;
  }
}

// One Aho-Corasick step (goto + failure) on the byte automaton.  `s` untagged, result tagged.
__device__ __forceinline__ uint32_t ac_step(const DevAutomaton& A, uint32_t s, uint32_t b) {
  for (;;) {
    if (s < A.dense_states) return __ldg(A.dense + (size_t)s * 256u + b);
    uint32_t c = edge_lookup(A, s, b);
    if (c != NONE) return c;
    s = __ldg(A.fail + s);
  }
}

// CTA-level staging of match keys: shared-memory slots reserved with a shared atomic, flushed
// to HBM with ONE global atomic per flush (2 M same-address global atomics would serialise in L2).
template <int CAP>
struct KeyStage {
  unsigned long long keys[CAP];
  unsigned int n;
  unsigned long long base;

  __device__ __forceinline__ void init() { if (threadIdx.x == 0) n = 0; }
  __device__ __forceinline__ void push(const ScanArgs& a, unsigned long long key) {
    unsigned int i = atomicAdd(&n, 1u);
    if (i < CAP) { keys[i] = key; return; }
    unsigned long long g = atomicAdd(a.d_count, 1ull);  // stage full: slow but correct direct append
    if (g < a.cap) a.d_keys[g] = key;
  }
  // All threads of the CTA must call this (contains __syncthreads).
  __device__ __forceinline__ void flush(const ScanArgs& a) {
    __syncthreads();
    unsigned int cnt = n < (unsigned)CAP ? n : (unsigned)CAP;
    if (threadIdx.x == 0 && cnt) base = atomicAdd(a.d_count, (unsigned long long)cnt);
    __syncthreads();
    if (cnt) {
      unsigned long long b = base;
      for (unsigned int i = threadIdx.x; i < cnt; i += blockDim.x)
        if (b + i < a.cap) a.d_keys[b + i] = keys[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) n = 0;
    __syncthreads();
  }
};

}  // namespace am
