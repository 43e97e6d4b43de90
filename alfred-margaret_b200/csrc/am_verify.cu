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

namespace am {

// Emit one match.  The key goes into the SEGMENT of its end position (128 KiB of text per segment, a fixed number of
// slots each; slot reserved with an atomic on the segment's own counter), so the list comes out ordered at segment
// granularity and a local rank sort per segment replaces the global radix sort (seg_sort_kernel).  A key whose
// segment is full goes to the overflow area; any overflow sends the host down the compact + radix sort path.
__device__ __forceinline__ void fk_emit(const DevAutomaton& A, const ScanArgs& a, uint64_t end, uint32_t rank) {
  const unsigned long long key = ((unsigned long long)(end + a.pos_base) << A.rank_bits) | rank;
  const uint32_t seg = (uint32_t)((end - a.report_begin - 1) >> a.seg_shift);
  const uint32_t slot = atomicAdd(a.seg_counts + seg, 1u);
  if (slot < a.seg_cap) { a.d_keys[(uint64_t)seg * a.seg_cap + slot] = key; return; }
  const unsigned long long o = atomicAdd(a.d_count, 1ull);   // EMIT: d_count counts the overflowed keys
  if (o < a.ovf_cap) a.d_keys[a.ovf_base + o] = key;
}

// ---- the text as the verification must see it -------------------------------------------------------------------------
// CaseSensitive (and IgnoreCase on a lowered copy): the bytes themselves.  IgnoreCase in one pass over the ORIGINAL text:
// `runLower` lower-cases every code point of the haystack (consumeInput, Automaton.hs:468-480; lowerCodePoint,
// Utf8.hs:145-151), so a survivor is verified on a stream that decodes (decodeN, Utf8.hs:344-350), lowers and re-encodes
// code point by code point.  A code point whose lower case has another UTF-8 length is passed through unchanged: the
// automaton holds the needle variants that match it (am_build.cpp step 1), so byte offsets in the stream are byte
// offsets in the text.
// Four text bytes from an arbitrary address as one word (two aligned loads + funnel shift); the caller guarantees that
// the eight bytes of the two aligned words lie in memory the kernel may read (p + 8 <= end of the text).
__device__ __forceinline__ uint32_t fk_load32u(const uint8_t* p) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~uintptr_t(3));
  return __funnelshift_r(__ldg(w), __ldg(w + 1), (uint32_t)(a & 3) * 8u);
}

template <bool LOWER>
struct TextStream;
template <>
struct TextStream<false> {
  const uint8_t* p; const uint8_t* end;
  __device__ __forceinline__ TextStream(const DevAutomaton&, const uint8_t* b, const uint8_t* e) : p(b), end(e) {}
  __device__ __forceinline__ int next() { return p < end ? (int)__ldg(p++) : -1; }
  __device__ __forceinline__ bool next4_ascii(uint32_t* out) {                // four bytes at once (no lowering here: any bytes)
    if (p + 8 > end) return false;
    *out = fk_load32u(p);
    p += 4;
    return true;
  }
};
// Lower one code point above ASCII of `n` bytes whose raw bytes are `raw` (byte k at bits 8 k): decode, `Char.toLower` table,
// re-encode; unchanged when its lower case has another UTF-8 length (the needle variants match it).  
static __device__ __forceinline__ uint32_t fk_lower_multibyte(
const DevAutomaton& A, uint32_t raw, uint32_t n) {
  const uint32_t c0 = raw & 0xFFu, c1 = (raw >> 8) & 0xFFu, c2 = (raw >> 16) & 0xFFu, c3 = raw >> 24;
  const uint32_t cp = n == 2 ? ((c0 & 0x1Fu) << 6) | (c1 & 0x3Fu)
                    : n == 3 ? ((c0 & 0x0Fu) << 12) | ((c1 & 0x3Fu) << 6) | (c2 & 0x3Fu)
                             : ((c0 & 0x07u) << 18) | ((c1 & 0x3Fu) << 12) | ((c2 & 0x3Fu) << 6) | (c3 & 0x3Fu);
  const uint32_t l = lower_cp(A, cp);
  const uint32_t ln = l < 0x80u ? 1u : l < 0x800u ? 2u : l < 0x10000u ? 3u : 4u;
  if (l == cp || ln != n) return raw;                         // unchanged, or kept because its lower case has another length
  if (n == 2) return (0xC0u | (l >> 6)) | ((0x80u | (l & 0x3Fu)) << 8);
  if (n == 3) return (0xE0u | (l >> 12)) | ((0x80u | ((l >> 6) & 0x3Fu)) << 8) | ((0x80u | (l & 0x3Fu)) << 16);
  return (0xF0u | (l >> 18)) | ((0x80u | ((l >> 12) & 0x3Fu)) << 8) | ((0x80u | ((l >> 6) & 0x3Fu)) << 16) | ((0x80u | (l & 0x3Fu)) << 24);
}

template <>
struct TextStream<true> {
  const DevAutomaton& A; const uint8_t* p; const uint8_t* end; uint32_t pend, npend;
  __device__ __forceinline__ TextStream(const DevAutomaton& A_, const uint8_t* b, const uint8_t* e) : A(A_), p(b), end(e), pend(0), npend(0) {}
  __device__ __forceinline__ int next() {
    if (npend == 0) {
      if (p >= end) return -1;
      const uint32_t c0 = __ldg(p);
      const uint32_t n = c0 < 0xC0u ? 1u : c0 < 0xE0u ? 2u : c0 < 0xF0u ? 3u : 4u;
      if (p + n > end) return -1;
      if (n == 1) {
        pend = c0 + ((c0 - 'A' < 26u) ? 0x20u : 0u);          // toLowerAscii (Utf8.hs:131-135)
      } else {
        const uint32_t c1 = __ldg(p + 1), c2 = n > 2 ? __ldg(p + 2) : 0u, c3 = n > 3 ? __ldg(p + 3) : 0u;
        pend = fk_lower_multibyte(A, c0 | c1 << 8 | c2 << 16 | c3 << 24, n);
      }
      npend = n; p += n;
    }
    const uint32_t b = pend & 0xFFu;
    pend >>= 8; npend--;
    return (int)b;
  }
  // Fast lane for ASCII text: the next four bytes lowered at once (toLowerAscii as SWAR), or false when the stream is inside a
  // code point, within 8 bytes of the end, or the four bytes hold one above ASCII -- then next() goes byte by byte.
  __device__ __forceinline__ bool next4_ascii(uint32_t* out) {
    if (npend != 0 || p + 8 > end) return false;
    const uint32_t w = fk_load32u(p);
    if (w & 0x80808080u) return false;
    *out = lower_ascii_word(w);
    p += 4;
    return true;
  }
};

template <int MODE>
__device__ __forceinline__ void fk_report_state(const DevAutomaton& A, const ScanArgs& a, uint32_t s, uint64_t end, unsigned long long& local_count) {
  const uint32_t olo = __ldg(A.own_off + s), ohi = __ldg(A.own_off + s + 1);   // all needles that end at this state (duplicates)
  if (MODE == MODE_COUNT) local_count += ohi - olo;
  else
    for (uint32_t j = olo; j < ohi; j++) fk_emit(A, a, end, __ldg(A.own_rank + j));
}

// Verify a survivor (its q-gram passed both filter levels): report every needle that is a prefix of the (lowered)
// text at i.  No failure links are needed because every start position is tried (failure-less, position-parallel
// formulation of Aho-Corasick).  The jump table maps the q-gram to its trie state -- or, when a single needle path
// hangs below it (nearly always), to that path's tail, which is compared with the text in one go.
template <int MODE, bool LOWER>
__device__ __forceinline__ void fk_deep_verify(const DevAutomaton& A, const ScanArgs& a, uint64_t i, uint32_t b_lo, uint32_t b_hi, unsigned long long& local_count) {
  // i: text index of the survivor; b_lo, b_hi: the eight text bytes there, carried from the scan kernel's window (bytes beyond
  // the text are arbitrary: every use below is bounded by text_len)
  if (i + A.min_len > a.text_len) return;
  const uint8_t* tp = a.text + i;
  const uint32_t q = A.q;
  uint32_t g_lo, g_hi = 0;
  uint32_t have = 8;                                          // text bytes [0, have) of the survivor are in (b_lo, b_hi), lowered if LOWER
  if (LOWER) {
    // `runLower` lowers every code point.  ASCII bytes are lowered in place (toLowerAscii as SWAR); a survivor with a byte above
    // ASCII among its first eight goes through the code point stream from its first byte.
    if ((b_lo & 0xC0u) == 0x80u) return;                      // inside a code point: no needle starts here
    if (((b_lo | b_hi) & 0x80808080u) == 0) { b_lo = lower_ascii_word(b_lo); b_hi = lower_ascii_word(b_hi); }
    else have = 0;
  }
  if (i + 8 > a.text_len) have = 0;                           // (the carried bytes run past the text: read it byte by byte instead)
  TextStream<LOWER> ts(A, tp + have, a.text + a.text_len);
  if (have) {
    g_lo = q < 4 ? b_lo & A.qmask : b_lo;
    if (q > 4) g_hi = q == 6 ? b_hi & 0xFFFFu : b_hi;
  } else {
    g_lo = 0;
    for (uint32_t k = 0; k < q; k++) {
      const int b = ts.next();
      if (b < 0) return;
      if (k < 4) g_lo |= (uint32_t)b << (8 * k); else g_hi |= (uint32_t)b << (8 * (k - 4));
    }
  }
  // bytes of the text after the q-gram come from the carried eight first, then from the stream
  uint32_t at = q;                                            // next text byte to compare (offset from i)
  auto next_byte = [&]() -> int {
    if (at < have) { const uint32_t b = at < 4 ? (b_lo >> (8 * at)) & 0xFFu : (b_hi >> (8 * (at - 4))) & 0xFFu; at++; return (int)b; }
    at++;
    return ts.next();
  };
  const uint32_t nhead = q > 4 ? q - 4 : 0;                   // bytes of the q-gram that head the tail of a slot
  uint32_t idx = jump_hash(g_lo, g_hi) & A.jump_mask;
  uint32_t st;
  for (;;) {
    const uint4 s = __ldg(reinterpret_cast<const uint4*>(A.jump) + idx);
    if (s.y == NONE) return;
    if (s.x == g_lo) {
#if FK_TAIL
      if (s.w & JUMP_SIMPLE) {
        const uint32_t tl = s.w & JUMP_TAIL_MASK;
        const uint8_t* np = A.tails + s.z;
        bool mine = true;
        for (uint32_t j = 0; j < nhead; j++) mine = mine && (uint32_t)__ldg(np + j) == ((g_hi >> (8 * j)) & 0xFFu);
        if (mine) {
          // one needle path below this q-gram: compare the rest of its tail with the text
          const uint32_t rest = tl - nhead;
          const uint64_t end = i + q + rest;
          if (end > a.text_len || end <= a.report_begin) return;
          uint32_t k = nhead;
          if (have == 8 && q == 4 && tl >= 4) {               // the first four tail bytes against the carried bytes 4..7, at once
            if (__ldg(reinterpret_cast<const uint32_t*>(np)) != b_hi) return;
            k = 4; at = 8;
          }
          while (k < tl) {
            uint32_t w4;
            if (at >= have && k + 4 <= tl && ((k & 3u) == 0) && ts.next4_ascii(&w4)) {   // (tails are 4-byte aligned in `tails`)
              if (w4 != __ldg(reinterpret_cast<const uint32_t*>(np + k))) return;
              k += 4; at += 4;
            } else {
              if (next_byte() != (int)__ldg(np + k)) return;
              k++;
            }
          }
          if (MODE == MODE_ANY) { *a.d_flag = 1; return; }
          if (s.w & JUMP_SINGLE) {
            if (MODE == MODE_COUNT) local_count += 1;
            else fk_emit(A, a, end, s.y);
          } else {
            fk_report_state<MODE>(A, a, s.y, end, local_count);
          }
          return;
        }
      } else
#endif
      if (s.z == g_hi) { st = s.y; break; }                  // not simple: the slot holds the depth-q state
    }
    idx = (idx + 1) & A.jump_mask;                            // (another q-gram, possibly one with the same first four bytes)
  }
  uint32_t d = q;
  for (;;) {
    if (st & OWN_FLAG) {
      const uint64_t end = i + d;
      if (end > a.report_begin) {
        if (MODE == MODE_ANY) { *a.d_flag = 1; return; }
        fk_report_state<MODE>(A, a, st & ID_MASK, end, local_count);
      }
    }
    const int ch = next_byte();
    if (ch < 0) return;
    st = edge_lookup(A, st & ID_MASK, (uint32_t)ch);
    if (st == NONE) return;
    d++;
  }
}


template <int MODE, bool LOWER>
__global__ void __launch_bounds__(256) verify_kernel(const __grid_constant__ DevAutomaton A, const __grid_constant__ ScanArgs a) {
  __shared__ unsigned long long red[8];
  unsigned long long n = *a.surv_count;
  if (n > a.surv_cap) n = a.surv_cap;                          // (the host sees the overflow in surv_count and repeats the scan with a larger list)
  unsigned long long local_count = 0;
  for (unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (unsigned long long)gridDim.x * blockDim.x) {
    if (MODE == MODE_ANY && *reinterpret_cast<volatile int*>(a.d_flag)) break;
    const ulonglong2 e = a.surv[k];
    fk_deep_verify<MODE, LOWER>(A, a, e.x, (uint32_t)e.y, (uint32_t)(e.y >> 32), local_count);
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
