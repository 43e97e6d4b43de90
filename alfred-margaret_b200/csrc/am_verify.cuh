// am_verify.cuh -- verification of a survivor of the q-gram filter (device functions shared by the verify kernel, am_verify.cu,
// and by the filter kernels that verify in line, am_filter_inline.cu).
#pragma once

#include "am_device.cuh"

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
  __device__ __forceinline__ TextStream(const DevAutomaton&, const uint8_t* b, const uint8_t* e, unsigned long long = 0, uint32_t = 0) : p(b), end(e) {}
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
  // Raw text bytes wait in a 64-bit register window (`buf`, nbuf bytes, the next byte lowest), seeded with the eight bytes the
  // survivor carries and refilled four at a time: the stream loads a word per four bytes instead of a byte per step, and
  // none at all for a survivor whose verification ends within its first eight bytes.  Lowered code points leave it through
  // `pend` (npend bytes).
  const DevAutomaton& A; const uint8_t* p; const uint8_t* end; unsigned long long buf; uint32_t nbuf, pend, npend;
  // b: address of the first byte that is NOT in the window; the window starts with the `n0` bytes of `seed`
  __device__ __forceinline__ TextStream(const DevAutomaton& A_, const uint8_t* b, const uint8_t* e, unsigned long long seed = 0, uint32_t n0 = 0)
      : A(A_), p(b), end(e), buf(seed), nbuf(n0), pend(0), npend(0) {}
  __device__ __forceinline__ void refill() {                 // at most 4 bytes in the window on entry
    if (p + 8 <= end) { buf |= (unsigned long long)fk_load32u(p) << (8 * nbuf); nbuf += 4; p += 4; }
    else if (p < end) { buf |= (unsigned long long)__ldg(p) << (8 * nbuf); nbuf += 1; p += 1; }
  }
  __device__ __forceinline__ int next() {
    if (npend == 0) {
      if (nbuf == 0) { refill(); if (nbuf == 0) return -1; }
      const uint32_t c0 = (uint32_t)buf & 0xFFu;
      const uint32_t n = c0 < 0xC0u ? 1u : c0 < 0xE0u ? 2u : c0 < 0xF0u ? 3u : 4u;
      while (nbuf < n) { const uint32_t before = nbuf; refill(); if (nbuf == before) return -1; }   // (a code point cut off by the end of the text)
      if (n == 1) pend = c0 + ((c0 - 'A' < 26u) ? 0x20u : 0u);          // toLowerAscii (Utf8.hs:131-135)
      else pend = fk_lower_multibyte(A, (uint32_t)buf & (0xFFFFFFFFu >> (32 - 8 * n)), n);
      npend = n; buf >>= 8 * n; nbuf -= n;
    }
    const uint32_t b = pend & 0xFFu;
    pend >>= 8; npend--;
    return (int)b;
  }
  // Fast lane for ASCII text: the next four bytes lowered at once (toLowerAscii as SWAR), or false when the stream is inside a
  // code point, at the end of the text, or the four bytes hold one above ASCII -- then next() goes byte by byte.
  __device__ __forceinline__ bool next4_ascii(uint32_t* out) {
    if (npend != 0) return false;
    if (nbuf < 4) { refill(); if (nbuf < 4) return false; }
    const uint32_t w = (uint32_t)buf;
    if (w & 0x80808080u) return false;
    *out = lower_ascii_word(w);
    buf >>= 32; nbuf -= 4;
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
// Third filter level (gp_hash, am_build.cpp; images with q = 4 and the bitmap second level): is one of the text's (folded) prefixes of
// 4 .. 8 bytes at i a closed needle resp. the 8-byte prefix of a longer one?  Five independent loads from an L2-resident bitmap,
// before anything is decoded or lowered.  b_lo, b_hi: the eight text bytes at i (as they stand in the text).
__device__ __forceinline__ bool fk_has_level3(const DevAutomaton& A) { return A.q == 4 && !A.t2_exact && A.gbits_shift < 32; }
__device__ __forceinline__ bool fk_level3_pass(const DevAutomaton& A, const ScanArgs& a, uint64_t i, uint32_t b_lo, uint32_t b_hi) {
  if (!fk_has_level3(A) || i + 8 > a.text_len) return true;   // (carried bytes that run past the text decide nothing)
  uint32_t f_lo = b_lo, f_hi = b_hi;
  if (A.ignore_case) { f_lo |= FOLD_MASK; f_hi |= FOLD_MASK; }
  uint32_t bit[5], word[5];
#pragma unroll
  for (uint32_t L = 4; L <= 8; L++) {
    const uint32_t h = L == 4 ? 0u : L == 8 ? f_hi : f_hi & ((1u << (8 * (L - 4))) - 1u);
    bit[L - 4] = gp_hash(f_lo, h, L) >> A.gbits_shift;
    word[L - 4] = __ldg(A.gbits + (bit[L - 4] >> 5));
  }
  uint32_t any = 0;
#pragma unroll
  for (int k = 0; k < 5; k++) any |= word[k] >> (bit[k] & 31);
  return (any & 1u) != 0;
}

template <int MODE, bool LOWER, bool CHECK_L3 = true>
__device__ __forceinline__ void fk_deep_verify(const DevAutomaton& A, const ScanArgs& a, uint64_t i, uint32_t b_lo, uint32_t b_hi, unsigned long long& local_count) {
  // i: text index of the survivor; b_lo, b_hi: the eight text bytes there, carried from the scan kernel's window (bytes beyond
  // the text are arbitrary: every use below is bounded by text_len)
  if (i + A.min_len > a.text_len) return;
  const uint8_t* tp = a.text + i;
  const uint32_t q = A.q;
  uint32_t g_lo = 0, g_hi = 0;
  const bool carried = i + 8 <= a.text_len;                   // (else the carried bytes run past the text: read it instead)
  if (CHECK_L3 && !fk_level3_pass(A, a, i, b_lo, b_hi)) return;
  // CaseSensitive: the carried bytes ARE the text; the stream starts behind them.  IgnoreCase in one pass: `runLower` lowers
  // every code point, so the carried bytes seed the lowering stream (ASCII is lowered in place, anything else decoded).
  uint32_t have = (!LOWER && carried) ? 8u : 0u;              // text bytes [0, have) are compared straight from (b_lo, b_hi)
  if (LOWER && carried && (b_lo & 0xC0u) == 0x80u) return;    // inside a code point: no needle starts here
  TextStream<LOWER> ts = LOWER ? TextStream<LOWER>(A, tp + (carried ? 8 : 0), a.text + a.text_len, carried ? ((unsigned long long)b_hi << 32) | b_lo : 0ull, carried ? 8u : 0u)
                               : TextStream<LOWER>(A, tp + have, a.text + a.text_len);
  if (have) {
    g_lo = q < 4 ? b_lo & A.qmask : b_lo;
    if (q > 4) g_hi = q == 6 ? b_hi & 0xFFFFu : b_hi;
  } else {
    if (LOWER && !carried && (__ldg(tp) & 0xC0u) == 0x80u) return;
    uint32_t w4;
    uint32_t k = 0;
    if (q >= 4 && ts.next4_ascii(&w4)) { g_lo = w4; k = 4; }
    for (; k < q; k++) {
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
          if (!LOWER) {
            // four bytes per round (independent loads), leaving at the first round that differs; the first round of a
            // 4-gram's tail is the carried bytes 4..7: one comparison, no load from the text
            const uint8_t* xp = tp + q;
            const uint8_t* tn = np + nhead;
            uint32_t k = 0;
            if (have == 8 && q == 4 && rest >= 4) {
              if (__ldg(reinterpret_cast<const uint32_t*>(tn)) != b_hi) return;   // (tails are 4-byte aligned in `tails`)
              k = 4;
            }
            for (; k < rest; k += 4) {
              uint32_t diff = 0;
#pragma unroll
              for (uint32_t j = 0; j < 4; j++)
                if (k + j < rest) diff |= (uint32_t)__ldg(xp + k + j) ^ (uint32_t)__ldg(tn + k + j);
              if (diff) return;
            }
          } else {
            uint32_t k = nhead;
            while (k < tl) {
              uint32_t w4;
              if (k + 4 <= tl && ((k & 3u) == 0) && ts.next4_ascii(&w4)) {   // (tails are 4-byte aligned in `tails`)
                if (w4 != __ldg(reinterpret_cast<const uint32_t*>(np + k))) return;
                k += 4;
              } else {
                if (ts.next() != (int)__ldg(np + k)) return;
                k++;
              }
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



// The same verification for the inline form of the scan kernel (CaseSensitive, q <= 4), as lean as it gets: the hot loop
// of filter_kernel carries this code at three sites, and every instruction in it is paid with most of the warp idle.  The
// jump table is keyed by the whole q-gram here (q <= 4), the tail of a slot starts right after it, the text is read as it
// stands.  g: the q-gram of the survivor at text index i (from the scan kernel's window).
template <int MODE>
__device__ __forceinline__ void fk_verify_short(const DevAutomaton& A, const ScanArgs& a, uint64_t i, uint32_t g, unsigned long long& local_count) {
  if (i + A.min_len > a.text_len) return;
  uint32_t idx = jump_hash(g) & A.jump_mask;
  uint32_t st;
  for (;;) {
    const uint4 s = __ldg(reinterpret_cast<const uint4*>(A.jump) + idx);
    if (s.y == NONE) return;
    if (s.x == g) {
#if FK_TAIL
      if (s.w & JUMP_SIMPLE) {
        // one needle path below this q-gram: compare its tail with the text in one go (independent loads)
        const uint32_t tl = s.w & JUMP_TAIL_MASK;
        const uint64_t end = i + A.q + tl;
        if (end > a.text_len || end <= a.report_begin) return;
        const uint8_t* tp = a.text + i + A.q;
        const uint8_t* np = A.tails + s.z;
        for (uint32_t k = 0; k < tl; k += 4) {               // four bytes per round, leaving at the first round that differs
          uint32_t diff = 0;
#pragma unroll
          for (uint32_t j = 0; j < 4; j++)
            if (k + j < tl) diff |= (uint32_t)__ldg(tp + k + j) ^ (uint32_t)__ldg(np + k + j);
          if (diff) return;
        }
        if (MODE == MODE_ANY) { *a.d_flag = 1; return; }
        if (s.w & JUMP_SINGLE) {
          if (MODE == MODE_COUNT) local_count += 1;
          else fk_emit(A, a, end, s.y);
        } else {
          fk_report_state<MODE>(A, a, s.y, end, local_count);   // duplicates of one needle: all ranks of the leaf
        }
        return;
      }
#endif
      st = s.y;                                            // not simple: the slot holds the depth-q state
      break;
    }
    idx = (idx + 1) & A.jump_mask;
  }
  uint32_t d = A.q;
  for (;;) {
    if (st & OWN_FLAG) {
      const uint64_t end = i + d;
      if (end > a.report_begin) {
        if (MODE == MODE_ANY) { *a.d_flag = 1; return; }
        fk_report_state<MODE>(A, a, st & ID_MASK, end, local_count);
      }
    }
    if (i + d >= a.text_len) return;
    st = edge_lookup(A, st & ID_MASK, (uint32_t)__ldg(a.text + i + d));
    if (st == NONE) return;
    d++;
  }
}

}  // namespace am
