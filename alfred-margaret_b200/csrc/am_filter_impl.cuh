// am_filter_impl.cuh -- the q-gram filter scan kernel of libam_b200 (see am_kernels.cu for the file-level notes).
// Included by am_filter_count.cu / am_filter_any.cu / am_filter_emit.cu, one translation unit per scan mode (FK_MODE), so
// that the template instantiations compile in parallel.
#include <cstddef>

#include "am_device.cuh"
#include "am_kernels.h"

// The CTA's dynamic shared memory under an unmangled name, so that the kernel can take its shared-space address
// with a plain `mov` (a compile-time constant) instead of converting a generic pointer.
extern "C" { extern __shared__ __align__(128) unsigned char am_fk_smem[]; }

namespace am {

// =====================================================================================================
// filter_kernel
// =====================================================================================================
// Per warp, per 4 KiB chunk, in pairs of 512-byte iterations:
//   1. stream two 16-byte granules per lane from HBM (next pair prefetched into registers) and mirror
//      them into the warp's 1 KiB shared-memory window;
//   2. probe the q-gram bitmap in shared memory: one probe per TWO text positions for q = 4 (stride-2 cells,
//      fk_probe16_s2), one per position for shorter q-grams (32 candidate bits per lane either way);
//   3. pop the candidate bits: re-read the q-gram from the window, test it against the second-level table T2
//      (shared memory: exact keys, or Bloom bits for large needle sets) -> "survivors";
//   4. survivors are queued per warp and, 32 at a time, walked through the goto trie in HBM/L2
//      (dense: all lanes busy); matches go to a per-warp stage, flushed with one global atomic.
// No CTA-wide barrier in the steady state; the only global atomics are per-warp stage flushes.
constexpr int FK_THREADS = 1024;                 // 32 warps, one CTA per SM (shared memory bound)
constexpr int FK_WARPS = FK_THREADS / 32;
#ifndef FK_NPAIRS
#define FK_NPAIRS 4
#endif
#ifndef FK_DRAIN_AT
#define FK_DRAIN_AT 32
#endif
constexpr int FK_PAIRS = FK_NPAIRS;              // pairs of 512-byte warp iterations per warp chunk
constexpr int FK_CHUNK = FK_PAIRS * 1024;        // bytes per warp chunk
constexpr int FK_TILE = FK_WARPS * FK_CHUNK;     // bytes per CTA tile (128 KiB)
constexpr int FK_WIN_WORDS = 256 + 4;            // window: 1 KiB pair + tail word (padded to 16 B)
constexpr int FK_SQ = 128;                       // survivor queue entries per warp
constexpr uint64_t FK_SPAN = 1ull << 40;         // bytes per launch (one launch per scan in practice; survivors carry 64-bit offsets)
// Build-time variants (A/B-tested on the GPU; the rejected ones -- warp-scan compaction of the candidates, bulk L2
// prefetch, IMAD.HI row addressing, an out-of-line survivor drain -- are recorded in profiles/README.md):
//   FK_DEBUG      compile the stage-isolation switches (AM_DEBUG_FLAGS=1: probes only, 2: no survivor walk)
#ifndef FK_DEBUG
#define FK_DEBUG 0
#endif

struct FilterSmem {
  uint32_t filter[FILTER_WORDS];                 // 128 KiB: [row][bank]
  uint32_t t2[T2_WORDS];                         // 32 KiB
  uint32_t window[FK_WARPS][FK_WIN_WORDS];       // 32.5 KiB
  unsigned long long sq_pos[FK_WARPS][FK_SQ];    // 32 KiB: survivor queue (offsets from v_begin)
  uint32_t sq_n[FK_WARPS];
  unsigned long long red[FK_WARPS];
  alignas(8) unsigned long long mbar;
};

// Context of the rare out-of-line work (survivor walk, flushes); everything in it is recomputed from kernel
// parameters and the thread index where it is used, so it costs the hot loop no registers.
struct FilterCtx {
  uint64_t v_begin; uint32_t a0, warp, lane;
  __device__ __forceinline__ FilterCtx(const ScanArgs& a, uint64_t vb)
      : v_begin(vb), a0((uint32_t)(reinterpret_cast<uintptr_t>(a.text) & 15)), warp(threadIdx.x >> 5), lane(threadIdx.x & 31) {}
};

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
template <bool LOWER>
struct TextStream;
template <>
struct TextStream<false> {
  const uint8_t* p; const uint8_t* end;
  __device__ __forceinline__ TextStream(const DevAutomaton&, const uint8_t* b, const uint8_t* e) : p(b), end(e) {}
  __device__ __forceinline__ int next() { return p < end ? (int)__ldg(p++) : -1; }
};
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
        const uint32_t raw = c0 | c1 << 8 | c2 << 16 | c3 << 24;
        const uint32_t cp = n == 2 ? ((c0 & 0x1Fu) << 6) | (c1 & 0x3Fu)
                          : n == 3 ? ((c0 & 0x0Fu) << 12) | ((c1 & 0x3Fu) << 6) | (c2 & 0x3Fu)
                                   : ((c0 & 0x07u) << 18) | ((c1 & 0x3Fu) << 12) | ((c2 & 0x3Fu) << 6) | (c3 & 0x3Fu);
        const uint32_t l = lower_cp(A, cp);
        const uint32_t ln = l < 0x80u ? 1u : l < 0x800u ? 2u : l < 0x10000u ? 3u : 4u;
        if (l == cp || ln != n) pend = raw;                   // unchanged, or kept because its lower case has another length
        else if (n == 2) pend = (0xC0u | (l >> 6)) | ((0x80u | (l & 0x3Fu)) << 8);
        else if (n == 3) pend = (0xE0u | (l >> 12)) | ((0x80u | ((l >> 6) & 0x3Fu)) << 8) | ((0x80u | (l & 0x3Fu)) << 16);
        else pend = (0xF0u | (l >> 18)) | ((0x80u | ((l >> 12) & 0x3Fu)) << 8) | ((0x80u | ((l >> 6) & 0x3Fu)) << 16) | ((0x80u | (l & 0x3Fu)) << 24);
      }
      npend = n; p += n;
    }
    const uint32_t b = pend & 0xFFu;
    pend >>= 8; npend--;
    return (int)b;
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
__device__ __forceinline__ void fk_deep_verify(const DevAutomaton& A, const ScanArgs& a, const FilterCtx& c, uint64_t v_rel, unsigned long long& local_count) {
  const uint64_t v = c.v_begin + v_rel;
  if (v < c.a0) return;
  const uint64_t i = v - c.a0;
  if (i + A.min_len > a.text_len) return;
  const uint8_t* tp = a.text + i;
  if (LOWER && (__ldg(tp) & 0xC0u) == 0x80u) return;          // inside a code point: no needle starts here
  TextStream<LOWER> ts(A, tp, a.text + a.text_len);
  const uint32_t q = A.q;
  uint32_t g_lo = 0, g_hi = 0;                                // the (lowered) q-gram; min_len >= q bytes exist
  for (uint32_t k = 0; k < q; k++) {
    const int b = ts.next();
    if (b < 0) return;
    if (k < 4) g_lo |= (uint32_t)b << (8 * k); else g_hi |= (uint32_t)b << (8 * (k - 4));
  }
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
            // four bytes per round (independent loads), leaving at the first round that differs: most survivors of a
            // needle set too large for the exact second level fail within the first bytes
            const uint8_t* xp = tp + q;
            np += nhead;
            for (uint32_t k = 0; k < rest; k += 4) {
              uint32_t diff = 0;
#pragma unroll
              for (uint32_t j = 0; j < 4; j++)
                if (k + j < rest) diff |= (uint32_t)__ldg(xp + k + j) ^ (uint32_t)__ldg(np + k + j);
              if (diff) return;
            }
          } else {
            for (uint32_t k = nhead; k < tl; k++)
              if (ts.next() != (int)__ldg(np + k)) return;
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
    const int ch = ts.next();
    if (ch < 0) return;
    st = edge_lookup(A, st & ID_MASK, (uint32_t)ch);
    if (st == NONE) return;
    d++;
  }
}

// Drain the warp's survivor queue (warp converged on entry and exit).
template <int MODE, bool LOWER>
__device__ __forceinline__ void fk_drain(const DevAutomaton& A, const ScanArgs& a, FilterSmem* sm, uint64_t v_begin,
                                         unsigned long long& local_count, uint32_t min_fill) {
  __syncwarp();
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t n = sm->sq_n[warp];
  if (n > FK_SQ) n = FK_SQ;
  if (n < min_fill || n == 0) return;
  const FilterCtx c(a, v_begin);
  for (uint32_t k = lane; k < n; k += 32) fk_deep_verify<MODE, LOWER>(A, a, c, sm->sq_pos[warp][k], local_count);
  __syncwarp();
  if (lane == 0) sm->sq_n[warp] = 0;
  __syncwarp();
}

__device__ __forceinline__ uint32_t lds32(uint32_t saddr) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr)); return v; }
__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
  uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr)); return v;
}

// 16 probes of one granule, stride-1 form (q < 4): w[0..3] own words, w[4] the word that follows.  `filt_lane` is the
// shared address of this lane's copy of the bitmap.
__device__ __forceinline__ uint32_t fk_probe16(uint32_t filt_lane, uint32_t qmask, uint32_t krow, uint32_t m, const uint32_t (&w)[6]) {
  // positions are visited last-to-first so that, after both granules, bit P of the mask is position P
#pragma unroll
  for (int k = 3; k >= 0; k--) {
#pragma unroll
    for (int j = 3; j >= 0; j--) {
      const uint32_t g = (j == 0 ? w[k] : __funnelshift_r(w[k], w[k + 1], 8 * j)) & qmask;
#if FK_WB
      const uint32_t y = g * HASH_MUL;                     // bit index = low 5 bits, row = top bits
      const uint32_t word = lds32((y >> (32 - FILTER_ROWBITS_S1)) * krow + filt_lane);   // SHF + IMAD(UR) + LDS
#else
      const uint32_t y = (g * HASH_MUL) >> 15;             // bits 0..4 bit index, top FILTER_ROWBITS_S1 bits row
      const uint32_t word = lds32((y & (((1u << FILTER_ROWBITS_S1) - 1u) << (17 - FILTER_ROWBITS_S1))) + filt_lane);
#endif
      const uint32_t t = __funnelshift_l(word, word, y);   // rotate the tested bit into bit 31
      m = __funnelshift_l(t, m, 1);                        // m = m << 1 | t >> 31
    }
  }
  return m;
}

// Stride-2 form of the 16 probes (q = 4, 6, 8; FK_S2): for every even position p ONE bitmap word answers both "a needle
// starts at p" and "a needle starts at p + 1".  The row is hashed from the q - 1 bytes the two q-grams share,
// text[p+1 .. p+q) -- q = 4: the 4-gram at p + 1 times HASH_MUL << 8, which discards its top byte; q = 6 / 8: plus the 4-gram at
// p + 3 / p + 5 times HASH_MUL_B << 8 (s2_hash) -- and the two bits are picked by rotating the word by text[p] and by
// text[p + q] (SHF uses the low 5 bits of the register, so any register whose low byte is that text byte serves).
// Per two text bytes, q = 4: 1.5 + 1 + 2 + 2 ALU-pipe instructions, 2 IMAD, 1 LDS (stride-1: 7.5, 4, 2); q > 4: one IMAD more.
// w[0..3]: the granule's words, w[4], w[5]: the two words that follow.
template <int ROWBITS>
__device__ __forceinline__ uint32_t fk_row_addr(uint32_t y, uint32_t krow, uint32_t filt_lane) {
  return (y >> (32 - ROWBITS)) * krow + filt_lane;   // SHF + IMAD (krow is a run-time value: keeps the address an IMAD)
}

template <int ROWBITS, int QK>
__device__ __forceinline__ uint32_t fk_probe16_s2(uint32_t filt_lane, uint32_t krow, uint32_t m, const uint32_t (&w)[6]) {
  uint32_t h[6], xa[5], xb[5];     // h[k]: low byte = text[4k + 2]; xa[k] / xb[k]: the 4-grams at 4k + 1 / 4k + 3
#pragma unroll
  for (int k = 0; k < 5; k++) {
    h[k] = __funnelshift_r(w[k], w[k + 1], 16);
    xa[k] = __funnelshift_r(w[k], w[k + 1], 8);
    xb[k] = __funnelshift_r(w[k], w[k + 1], 24);
  }
  h[5] = w[5] >> 16;
#pragma unroll
  for (int k = 3; k >= 0; k--) {
    {  // p = 4k + 2: positions 4k + 3 (cell B, private byte text[4k + 2 + q]) and 4k + 2 (cell A, text[4k + 2])
      const uint32_t y = QK == 4 ? xb[k] * HASH_MUL_S2 : QK == 6 ? xb[k] * HASH_MUL + xa[k + 1] * HASH_MUL_BS : xb[k] * HASH_MUL + xb[k + 1] * HASH_MUL_BS;
      const uint32_t word = lds32(fk_row_addr<ROWBITS>(y, krow, filt_lane));
      const uint32_t pb = QK == 4 ? h[k + 1] : QK == 6 ? w[k + 2] : h[k + 2];
      m = __funnelshift_l(__funnelshift_l(word, word, pb), m, 1);
      m = __funnelshift_l(__funnelshift_l(word, word, h[k]), m, 1);
    }
    {  // p = 4k: positions 4k + 1 (cell B, text[4k + q]) and 4k (cell A, text[4k])
      const uint32_t y = QK == 4 ? xa[k] * HASH_MUL_S2 : QK == 6 ? xa[k] * HASH_MUL + xb[k] * HASH_MUL_BS : xa[k] * HASH_MUL + xa[k + 1] * HASH_MUL_BS;
      const uint32_t word = lds32(fk_row_addr<ROWBITS>(y, krow, filt_lane));
      const uint32_t pb = QK == 4 ? w[k + 1] : QK == 6 ? h[k + 1] : w[k + 2];
      m = __funnelshift_l(__funnelshift_l(word, word, pb), m, 1);
      m = __funnelshift_l(__funnelshift_l(word, word, w[k]), m, 1);
    }
  }
  return m;
}

// Second-level test of one candidate at byte offset `o` of the warp's window: recover the q-gram (folded for IgnoreCase
// automata: the window holds the text as it is), look it up in T2 -- exact keys + the byte that must follow, or Bloom bits
// for large needle sets.  win_s / t2_s: shared-space addresses of the warp's window and of T2.
template <int QK, bool T2X, bool FOLD>
__device__ __forceinline__ bool fk_phase_a(const DevAutomaton& A, uint32_t win_s, uint32_t t2_s, uint32_t o) {
  const uint32_t wa = win_s + (o & ~3u);
  uint32_t lo = lds32(wa), hi = lds32(wa + 4);
  if (FOLD) { lo = fold8(lo); hi = fold8(hi); }
  const uint32_t sh = (o & 3u) * 8u;
  uint32_t g = __funnelshift_r(lo, hi, sh);
  if (QK == 0) g &= A.qmask;
  if (T2X) {
    uint32_t hb = (g * HASH_MUL2) >> (32 - T2_LOG2_BUCKETS);
    uint4 b = lds128(t2_s + (hb << 4));
    if (b.x != g && b.z != g) {
      if (!(b.w & T2_AUX_OVERFLOW)) return false;          // the common exit: one probe, no key of this bucket matches
      do {                                                 // the bucket overflowed at build time: the key may sit in a later one
        hb = (hb + 1) & ((1u << T2_LOG2_BUCKETS) - 1);
        b = lds128(t2_s + (hb << 4));
        if (b.x == g || b.z == g) break;
      } while (b.w & T2_AUX_OVERFLOW);
      if (b.x != g && b.z != g) return false;
    }
    const uint32_t aux = b.x == g ? b.y : b.w;
    if (aux & T2_AUX_ANY) return true;
    uint32_t nb;                                           // text byte right after the q-gram
    if (QK == 4) nb = (hi >> sh) & 0xFFu;
    else nb = (uint32_t)((((unsigned long long)hi << 32) | lo) >> (sh + 8u * A.q)) & 0xFFu;
    return nb == (aux & 0xFFu);
  } else if (QK > 4) {
    // Bloom filter over the whole q-gram: one bit in each half of T2
    uint32_t hi2 = lds32(wa + 8);
    if (FOLD) hi2 = fold8(hi2);
    uint32_t ghi = __funnelshift_r(hi, hi2, sh);
    if (QK == 6) ghi &= 0xFFFFu;
    const uint32_t b0 = t2q_bit0(g, ghi), b1 = t2q_bit1(g, ghi);
    const uint32_t w0 = lds32(t2_s + ((b0 >> 5) << 2));
    const uint32_t w1 = lds32(t2_s + ((T2Q_WORD1 + (b1 >> 5)) << 2));
    return ((w0 >> (b0 & 31)) & (w1 >> (b1 & 31))) & 1u;
  } else if (QK == 4) {
    // needle set too large for the exact table: closed 4-grams (T2A) or a fifth byte that continues a needle (T2B, T2C)
    const uint32_t nb = (hi >> sh) & 0xFFu;                 // text byte right after the 4-gram
    const uint32_t ba = t2a_bit(g), bb = t2b_bit(g, nb), bc = t2c_bit(g, nb);
    const uint32_t xa = lds32(t2_s + ((T2A_WORD0 + (ba >> 5)) << 2));
    const uint32_t xb = lds32(t2_s + ((T2B_WORD0 + (bb >> 5)) << 2));
    const uint32_t xc = lds32(t2_s + ((T2C_WORD0 + (bc >> 5)) << 2));
    return ((xa >> (ba & 31)) | ((xb >> (bb & 31)) & (xc >> (bc & 31)))) & 1u;
  } else {
    const uint32_t b2 = (g * HASH_MUL2) >> (32 - FILTER2_LOG2_BITS);
    return (lds32(t2_s + ((b2 >> 5) << 2)) >> (b2 & 31)) & 1u;
  }
}

// QK:   0 = stride-1 probe of q < 4 grams; 4, 6, 8 = stride-2 probe of q-grams.
// CASE: 0 = CaseSensitive; 1 = IgnoreCase on a lowered copy of the text; 2 = IgnoreCase in ONE pass over the original text (the
//       survivors are lowered on the fly, TextStream<true>).  Both IgnoreCase forms probe FOLDED bytes (fold8).
template <int MODE, int QK, bool T2X, int CASE>
__global__ void __launch_bounds__(FK_THREADS, 1) filter_kernel(const __grid_constant__ DevAutomaton A, const __grid_constant__ ScanArgs a, uint64_t v_begin, uint64_t num_tiles) {
  FilterSmem* sm = reinterpret_cast<FilterSmem*>(am_fk_smem);
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr bool LOWER = CASE == 2;
  constexpr bool FOLD = CASE != 0;
  constexpr bool S2 = QK >= 4;
  constexpr int COPIES = S2 ? filter_copies_s2(T2X) : FK_COPIES_S1;

  // ---- stage the filter bitmap and T2 into shared memory with TMA bulk copies -----------------------
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sm->mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    constexpr uint32_t total = FILTER_WORDS * 4 + T2_WORDS * 4;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&sm->mbar)), "r"(total) : "memory");
    constexpr uint32_t CH = 16384;
    for (uint32_t off = 0; off < FILTER_WORDS * 4; off += CH)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_u32(reinterpret_cast<unsigned char*>(sm->filter) + off)),
                   "l"(reinterpret_cast<const unsigned char*>(A.filter) + off), "r"(CH), "r"(smem_u32(&sm->mbar))
                   : "memory");
    for (uint32_t off = 0; off < T2_WORDS * 4; off += CH)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_u32(reinterpret_cast<unsigned char*>(sm->t2) + off)),
                   "l"(reinterpret_cast<const unsigned char*>(A.filter2) + off), "r"(CH), "r"(smem_u32(&sm->mbar))
                   : "memory");
  }
  if (lane == 0) sm->sq_n[warp] = 0;
  __syncthreads();
  {
    uint32_t done = 0;
    while (!done)
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                   : "=r"(done) : "r"(smem_u32(&sm->mbar)), "r"(0u) : "memory");
  }

  // ---- main loop ---------------------------------------------------------------------------------------------------
  // Software pipeline over (tile, pair): everything a pair needs -- its two granules per lane and the two words that
  // follow the pair -- was requested one pair earlier, including across chunk and tile boundaries.  The hot path
  // carries no bookkeeping: the pair loop is unrolled by two (register ping-pong instead of moves), the position of the next
  // pair is one warp-uniform granule index, loads that could leave the text are CLAMPED to its last granule instead
  // of being guarded (bytes beyond the text only ever reach candidates that the exact verification rejects on
  // bounds), shared-memory addresses are compile-time offsets from the CTA's dynamic shared memory base.
  const uintptr_t addr0 = reinterpret_cast<uintptr_t>(a.text);
  const uint32_t a0 = (uint32_t)(addr0 & 15);
  const uint4* base16 = reinterpret_cast<const uint4*>(addr0 - a0);
  const uint64_t nvec = (a0 + a.text_len + 15) >> 4;      // 16-byte granules overlapping the text (>= 1 here)
  uint32_t smem0;
  asm("mov.u32 %0, am_fk_smem;" : "=r"(smem0));
  const uint32_t filt_lane = smem0 + (uint32_t)offsetof(FilterSmem, filter) + ((lane & (COPIES - 1u)) << 2);
  const uint32_t win_s = smem0 + (uint32_t)offsetof(FilterSmem, window) + warp * (FK_WIN_WORDS * 4u);
  const uint32_t win_lane = win_s + (lane << 4);
  const uint32_t t2_s = smem0 + (uint32_t)offsetof(FilterSmem, t2);
  unsigned long long local_count = 0;

  auto load_pair = [&](uint64_t g, uint4& qa, uint4& qb, uint2& tail) {   // g: first granule of the pair (warp-uniform)
    if (g + 65 <= nvec) {                                  // granules g .. g + 64 exist
      const uint4* p = base16 + g + lane;
      qa = ld_stream_v4(p);
      qb = ld_stream_v4(p + 32);
      tail = __ldg(reinterpret_cast<const uint2*>(base16 + g + 64));
    } else {
      const uint64_t last = nvec - 1;
      const uint64_t ga = g + lane < last ? g + lane : last, gb = g + lane + 32 < last ? g + lane + 32 : last;
      const uint64_t gt = g + 64 < last ? g + 64 : last;
      qa = ld_stream_v4(base16 + ga);
      qb = ld_stream_v4(base16 + gb);
      tail = __ldg(reinterpret_cast<const uint2*>(base16 + gt));
    }
  };
  auto process_pair = [&](const uint4& qa_in, const uint4& qb_in, const uint2& tail_in, uint64_t tile_rel, uint32_t pair_rel) {
    uint4 qa = qa_in, qb = qb_in;
    uint2 tail = tail_in;
    // mirror the pair into the window (q-gram recovery for the few candidates)
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(win_lane), "r"(qa.x), "r"(qa.y), "r"(qa.z), "r"(qa.w) : "memory");
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(win_lane + 512u), "r"(qb.x), "r"(qb.y), "r"(qb.z), "r"(qb.w) : "memory");
    if (lane == 0) asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(win_s + 1024u), "r"(tail.x), "r"(tail.y) : "memory");
    if (FOLD) {
      // IgnoreCase automata hold the cells of FOLDED q-grams (fold8): the probe costs one OR per word on ASCII text and
      // four instructions per word where the warp meets a byte above ASCII.  The text itself stays in the window.
      const uint32_t high = (qa.x | qa.y | qa.z | qa.w | qb.x | qb.y | qb.z | qb.w | tail.x | tail.y) & 0x80808080u;
      if (__any_sync(0xFFFFFFFFu, high != 0)) {
        qa.x = fold8(qa.x); qa.y = fold8(qa.y); qa.z = fold8(qa.z); qa.w = fold8(qa.w);
        qb.x = fold8(qb.x); qb.y = fold8(qb.y); qb.z = fold8(qb.z); qb.w = fold8(qb.w);
        tail.x = fold8(tail.x); tail.y = fold8(tail.y);
      } else {
        qa.x |= 0x20202020u; qa.y |= 0x20202020u; qa.z |= 0x20202020u; qa.w |= 0x20202020u;
        qb.x |= 0x20202020u; qb.y |= 0x20202020u; qb.z |= 0x20202020u; qb.w |= 0x20202020u;
        tail.x |= 0x20202020u; tail.y |= 0x20202020u;
      }
    }
    // the words that follow each granule: the next lane's (lane 31: granule B of lane 0, resp. the tail)
    const uint32_t nl = (lane + 1) & 31;
    const uint32_t w4A = __shfl_sync(0xFFFFFFFFu, lane == 0 ? qb.x : qa.x, nl);
    const uint32_t w4B = __shfl_sync(0xFFFFFFFFu, lane == 0 ? tail.x : qb.x, nl);
    uint32_t w5A = 0, w5B = 0;
    if (QK > 4) {
      w5A = __shfl_sync(0xFFFFFFFFu, lane == 0 ? qb.y : qa.y, nl);
      w5B = __shfl_sync(0xFFFFFFFFu, lane == 0 ? tail.y : qb.y, nl);
    }
    const uint32_t wB[6] = {qb.x, qb.y, qb.z, qb.w, w4B, w5B}, wA[6] = {qa.x, qa.y, qa.z, qa.w, w4A, w5A};
    uint32_t m = 0;                                        // bit P <-> position P of the lane's 32 (0..15 granule A, 16..31 B)
    if (S2) {
      m = fk_probe16_s2<filter_rowbits(COPIES), QK>(filt_lane, a.krow, m, wB);
      m = fk_probe16_s2<filter_rowbits(COPIES), QK>(filt_lane, a.krow, m, wA);
    } else {
      m = fk_probe16(filt_lane, A.qmask, a.krow, m, wB);
      m = fk_probe16(filt_lane, A.qmask, a.krow, m, wA);
    }
    __syncwarp();
#if FK_DEBUG
    if (a.debug & 1u) { local_count += __popc(m); m = 0; }
#endif
    // ---- every lane pops its own candidate bits and tests them against T2 -------------------------------
    while (m) {
      uint32_t P;
      asm("bfind.u32 %0, %1;" : "=r"(P) : "r"(m));         // highest candidate position
      m ^= 1u << P;
      const uint32_t o = (P & 16u) * 31u + P;              // byte offset from the lane's granule A: (P >> 4) * 512 + (P & 15)
      if (fk_phase_a<QK, T2X, FOLD>(A, win_lane, t2_s, o)) {
#if FK_DEBUG
        if (a.debug & 2u) { local_count++; continue; }
#endif
        const uint64_t rel = tile_rel + (pair_rel + o + (lane << 4));
        const uint32_t qi = atomicAdd(&sm->sq_n[warp], 1u);
        if (qi < FK_SQ) sm->sq_pos[warp][qi] = rel;
        else fk_deep_verify<MODE, LOWER>(A, a, FilterCtx(a, v_begin), rel, local_count);   // queue full: verify in place
      }
    }
    fk_drain<MODE, LOWER>(A, a, sm, v_begin, local_count, FK_DRAIN_AT);       // only when a full round of survivors waits
  };

  static_assert(FK_PAIRS % 2 == 0, "the pair loop is unrolled by two");
  const uint64_t tile_stride_granules = (uint64_t)gridDim.x * (FK_TILE / 16);
  uint64_t g_next = ((v_begin + (uint64_t)blockIdx.x * FK_TILE + (uint64_t)warp * FK_CHUNK) >> 4);   // granule of the pair in flight
  uint4 cA, cB, nA, nB;
  uint2 tC, tN;
  if (blockIdx.x < num_tiles) load_pair(g_next, cA, cB, tC);
  for (uint64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    if (MODE == MODE_ANY && *reinterpret_cast<volatile int*>(a.d_flag)) break;
    const uint64_t tile_rel = tile * FK_TILE;              // this tile, relative to v_begin
    const uint32_t chunk_rel = warp * FK_CHUNK;            // this warp's chunk, relative to the tile
#pragma unroll 1
    for (int pair = 0; pair < FK_PAIRS; pair += 2) {
      g_next += 64;                                        // odd pair of the same chunk
      load_pair(g_next, nA, nB, tN);
      process_pair(cA, cB, tC, tile_rel, chunk_rel + (uint32_t)pair * 1024u);
      g_next += pair + 2 < FK_PAIRS ? 64 : tile_stride_granules - (FK_PAIRS - 1) * 64;   // next even pair: same chunk, or this warp's chunk in the CTA's next tile
      load_pair(g_next, cA, cB, tC);                       // beyond the CTA's last tile this is a clamped, unused load
      process_pair(nA, nB, tN, tile_rel, chunk_rel + (uint32_t)pair * 1024u + 1024u);
    }
  }
  fk_drain<MODE, LOWER>(A, a, sm, v_begin, local_count, 1);

  if (MODE == MODE_COUNT) {
    for (int o = 16; o > 0; o >>= 1) local_count += __shfl_down_sync(0xFFFFFFFFu, local_count, o);
    if (lane == 0) sm->red[warp] = local_count;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long s = 0;
      for (int i = 0; i < FK_WARPS; i++) s += sm->red[i];
      if (s) atomicAdd(a.d_count, s);
    }
  }
}

template <int MODE, int QK, bool T2X, int CASE>
static cudaError_t launch_filter_t(const DevAutomaton& A, const ScanArgs& a, cudaStream_t st) {
  if (a.text_len <= a.report_begin) return cudaSuccess;
  static std::atomic<uint64_t> attr_done{0};   // per device (am_options.device: one process may use several GPUs)
  {
    cudaError_t e = ensure_dynamic_smem(filter_kernel<MODE, QK, T2X, CASE>, (int)sizeof(FilterSmem), attr_done);
    if (e != cudaSuccess) return e;
  }
  const uint32_t a0 = (uint32_t)(reinterpret_cast<uintptr_t>(a.text) & 15);
  // first start position that can produce a match ending after report_begin
  const uint64_t first = a.report_begin + 1 > A.max_len ? a.report_begin + 1 - A.max_len : 0;
  const uint64_t v_end = a0 + a.text_len;
  for (uint64_t v0 = (first + a0) & ~15ull; v0 < v_end; v0 += FK_SPAN) {
    const uint64_t span = v_end - v0 < FK_SPAN ? v_end - v0 : FK_SPAN;
    const uint64_t tiles = (span + FK_TILE - 1) / FK_TILE;
    const uint64_t blocks = tiles < (uint64_t)sm_count() ? tiles : (uint64_t)sm_count();
    g_kernel_launches++;
    filter_kernel<MODE, QK, T2X, CASE><<<(unsigned)blocks, FK_THREADS, sizeof(FilterSmem), st>>>(A, a, v0, tiles);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

template <int MODE, int CASE>
static cudaError_t launch_filter_c(const DevAutomaton& A, const ScanArgs& a, cudaStream_t st) {
  const bool x = A.t2_exact != 0;
  switch (A.q) {
    case 8: return x ? cudaErrorInvalidValue : launch_filter_t<MODE, 8, false, CASE>(A, a, st);
    case 6: return x ? cudaErrorInvalidValue : launch_filter_t<MODE, 6, false, CASE>(A, a, st);
    case 4: return x ? launch_filter_t<MODE, 4, true, CASE>(A, a, st) : launch_filter_t<MODE, 4, false, CASE>(A, a, st);
    case 1: case 2: case 3: return x ? launch_filter_t<MODE, 0, true, CASE>(A, a, st) : launch_filter_t<MODE, 0, false, CASE>(A, a, st);
    default: return cudaErrorInvalidValue;
  }
}

// One scan mode per translation unit (FK_MODE): launch_filter (am_filter_count.cu) dispatches on the mode.
template <int MODE>
cudaError_t launch_filter_mode(const DevAutomaton& A, const ScanArgs& a, cudaStream_t st) {
  if (!A.ignore_case) return launch_filter_c<MODE, 0>(A, a, st);
  return a.ic_one_pass ? launch_filter_c<MODE, 2>(A, a, st) : launch_filter_c<MODE, 1>(A, a, st);
}

}  // namespace am
