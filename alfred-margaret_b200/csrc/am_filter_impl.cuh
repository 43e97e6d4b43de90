// am_filter_impl.cuh -- the q-gram filter scan kernel of libam_b200 (see am_kernels.cu for the file-level notes).
// Two forms of one kernel (template parameter IMODE):
//   * IMODE = -1, "list": the kernel only FILTERS; the positions that pass its two levels are appended to a list in global
//     memory and verified by verify_kernel (am_verify.cu), one survivor per thread.  IgnoreCase automata (whose verification
//     lowers code points: a lot of code) and the long q-grams of large needle sets take this form: the scan warps never
//     wait on a verification, and the hot loop's instruction footprint stays small.
//   * IMODE = a ScanMode, "inline": CaseSensitive automata with q <= 4 verify their survivors in the kernel, 32 at a time,
//     while the text they sit in is still in L2 -- measured 7 % faster on C2 than listing + verifying after the scan.
// Instantiated by am_filter_list.cu and am_filter_inline.cu (two translation units, compiled in parallel).
#include <cstddef>

#include "am_device.cuh"
#include "am_kernels.h"
#include "am_verify.cuh"

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
//   4. survivors are staged per warp and appended, 32 or more at a time, to the survivor list in global memory (one
//      atomic per flush); verify_kernel walks them through the goto trie afterwards, one survivor per thread.
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
constexpr int FK_SQ = 48;                        // survivor stage entries per warp (16 bytes each; 24 KiB in all: with 32 KiB the kernel loses 4 % -- the last KiBs of L1)
constexpr int FK_CAND = 384;                     // candidate list entries per warp (global second level)
constexpr int FK_CAND0 = 128;                    // candidate list entries per warp (bitmap second level in shared memory: 4 positions x 32 lanes always fit)
constexpr uint64_t FK_SPAN = 1ull << 40;         // bytes per launch (one launch per scan in practice; survivors carry 64-bit offsets)
// Build-time variants (A/B-tested on the GPU; the rejected ones -- warp-scan compaction of the candidates, bulk L2
// prefetch, IMAD.HI row addressing, an out-of-line survivor drain -- are recorded in profiles/README.md):
//   FK_DEBUG      compile the stage-isolation switches (AM_DEBUG_FLAGS=1: count the level-1 candidates and stop there,
//                 2: count the level-2 survivors instead of listing them)
#ifndef FK_DEBUG
#define FK_DEBUG 0
#endif

struct FilterSmem {
  uint32_t filter[FILTER_WORDS];                 // 128 KiB: [row][bank]
  union {
    uint32_t t2[T2_WORDS];                       // 32 KiB: second level (exact keys or bitmaps) ...
    uint16_t cand[FK_WARPS][FK_CAND];            // ... or, when the second level lives in global memory (q > 4), the warps' candidate lists
  };
  uint32_t window[FK_WARPS][FK_WIN_WORDS];       // 32.5 KiB
  ulonglong2 sq[FK_WARPS][FK_SQ];                // 24 KiB: survivor stage {text index, the eight text bytes there}
  uint16_t cand0[FK_WARPS][FK_CAND0];            //  8 KiB: the warps' candidate lists when the second level is the bitmaps in `t2` (large needle sets, q <= 4)
  uint32_t sq_n[FK_WARPS];
  uint32_t surv_n[FK_WARPS];                     // inline form: survivors verified so far (the host's survivor-rate monitor reads their sum)
  unsigned long long red[FK_WARPS];
  alignas(8) unsigned long long mbar;
};
static_assert(sizeof(uint16_t) * FK_WARPS * FK_CAND <= sizeof(uint32_t) * T2_WORDS, "candidate lists reuse the T2 region");

// Stage one survivor: virtual index v = a0 + text index, and the eight text bytes at that position (as they stand in the
// text, not folded): the verification of most survivors ends within them and never touches the text again.  A full stage
// appends to the list directly (list form) or verifies in place (inline form).
template <int IMODE, int QK>
__device__ __forceinline__ void fk_push(const DevAutomaton& A, FilterSmem* sm, const ScanArgs& a, uint32_t warp, uint32_t a0, uint64_t v, uint32_t b_lo, uint32_t b_hi,
                                        unsigned long long& local_count) {
  // (one structured if / else and no early return: ptxas then proves that the warp reconverges after the candidate loop.
  // With early returns here it did not -- a plain BSSY instead of BSSY.RECONVERGENT -- and 18 % of the pairs went on in
  // pieces: +9 % instructions, the flush check and the next pair's shuffles on their slow paths)
  const ulonglong2 e = make_ulonglong2(v, (unsigned long long)b_lo | ((unsigned long long)b_hi << 32));   // v < a0: bytes of the first granule that precede the text (dropped at the flush)
  const uint32_t qi = atomicAdd(&sm->sq_n[warp], 1u);
  if (qi < FK_SQ) {
    sm->sq[warp][qi] = e;
  } else if (v >= a0) {
    if (IMODE < 0) {
      const unsigned long long o = atomicAdd(a.surv_counts + blockIdx.x, 1ull);
      if (o < a.surv_cap_cta) a.surv[(uint64_t)blockIdx.x * a.surv_cap_cta + o] = e;
    } else {
      fk_verify_short<IMODE < 0 ? 0 : IMODE>(A, a, v - a0, QK == 0 ? b_lo & A.qmask : b_lo, local_count);
    }
  }
}
// The warp's staged survivors (warp converged on entry and exit): appended to the global list with one atomic and coalesced
// stores (list form), or verified, one per lane (inline form).
template <int IMODE, int QK>
__device__ __forceinline__ void fk_flush(const DevAutomaton& A, FilterSmem* sm, const ScanArgs& a, uint32_t warp, uint32_t lane, uint32_t a0, uint32_t min_fill,
                                         unsigned long long& local_count) {
  __syncwarp();
  uint32_t n = sm->sq_n[warp];
  if (n > FK_SQ) n = FK_SQ;
  if (n >= min_fill && n != 0) {
    if (IMODE < 0) {
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(a.surv_counts + blockIdx.x, (unsigned long long)n);   // this CTA's own counter
      base = __shfl_sync(0xFFFFFFFFu, base, 0);
      ulonglong2* region = a.surv + (uint64_t)blockIdx.x * a.surv_cap_cta;
      for (uint32_t k = lane; k < n; k += 32)
        if (base + k < a.surv_cap_cta) region[base + k] = sm->sq[warp][k];   // (virtual indices: verify_kernel drops v < a0 and subtracts a0)
    } else {
      if (lane == 0) sm->surv_n[warp] += n;
      for (uint32_t k = lane; k < n; k += 32) {
        const ulonglong2 e = sm->sq[warp][k];
        if (e.x >= a0) fk_verify_short<IMODE < 0 ? 0 : IMODE>(A, a, e.x - a0, QK == 0 ? (uint32_t)e.y & A.qmask : (uint32_t)e.y, local_count);
      }
    }
    __syncwarp();
    if (lane == 0) sm->sq_n[warp] = 0;
    __syncwarp();
  }
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

// Stride-2 form of the 16 probes (q = 4, FK_S2): for every even position p ONE bitmap word answers both "a needle
// starts at p" and "a needle starts at p + 1".  The row is hashed from text[p+1..p+4) -- the 4-gram at p + 1 times
// HASH_MUL << 8, which discards its top byte -- and the two bits are picked by rotating the word by text[p] and by
// text[p + 4] (SHF uses the low 5 bits of the register, so any register whose low byte is that text byte serves).
// Per two text bytes: 1.5 + 1 + 2 + 2 ALU-pipe instructions, 2 IMAD, 1 LDS (stride-1: 7.5, 4, 2).
// w[0..3]: the granule's words, w[4]: the word that follows.
template <int ROWBITS>
__device__ __forceinline__ uint32_t fk_row_addr(uint32_t y, uint32_t krow, uint32_t filt_lane) {
  return (y >> (32 - ROWBITS)) * krow + filt_lane;   // SHF + IMAD (krow is a run-time value: keeps the address an IMAD)
}

template <int ROWBITS>
__device__ __forceinline__ uint32_t fk_probe16_s2(uint32_t filt_lane, uint32_t krow, uint32_t m, const uint32_t (&w)[6]) {
  // h[k]: register whose low byte is text[4k + 2]
  uint32_t h[5];
#pragma unroll
  for (int k = 0; k < 4; k++) h[k] = __funnelshift_r(w[k], w[k + 1], 16);
  h[4] = w[4] >> 16;
#pragma unroll
  for (int k = 3; k >= 0; k--) {
    {  // p = 4k + 2: positions 4k + 3 (cell B, private byte text[4k + 6]) and 4k + 2 (cell A, text[4k + 2])
      const uint32_t y = __funnelshift_r(w[k], w[k + 1], 24) * HASH_MUL_S2;
      const uint32_t word = lds32(fk_row_addr<ROWBITS>(y, krow, filt_lane));
      m = __funnelshift_l(__funnelshift_l(word, word, h[k + 1]), m, 1);
      m = __funnelshift_l(__funnelshift_l(word, word, h[k]), m, 1);
    }
    {  // p = 4k: positions 4k + 1 (cell B, text[4k + 4]) and 4k (cell A, text[4k])
      const uint32_t y = __funnelshift_r(w[k], w[k + 1], 8) * HASH_MUL_S2;
      const uint32_t word = lds32(fk_row_addr<ROWBITS>(y, krow, filt_lane));
      m = __funnelshift_l(__funnelshift_l(word, word, w[k + 1]), m, 1);
      m = __funnelshift_l(__funnelshift_l(word, word, w[k]), m, 1);
    }
  }
  return m;
}

// Long q-grams (q = 6, 8): ONE cell per needle, probed at every position, that holds two bits of its word (long_cell):
//   t = X(p) * HASH_MUL,  y = t + X(p + q - 4) * HASH_MUL_B,  word = row y >> 17,  candidate = word[31 - (y & 31)] & word[31 - (t & 31)]
// The 4-grams X(0 .. 15 + q - 4) of the granule come out of the word pairs by funnel shifts, each used twice (as X(p) and as
// X(p' + q - 4)); per position 2 IMAD for the hash, 1 for the address, SHF (row) + 2 SHF (rotations) + LOP + SHF (append), 1 LDS.
// w[0..3]: the granule's words, w[4], w[5]: the two words that follow.
template <int QK>
__device__ __forceinline__ uint32_t fk_probe16_long(uint32_t filt_base, uint32_t m, const uint32_t (&w)[6]) {
  constexpr int D = QK - 4;                                  // distance of the second 4-gram
  uint32_t G[16 + D];
#pragma unroll
  for (int i = 0; i < 16 + D; i++) G[i] = (i & 3) == 0 ? w[i >> 2] : __funnelshift_r(w[i >> 2], w[(i >> 2) + 1], 8 * (i & 3));
  // positions are visited last-to-first so that, after both granules, bit P of the mask is position P
#pragma unroll
  for (int j = 15; j >= 0; j--) {
    const uint32_t t = G[j] * HASH_MUL;
    const uint32_t y = G[j + D] * HASH_MUL_B + t;
    const uint32_t word = lds32(((y >> (32 - FILTER_ROWBITS_LONG)) << 2) + filt_base);
    const uint32_t both = __funnelshift_l(word, word, y) & __funnelshift_l(word, word, t);   // bit 31: both bits of the cell are set
    m = __funnelshift_l(both, m, 1);
  }
  return m;
}

// Second-level test of one candidate at byte offset `o` of the warp's window: recover the q-gram (folded for IgnoreCase
// automata: the window holds the text as it is), look it up in T2 -- exact keys + the byte that must follow (T2M = 1), or
// bitmaps (T2M = 0).  win_s / t2_s: shared-space addresses of the warp's window and of T2.
template <int QK, int T2M, bool FOLD>
__device__ __forceinline__ bool fk_phase_a(const DevAutomaton& A, uint32_t win_s, uint32_t t2_s, uint32_t o, uint32_t* raw_lo) {
  const uint32_t wa = win_s + (o & ~3u);
  uint32_t lo = lds32(wa), hi = lds32(wa + 4);
  const uint32_t sh = (o & 3u) * 8u;
  *raw_lo = __funnelshift_r(lo, hi, sh);                     // the four text bytes at the candidate, as they are
  if (FOLD) { lo = fold20(lo); hi = fold20(hi); }
  uint32_t g = __funnelshift_r(lo, hi, sh);
  if (QK == 0) g &= A.qmask;
  if (T2M == 1) {
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

// Compaction of the warp's candidates (bit P of a lane's m = position P of its 32).  fk_scan: inclusive prefix sum of the lanes'
// candidate counts, returns the warp's total.  fk_scatter: lane l writes the window offsets of its candidates at
// [incl_l - popc(m_l), incl_l) of the warp's list in shared memory.
__device__ __forceinline__ uint32_t fk_scan(uint32_t m, uint32_t lane, uint32_t* incl_out) {
  uint32_t incl = __popc(m);
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= (uint32_t)d) incl += t; }
  *incl_out = incl;
  return __shfl_sync(0xFFFFFFFFu, incl, 31);
}
__device__ __forceinline__ void fk_scatter(uint32_t m, uint32_t incl, uint32_t cand_s, uint32_t lane) {
  uint32_t at = cand_s + ((incl - __popc(m)) << 1);
  while (m) {
    uint32_t P;
    asm("bfind.u32 %0, %1;" : "=r"(P) : "r"(m));
    m ^= 1u << P;
    const uint32_t o = (P & 16u) * 31u + P + (lane << 4);   // byte offset in the warp's window: (P >> 4) * 512 + (P & 15) + 16 * lane
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(at), "h"((unsigned short)o) : "memory");
    at += 2;
  }
  __syncwarp();
}

// Second level in GLOBAL memory (T2M = 2; q = 6, 8: needle sets of 10^4 .. 10^6 keys, whose level-1 bitmap passes several
// per cent of the positions).  The candidates of the warp's pair are compacted into a list in shared memory (a warp
// scan over the lanes' candidate counts) and looked up 32 at a time: every lane recovers one candidate's q-gram from the
// window and loads ONE word of the L2-resident bitmap -- all lanes busy, 64 independent loads in flight per warp --
// instead of every lane popping its own candidates while the others wait.  Survivors (true q-gram hits + ~0.4 %) go to
// the stage.  `m`: the lanes' candidate masks; cand_s: the warp's list (shared-space address).
template <int QK, bool FOLD, int IMODE>
__device__ __forceinline__ void fk_global_rounds(const DevAutomaton& A, const ScanArgs& a, FilterSmem* sm, uint32_t m, uint32_t incl, uint32_t total, uint32_t win_s,
                                                 uint32_t cand_s, uint32_t lane, uint32_t warp, uint32_t a0, uint64_t pair_v0, unsigned long long& local_count) {
  if (total == 0) return;
  fk_scatter(m, incl, cand_s, lane);
  // ---- rounds of 64 candidates: two independent bitmap loads per lane ----
  for (uint32_t r = 0; r < total; r += 64) {
    uint32_t o[2], glo[2], ghi[2], word[2], bit[2];
    bool live[2];
#pragma unroll
    for (int u = 0; u < 2; u++) {
      const uint32_t idx = r + u * 32 + lane;
      live[u] = idx < total;
      unsigned short ov = 0;
      if (live[u]) asm volatile("ld.shared.u16 %0, [%1];" : "=h"(ov) : "r"(cand_s + (idx << 1)));
      o[u] = ov;
      const uint32_t wa = win_s + (o[u] & ~3u);
      const uint32_t lo = lds32(wa), hi = lds32(wa + 4), hi2 = lds32(wa + 8);
      const uint32_t sh = (o[u] & 3u) * 8u;
      glo[u] = __funnelshift_r(lo, hi, sh);                  // the eight text bytes at the candidate, as they are
      ghi[u] = __funnelshift_r(hi, hi2, sh);
      uint32_t flo = glo[u], fhi = ghi[u];
      if (FOLD) { flo = fold20(flo); fhi = fold20(fhi); }
      if (QK == 6) fhi &= 0xFFFFu;
      bit[u] = gq_hash(flo, fhi) >> A.gbits_shift;
      word[u] = live[u] ? __ldg(A.gbits + (bit[u] >> 5)) : 0u;
    }
#pragma unroll
    for (int u = 0; u < 2; u++) {
      if (live[u] && ((word[u] >> (bit[u] & 31)) & 1u)) {
#if FK_DEBUG
        if (a.debug & 2u) { local_count++; continue; }
#endif
        fk_push<IMODE, QK>(A, sm, a, warp, a0, pair_v0 + o[u], glo[u], ghi[u], local_count);
      }
    }
    __syncwarp();
  }
}

// Second level = the bitmaps in shared memory (T2M = 0: needle sets beyond the exact table, q <= 4), candidates compacted as
// above and tested 32 at a time, one per lane: with 10^4 needles several per cent of the positions pass the first level (C3:
// 6 %, ~60 per pair), and popping them lane by lane costs the warp max-over-lanes rounds of one candidate each.
template <int QK, bool FOLD, int IMODE>
__device__ __forceinline__ void fk_shared_rounds(const DevAutomaton& A, const ScanArgs& a, FilterSmem* sm, uint32_t m, uint32_t incl, uint32_t total, uint32_t win_s,
                                                 uint32_t cand_s, uint32_t t2_s, uint32_t lane, uint32_t warp, uint32_t a0, uint64_t pair_v0, unsigned long long& local_count) {
  fk_scatter(m, incl, cand_s, lane);
  for (uint32_t r = 0; r < total; r += 32) {
    const uint32_t idx = r + lane;
    if (idx < total) {
      unsigned short ov;
      asm volatile("ld.shared.u16 %0, [%1];" : "=h"(ov) : "r"(cand_s + (idx << 1)));
      const uint32_t o = ov;
      uint32_t g;
      if (fk_phase_a<QK, 0, FOLD>(A, win_s, t2_s, o, &g)) {
#if FK_DEBUG
        if (a.debug & 2u) { local_count++; continue; }
#endif
        uint32_t raw_hi = 0;
        if (IMODE < 0) {
          const uint32_t wb = win_s + (o & ~3u) + 4u;
          raw_hi = __funnelshift_r(lds32(wb), lds32(wb + 4u), (o & 3u) * 8u);
        }
        fk_push<IMODE, QK>(A, sm, a, warp, a0, pair_v0 + o, g, raw_hi, local_count);
      }
    }
    __syncwarp();
  }
}

// QK:   0 = stride-1 probe of q < 4 grams; 4 = stride-2 probe of 4-grams; 6, 8 = stride-1 probe of long q-grams, two bits per cell.
// T2M:  second level: 0 = bitmaps in shared memory, 1 = exact keys in shared memory, 2 = bitmap in global memory (QK > 4).
// FOLD: IgnoreCase automata (on a lowered copy of the text or, in one pass, on the original text: verify_kernel lowers the
//       survivors on the fly): probe and second level see FOLDED bytes (every byte | 0x20).
template <int QK, int T2M, bool FOLD, int IMODE>
__global__ void __launch_bounds__(FK_THREADS, 1) filter_kernel(const __grid_constant__ DevAutomaton A, const __grid_constant__ ScanArgs a, uint64_t v_begin, uint64_t num_tiles) {
  FilterSmem* sm = reinterpret_cast<FilterSmem*>(am_fk_smem);
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr bool S2 = QK == 4;
  constexpr int COPIES = QK > 4 ? 1 : S2 ? filter_copies_s2(T2M == 1) : FK_COPIES_S1;
  static_assert(T2M != 2 || QK > 4, "the global second level serves the long q-grams");
  // Exact second level (survivors are rare: one in 2 000 positions at C2): look at the survivor stage once per TWO pairs -- one copy
  // of the flush / verification code in the unrolled loop instead of two (C2: -1.5 % COUNT, -2.2 % EMIT).  Sets that survive more
  // often would overflow the stage into the in-place verification (4 x 10^4 needles: +20 %), they look after every pair.
  constexpr bool FLUSH_PER_TWO = T2M == 1;
  constexpr bool TAIL8 = QK > 4 || IMODE < 0;                // the window carries two words beyond the pair (long q-grams; the list form's eight text bytes per survivor)

  // ---- stage the filter bitmap and T2 into shared memory with TMA bulk copies -----------------------
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sm->mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    constexpr uint32_t total = FILTER_WORDS * 4 + (T2M == 2 ? 0 : T2_WORDS * 4);   // (T2M = 2: the T2 region holds the candidate lists)
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&sm->mbar)), "r"(total) : "memory");
    constexpr uint32_t CH = 16384;
    for (uint32_t off = 0; off < FILTER_WORDS * 4; off += CH)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_u32(reinterpret_cast<unsigned char*>(sm->filter) + off)),
                   "l"(reinterpret_cast<const unsigned char*>(A.filter) + off), "r"(CH), "r"(smem_u32(&sm->mbar))
                   : "memory");
    for (uint32_t off = 0; T2M != 2 && off < T2_WORDS * 4; off += CH)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_u32(reinterpret_cast<unsigned char*>(sm->t2) + off)),
                   "l"(reinterpret_cast<const unsigned char*>(A.filter2) + off), "r"(CH), "r"(smem_u32(&sm->mbar))
                   : "memory");
  }
  if (lane == 0) { sm->sq_n[warp] = 0; sm->surv_n[warp] = 0; }
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
      if (TAIL8) tail = __ldg(reinterpret_cast<const uint2*>(base16 + g + 64));
      else tail = make_uint2(__ldg(reinterpret_cast<const uint32_t*>(base16 + g + 64)), 0u);   // (the inline form looks one word ahead)
    } else {
      const uint64_t last = nvec - 1;
      const uint64_t ga = g + lane < last ? g + lane : last, gb = g + lane + 32 < last ? g + lane + 32 : last;
      const uint64_t gt = g + 64 < last ? g + 64 : last;
      qa = ld_stream_v4(base16 + ga);
      qb = ld_stream_v4(base16 + gb);
      if (TAIL8) tail = __ldg(reinterpret_cast<const uint2*>(base16 + gt));
      else tail = make_uint2(__ldg(reinterpret_cast<const uint32_t*>(base16 + gt)), 0u);
    }
  };
  auto process_pair = [&](const uint4& qa_in, const uint4& qb_in, const uint2& tail_in, uint64_t tile_rel, uint32_t pair_rel) {
    uint4 qa = qa_in, qb = qb_in;
    uint2 tail = tail_in;
    // mirror the pair into the window (q-gram recovery for the few candidates)
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(win_lane), "r"(qa.x), "r"(qa.y), "r"(qa.z), "r"(qa.w) : "memory");
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(win_lane + 512u), "r"(qb.x), "r"(qb.y), "r"(qb.z), "r"(qb.w) : "memory");
    if (lane == 0) {
      if (TAIL8) asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(win_s + 1024u), "r"(tail.x), "r"(tail.y) : "memory");
      else asm volatile("st.shared.u32 [%0], %1;" ::"r"(win_s + 1024u), "r"(tail.x) : "memory");
    }
    if (FOLD) {
      // IgnoreCase automata hold the cells of FOLDED q-grams: one OR per word here.  The text itself stays in the window.
      qa.x |= FOLD_MASK; qa.y |= FOLD_MASK; qa.z |= FOLD_MASK; qa.w |= FOLD_MASK;
      qb.x |= FOLD_MASK; qb.y |= FOLD_MASK; qb.z |= FOLD_MASK; qb.w |= FOLD_MASK;
      tail.x |= FOLD_MASK; tail.y |= FOLD_MASK;
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
    if (QK > 4) {
      m = fk_probe16_long<QK>(filt_lane, m, wB);
      m = fk_probe16_long<QK>(filt_lane, m, wA);
    } else if (S2) {
      m = fk_probe16_s2<filter_rowbits(COPIES)>(filt_lane, a.krow, m, wB);
      m = fk_probe16_s2<filter_rowbits(COPIES)>(filt_lane, a.krow, m, wA);
    } else {
      m = fk_probe16(filt_lane, A.qmask, a.krow, m, wB);
      m = fk_probe16(filt_lane, A.qmask, a.krow, m, wA);
    }
    __syncwarp();
#if FK_DEBUG
    if (a.debug & 1u) { local_count += __popc(m); m = 0; }
#endif
    if (T2M == 2) {
      // ---- second level in global memory: compact the warp's candidates, look them up 64 at a time ----------------------
      const uint32_t cand_s = t2_s + warp * (FK_CAND * 2u);   // (sm->g.cand[warp]: the T2 region starts with the candidate lists)
      const uint64_t rel0 = v_begin + tile_rel + pair_rel;   // virtual index of the window's first byte
      uint32_t incl;
      const uint32_t total = fk_scan(m, lane, &incl);
      if (total <= FK_CAND) {
        fk_global_rounds<QK, FOLD, IMODE>(A, a, sm, m, incl, total, win_s, cand_s, lane, warp, a0, rel0, local_count);
      } else {                                             // more than the list holds: a quarter of the positions at a time (<= 256 each)
#pragma unroll 1
        for (uint32_t part = 0; part < 4; part++) {
          const uint32_t mp = m & (0xFFu << (8 * part));
          uint32_t ip;
          const uint32_t tp = fk_scan(mp, lane, &ip);
          fk_global_rounds<QK, FOLD, IMODE>(A, a, sm, mp, ip, tp, win_s, cand_s, lane, warp, a0, rel0, local_count);
        }
      }
    } else {
      // ---- second level in shared memory -------------------------------------------------------------------------------------
      // Bitmaps (T2M = 0, needle sets beyond the exact table): 10^4 needles pass several per cent of the positions at the first
      // level -- a few per lane and pair, unevenly spread -- so the warp compacts them and tests 32 at a time (C3: -10 %).  When
      // there are more than the list holds (4 x 10^4 needles: ~6 per lane) every lane is busy anyway and pops its own.
      bool popped = false;
      if (T2M == 0) {
        uint32_t incl;
        const uint32_t total = fk_scan(m, lane, &incl);
        if (total <= FK_CAND0) {
          const uint32_t cand_s = smem0 + (uint32_t)offsetof(FilterSmem, cand0) + warp * (FK_CAND0 * 2u);
          if (total) fk_shared_rounds<QK, FOLD, IMODE>(A, a, sm, m, incl, total, win_s, cand_s, t2_s, lane, warp, a0, v_begin + tile_rel + pair_rel, local_count);
          popped = true;
        }
      }
      if (popped) m = 0;
      // every lane pops its own candidate bits and tests them against T2 (exact table: candidates are rare)
      while (m) {
        uint32_t P;
        asm("bfind.u32 %0, %1;" : "=r"(P) : "r"(m));         // highest candidate position
        m ^= 1u << P;
        const uint32_t o = (P & 16u) * 31u + P;              // byte offset from the lane's granule A: (P >> 4) * 512 + (P & 15)
        uint32_t g;
        if (fk_phase_a<QK, T2M, FOLD>(A, win_lane, t2_s, o, &g)) {
#if FK_DEBUG
          if (a.debug & 2u) { local_count++; continue; }
#endif
          // the four bytes after them (survivors only: two more window words)
          uint32_t raw_hi = 0;                                 // (the inline form verifies from the q-gram alone)
          if (IMODE < 0) {
            const uint32_t wb = win_lane + (o & ~3u) + 4u;
            raw_hi = __funnelshift_r(lds32(wb), lds32(wb + 4u), (o & 3u) * 8u);
          }
          fk_push<IMODE, QK>(A, sm, a, warp, a0, v_begin + tile_rel + (pair_rel + o + (lane << 4)), g, raw_hi, local_count);
        }
      }
    }
    if (!FLUSH_PER_TWO) fk_flush<IMODE, QK>(A, sm, a, warp, lane, a0, FK_DRAIN_AT, local_count);   // only when a full round of survivors waits
    else __syncwarp();                                       // (lanes still popping candidates read the window the next pair overwrites)
  };

  static_assert(FK_PAIRS % 2 == 0, "the pair loop is unrolled by two");
  const uint64_t tile_stride_granules = (uint64_t)gridDim.x * (FK_TILE / 16);
  uint64_t g_next = ((v_begin + (uint64_t)blockIdx.x * FK_TILE + (uint64_t)warp * FK_CHUNK) >> 4);   // granule of the pair in flight
  uint4 cA, cB, nA, nB;
  uint2 tC, tN;
  if (blockIdx.x < num_tiles) load_pair(g_next, cA, cB, tC);
  for (uint64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    // containsAny: stop once a match has been flagged.  Inline form: by any CTA of this launch, hence a volatile poll -- compiled
    // into the ANY kernels only: around a STRONG load that decides a loop exit ptxas assumes a spin loop, plants YIELD and drops
    // the reconvergence guarantee of the loops inside (measured on C2: 18 % of the pairs then ran in pieces, +9 % instructions,
    // -10 % throughput -- in every mode, when the poll was a run-time switch).  List form: only verify_kernel sets the flag, i.e.
    // an EARLIER launch (a previous chunk of the text); a weak load is enough for that and keeps the loops reconvergent.
    if (IMODE == MODE_ANY) {
      if (*reinterpret_cast<volatile int*>(a.d_flag)) break;
    } else if (IMODE < 0 && a.any_mode) {
      int f;
      asm volatile("ld.global.s32 %0, [%1];" : "=r"(f) : "l"(a.d_flag) : "memory");
      if (f) break;
    }
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
      if (FLUSH_PER_TWO) fk_flush<IMODE, QK>(A, sm, a, warp, lane, a0, FK_DRAIN_AT, local_count);
    }
  }
  fk_flush<IMODE, QK>(A, sm, a, warp, lane, a0, 1, local_count);
  if (IMODE >= 0 && lane == 0 && sm->surv_n[warp]) atomicAdd(a.surv_count + 1, (unsigned long long)sm->surv_n[warp]);

  if (IMODE == MODE_COUNT || (FK_DEBUG && a.debug)) {        // (the stage-isolation counters of FK_DEBUG builds land in d_count too)
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

template <int QK, int T2M, bool FOLD, int IMODE>
static cudaError_t launch_filter_t(const DevAutomaton& A, const ScanArgs& a, cudaStream_t st) {
  static std::atomic<uint64_t> attr_done{0};   // per device (am_options.device: one process may use several GPUs)
  {
    cudaError_t e = ensure_dynamic_smem(filter_kernel<QK, T2M, FOLD, IMODE>, (int)sizeof(FilterSmem), attr_done);
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
    filter_kernel<QK, T2M, FOLD, IMODE><<<(unsigned)blocks, FK_THREADS, sizeof(FilterSmem), st>>>(A, a, v0, tiles);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

}  // namespace am
