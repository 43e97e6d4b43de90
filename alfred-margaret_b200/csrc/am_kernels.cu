// am_kernels.cu -- sm_100a kernels of libam_b200 other than the q-gram filter scan (am_filter.cu).
//
// The library replaces the inner loop of `runWithCase` (src/Data/Text/AhoCorasick/Automaton.hs:442-534):
// consumeInput / followCodePoint / lookupTransition / collectMatches.  Two formulations:
//
//   walk_kernel    (this file) every thread walks one text segment (plus a halo of max-needle-length - 1
//                  bytes of warm-up) through the byte-level goto+failure automaton and reports
//                  the matches that END inside its segment.  General: handles empty needles,
//                  IgnoreCase (decode -> Char.toLower table -> re-encode on the fly), any density.
//
//   filter_kernel  (am_filter.cu) position-parallel: every text position is tested against a q-gram membership
//                  bitmap held in shared memory, staged there by TMA bulk copies -- one probe per two positions;
//                  the few surviving positions pass an exact second level and are verified against the
//                  failure-less goto trie.  The haystack is read once from HBM with 128-bit streaming loads.
//
// Both produce (end_pos << rank_bits | rank) keys; ordering them restores the reference's callback order
// (end_pos ascending, then longest needle / later duplicate first).  Also here: lower_kernel (IgnoreCase front
// end: the lowered copy of the text), the segment scan / sort / compaction that orders the filter kernel's keys
// without a global sort, the radix sort for the walk kernel's keys, unpack_kernel (key -> am_match).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <thrust/iterator/transform_iterator.h>

#include "am_device.cuh"
#include "am_kernels.h"

namespace am {

// =====================================================================================================
// walk_kernel
// =====================================================================================================
// Every thread walks one text segment through the failure-resolved, class-compressed automaton:
//   class = cls[byte]                 (256-byte table in shared memory)
//   next  = row(state)[class]         (rows of the shallowest states are staged into shared memory with a TMA
//                                      bulk copy; deeper rows come from L2 / HBM through the read-only path)
// The thread starts max-needle-length - 1 bytes (the halo) before its segment so that its state is exact by
// the time it reaches the segment, and reports the matches that END inside the segment.
// A warp step costs the latency of its SLOWEST lane, so one row lookup that leaves shared memory (L2: ~700
// cycles) stalls 32 lanes: the fraction of lookups served from shared memory matters more than occupancy.
// Hence one CTA of 1024 threads per SM with as many hot rows as fit (ncu: with 88 KiB of rows per CTA 86 % of the
// warp steps on C2 waited on L2).
// For larger automata (hot set >> shared memory) the opposite holds: L1 capacity for the row lookups and
// two CTAs per SM win (measured: 100 k needles 214 vs 112 GB/s).  The launcher picks per automaton.
constexpr int WALK_THREADS = 1024;             // upper bound of the block size (launch bounds)
constexpr int WALK_STAGE = 1024;
constexpr int WALK_HOT_BYTES_BIG = 208 * 1024; // hot rows, 1 CTA of 1024 threads per SM (small automata)
constexpr int WALK_HOT_BYTES_SMALL = 88 * 1024;// hot rows, 2 CTAs of 512 threads per SM (large automata)
constexpr uint64_t WALK_SMALL_AUTOMATON_BYTES = 2ull << 20;   // all rows <= 2 MiB => the big-hot-set configuration

struct WalkSmem {
  uint8_t cls[256];
  KeyStage<WALK_STAGE> stage;
  unsigned long long red[WALK_THREADS / 32];
  alignas(8) unsigned long long mbar;
  alignas(16) uint32_t hot[1];                 // hot rows follow (dynamic shared memory)
};


template <int MODE>
__device__ __forceinline__ void report_chain(const DevAutomaton& A, const ScanArgs& a, uint32_t tagged,
                                             uint64_t pos, unsigned long long& local_count,
                                             KeyStage<WALK_STAGE>* stage) {
  // collectMatches (Automaton.hs:522-534) over values[state] = own ++ values[fail state]
  const uint32_t s = tagged & ID_MASK;
  if (MODE == MODE_COUNT) {
    local_count += __ldg(A.chain_count + s);
  } else if (MODE == MODE_ANY) {
    *a.d_flag = 1;
  } else {
    for (uint32_t t = __ldg(A.first_out + s); t != NONE; t = __ldg(A.next_out + t)) {
      const uint32_t lo = __ldg(A.own_off + t), hi = __ldg(A.own_off + t + 1);
      for (uint32_t j = lo; j < hi; j++)
        stage->push(a, ((unsigned long long)(pos + a.pos_base) << A.rank_bits) | __ldg(A.own_rank + j));
    }
  }
}

// One automaton step on byte `b`; `state` untagged, result tagged.
__device__ __forceinline__ uint32_t walk_step(const DevAutomaton& A, const WalkSmem* sm, uint32_t hot_states, uint32_t state, uint32_t b) {
  const uint32_t idx = (state << A.cdfa_shift) + sm->cls[b];
  if (state < hot_states) return sm->hot[idx];
  if (state < A.cdfa_states) return __ldg(A.cdfa + idx);
  return ac_step(A, state, b);                             // beyond the row budget (> 256 MiB of rows): goto + failure
}

// Incremental decodeN (Utf8.hs:344-350): returns true when `cp` is complete.
__device__ __forceinline__ bool walk_decode(uint32_t byte, uint32_t& cp, uint32_t& rem) {
  if (rem == 0) {
    if (byte < 0xC0u) { cp = byte; return true; }
    if (byte < 0xE0u) { cp = byte & 0x1Fu; rem = 1; return false; }
    if (byte < 0xF0u) { cp = byte & 0x0Fu; rem = 2; return false; }
    cp = byte & 0x07u; rem = 3; return false;
  }
  cp = (cp << 6) | (byte & 0x3Fu);
  return --rem == 0;
}

// Feed the UTF-8 bytes of one (lowered) code point; returns the tagged state after its last byte.
__device__ __forceinline__ uint32_t walk_feed_cp(const DevAutomaton& A, const WalkSmem* sm, uint32_t hot_states, uint32_t state, uint32_t l) {
  uint32_t t;
  if (l < 0x80u) return walk_step(A, sm, hot_states, state, l);
  if (l < 0x800u) {
    t = walk_step(A, sm, hot_states, state, 0xC0u | (l >> 6));
    return walk_step(A, sm, hot_states, t & ID_MASK, 0x80u | (l & 0x3Fu));
  }
  if (l < 0x10000u) {
    t = walk_step(A, sm, hot_states, state, 0xE0u | (l >> 12));
    t = walk_step(A, sm, hot_states, t & ID_MASK, 0x80u | ((l >> 6) & 0x3Fu));
    return walk_step(A, sm, hot_states, t & ID_MASK, 0x80u | (l & 0x3Fu));
  }
  t = walk_step(A, sm, hot_states, state, 0xF0u | (l >> 18));
  t = walk_step(A, sm, hot_states, t & ID_MASK, 0x80u | ((l >> 12) & 0x3Fu));
  t = walk_step(A, sm, hot_states, t & ID_MASK, 0x80u | ((l >> 6) & 0x3Fu));
  return walk_step(A, sm, hot_states, t & ID_MASK, 0x80u | (l & 0x3Fu));
}

// IgnoreCase: consume one text byte of a chain (consumeInput, Automaton.hs:468-480): decode incrementally
// (Utf8.hs:337-350), lower (Utf8.hs:145-151), feed the bytes of the lowered code point.  Returns false when the
// byte did not complete a code point (no state change, nothing to report).
__device__ __forceinline__ bool walk_ic_byte(const DevAutomaton& A, const WalkSmem* sm, uint32_t hot_states, uint32_t byte,
                                             uint32_t& state, uint32_t& cp, uint32_t& rem, bool& synced, uint32_t& t) {
  if (!synced) { if ((byte & 0xC0u) == 0x80u) return false; synced = true; }   // start on a code point boundary
  if (rem == 0 && byte < 0x80u) {                            // ASCII fast path: toLowerAscii (Utf8.hs:131-135)
    t = walk_step(A, sm, hot_states, state, byte + ((byte - 'A' < 26u) ? 0x20u : 0u));
  } else {
    if (!walk_decode(byte, cp, rem)) return false;
    t = walk_feed_cp(A, sm, hot_states, state, lower_cp(A, cp));
  }
  state = t & ID_MASK;
  return true;
}

// Walk ONE segment: the matches whose end position lies in (b, e].  General (edge) path.
template <bool IGNORE_CASE, int MODE>
__device__ __forceinline__ void walk_one_segment(const DevAutomaton& A, const ScanArgs& a, WalkSmem* sm, uint32_t hot_states, uint64_t seg,
                                                 uint64_t seg_bytes, uint32_t a0, const uint4* base16, unsigned long long& local_count) {
  const uint64_t b = a.report_begin + seg * seg_bytes;
  uint64_t e = b + seg_bytes; if (e > a.text_len) e = a.text_len;
  const uint64_t w = b > A.halo ? b - A.halo : 0;  // warm-up start: depth(state) <= max needle length
  uint32_t state = 0;                              // untagged
  uint32_t cp = 0, rem = 0;                        // incremental UTF-8 decoder (IgnoreCase)
  bool synced = !(IGNORE_CASE && (w > 0 || a.report_begin > 0));
  uint64_t v = w + a0; const uint64_t vend = e + a0;
  uint64_t c = v >> 4;
  uint4 q4 = __ldg(base16 + c);
  while (v < vend) {
    const uint64_t cn = c + 1;
    uint4 nq = make_uint4(0, 0, 0, 0);
    if ((cn << 4) < vend) nq = __ldg(base16 + cn);          // next granule, requested before this one is walked
    const uint32_t words[4] = {q4.x, q4.y, q4.z, q4.w};
    const uint32_t jlo = (uint32_t)(v & 15);
    const uint64_t left = vend - (c << 4);
    const uint32_t jhi = left < 16 ? (uint32_t)left : 16u;
#pragma unroll
    for (uint32_t j = 0; j < 16; j++) {
      if (j < jlo || j >= jhi) continue;
      const uint32_t byte = (words[j >> 2] >> (8 * (j & 3))) & 0xFFu;
      const uint64_t pos = (c << 4) + j - a0 + 1;            // offset one past this byte
      uint32_t t;
      if (!IGNORE_CASE) { t = walk_step(A, sm, hot_states, state, byte); state = t & ID_MASK; }
      else if (!walk_ic_byte(A, sm, hot_states, byte, state, cp, rem, synced, t)) continue;
      if ((t & OUT_FLAG) && pos > b) report_chain<MODE>(A, a, t, pos, local_count, &sm->stage);
    }
    v = cn << 4; c = cn; q4 = nq;
  }
}

// Walk TWO full-length interior segments in lockstep.  A warp step costs the latency of its slowest lane, and for
// automata whose hot set exceeds shared memory that latency is an L2 round trip; two independent chains per
// thread put two lookups in flight per lane.  Both segments have the same length and the same alignment, so
// they share every byte mask; both warm up over the same (16-byte rounded) halo.
template <bool IGNORE_CASE, int MODE>
__device__ __forceinline__ void walk_two_segments(const DevAutomaton& A, const ScanArgs& a, WalkSmem* sm, uint32_t hot_states, uint64_t seg,
                                                  uint64_t seg_bytes, uint32_t halo_al, uint32_t a0, const uint4* base16,
                                                  unsigned long long& local_count) {
  const uint64_t b0 = a.report_begin + seg * seg_bytes, b1 = b0 + seg_bytes;
  const uint64_t w0 = b0 - halo_al, w1 = b1 - halo_al;
  const uint64_t v0 = w0 + a0;                                // virtual index of chain 0's first byte
  const uint64_t vend0 = b0 + seg_bytes + a0;
  const uint64_t c0 = v0 >> 4, c1 = (w1 + a0) >> 4;
  const uint32_t nsteps = (uint32_t)(((vend0 + 15) >> 4) - c0);
  const uint32_t r = (uint32_t)(v0 & 15);                     // bytes of the first granule that precede the chains
  uint32_t s0 = 0, s1 = 0;                                    // untagged states
  uint32_t cp0 = 0, cp1 = 0, rem0 = 0, rem1 = 0;
  bool syn0 = !IGNORE_CASE, syn1 = !IGNORE_CASE;              // IgnoreCase chains start on a code point boundary
  uint4 q0 = __ldg(base16 + c0), q1 = __ldg(base16 + c1);
  for (uint32_t step = 0; step < nsteps; step++) {
    uint4 n0 = make_uint4(0, 0, 0, 0), n1 = n0;
    if (step + 1 < nsteps) { n0 = __ldg(base16 + c0 + step + 1); n1 = __ldg(base16 + c1 + step + 1); }
    const uint32_t wa[4] = {q0.x, q0.y, q0.z, q0.w}, wb[4] = {q1.x, q1.y, q1.z, q1.w};
    const uint32_t jlo = step == 0 ? r : 0u;
    const uint32_t jhi = step + 1 == nsteps ? (uint32_t)((vend0 - 1) & 15) + 1u : 16u;
    const uint32_t tbase = step * 16u - r;                    // chain-relative index of byte j = 0 (wraps for j < r; unused there)
#pragma unroll
    for (uint32_t j = 0; j < 16; j++) {
      if (j < jlo || j >= jhi) continue;
      const uint32_t x0 = (wa[j >> 2] >> (8 * (j & 3))) & 0xFFu, x1 = (wb[j >> 2] >> (8 * (j & 3))) & 0xFFu;
      const bool rep = tbase + j >= halo_al;                  // end position beyond the segment start
      uint32_t t0 = 0, t1 = 0;
      bool ok0 = true, ok1 = true;
      if (!IGNORE_CASE || (syn0 && syn1 && ((rem0 | rem1) == 0) && ((x0 | x1) < 0x80u))) {
        // both chains take one plain automaton step: issue the two class lookups, then the two row lookups
        const uint32_t y0 = IGNORE_CASE ? x0 + ((x0 - 'A' < 26u) ? 0x20u : 0u) : x0;
        const uint32_t y1 = IGNORE_CASE ? x1 + ((x1 - 'A' < 26u) ? 0x20u : 0u) : x1;
        const uint32_t k0 = sm->cls[y0], k1 = sm->cls[y1];
        const uint32_t i0 = (s0 << A.cdfa_shift) + k0, i1 = (s1 << A.cdfa_shift) + k1;
        t0 = s0 < hot_states ? sm->hot[i0] : __ldg(A.cdfa + i0);
        t1 = s1 < hot_states ? sm->hot[i1] : __ldg(A.cdfa + i1);
        s0 = t0 & ID_MASK; s1 = t1 & ID_MASK;
      } else {
        ok0 = walk_ic_byte(A, sm, hot_states, x0, s0, cp0, rem0, syn0, t0);
        ok1 = walk_ic_byte(A, sm, hot_states, x1, s1, cp1, rem1, syn1, t1);
      }
      if (rep && ((t0 | t1) & OUT_FLAG)) {
        if (ok0 && (t0 & OUT_FLAG)) report_chain<MODE>(A, a, t0, w0 + tbase + j + 1, local_count, &sm->stage);
        if (ok1 && (t1 & OUT_FLAG)) report_chain<MODE>(A, a, t1, w1 + tbase + j + 1, local_count, &sm->stage);
      }
    }
    q0 = n0; q1 = n1;
  }
}

template <bool IGNORE_CASE, int MODE>
__global__ void __launch_bounds__(WALK_THREADS, 1) walk_kernel(DevAutomaton A, ScanArgs a, uint64_t seg_bytes, uint64_t num_segs, uint32_t hot_states) {
  extern __shared__ __align__(128) unsigned char walk_smem_raw[];
  WalkSmem* sm = reinterpret_cast<WalkSmem*>(walk_smem_raw);
  // ---- stage the class map and the hot rows (TMA bulk copy of the row prefix) -----------------------------
  const uint32_t hot_bytes = (hot_states << A.cdfa_shift) * 4u;    // multiple of 16: a row is >= 2 words, hot_states is even or rows >= 16 B
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sm->mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&sm->mbar)), "r"(hot_bytes + 256u) : "memory");
    for (uint32_t off = 0; off < hot_bytes; off += 16384) {
      const uint32_t n = hot_bytes - off < 16384 ? hot_bytes - off : 16384;
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_u32(reinterpret_cast<unsigned char*>(sm->hot) + off)),
                   "l"(reinterpret_cast<const unsigned char*>(A.cdfa) + off), "r"(n), "r"(smem_u32(&sm->mbar))
                   : "memory");
    }
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sm->cls)),
                 "l"(A.cls), "r"(256u), "r"(smem_u32(&sm->mbar))
                 : "memory");
  }
  if (MODE == MODE_EMIT) sm->stage.init();
  __syncthreads();
  {
    uint32_t done = 0;
    while (!done)
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                   : "=r"(done) : "r"(smem_u32(&sm->mbar)), "r"(0u) : "memory");
  }
  unsigned long long local_count = 0;

  const uintptr_t addr0 = reinterpret_cast<uintptr_t>(a.text);
  const uint32_t a0 = (uint32_t)(addr0 & 15);
  const uint4* base16 = reinterpret_cast<const uint4*>(addr0 - a0);
  const uint32_t halo_al = (A.halo + 15u) & ~15u;
  // the lockstep path indexes the row table directly; beyond an L2-sized table the row lookups are bandwidth-
  // rather than latency-bound and a second chain only adds pressure (measured: 100 k needles 213 vs 196 GB/s)
  const bool pairable = A.cdfa_states == A.num_states && ((uint64_t)A.cdfa_states << A.cdfa_shift) * 4 <= (64ull << 20);
  const uint64_t num_pairs = (num_segs + 1) >> 1;

  for (uint64_t p0 = (uint64_t)blockIdx.x * blockDim.x; p0 < num_pairs; p0 += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t pr = p0 + threadIdx.x;
    bool live = pr < num_pairs;
    if (MODE == MODE_ANY && live && *reinterpret_cast<volatile int*>(a.d_flag)) live = false;
    if (live) {
      const uint64_t seg = pr * 2;
      const uint64_t b0 = a.report_begin + seg * seg_bytes;
      // interior pair: both segments exist with full length and the rounded halo does not reach before the text
      if (pairable && seg + 1 < num_segs && b0 >= halo_al && b0 + 2 * seg_bytes <= a.text_len) {
        walk_two_segments<IGNORE_CASE, MODE>(A, a, sm, hot_states, seg, seg_bytes, halo_al, a0, base16, local_count);
      } else {
        walk_one_segment<IGNORE_CASE, MODE>(A, a, sm, hot_states, seg, seg_bytes, a0, base16, local_count);
        if (seg + 1 < num_segs) walk_one_segment<IGNORE_CASE, MODE>(A, a, sm, hot_states, seg + 1, seg_bytes, a0, base16, local_count);
      }
    }
    if (MODE == MODE_EMIT) sm->stage.flush(a);
  }

  if (MODE == MODE_COUNT) {
    // CTA reduction, one global atomic per CTA
    for (int o = 16; o > 0; o >>= 1) local_count += __shfl_down_sync(0xFFFFFFFFu, local_count, o);
    if ((threadIdx.x & 31) == 0) sm->red[threadIdx.x >> 5] = local_count;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long s = 0;
      for (unsigned i = 0; i < blockDim.x / 32; i++) s += sm->red[i];
      if (s) atomicAdd(a.d_count, s);
    }
  }
}

// =====================================================================================================
// lower_kernel -- IgnoreCase front end of the filter kernel
// =====================================================================================================
// `runLower` (Automaton.hs:551-553) lower-cases every haystack code point on the fly (Utf8.lowerCodePoint,
// Utf8.hs:145-151) and matches the lowered code points against needles the caller lowered.  When a lowering
// keeps the UTF-8 length of its code point -- true for all but a handful of code points (K U+212A, Å U+212B,
// ẞ U+1E9E, İ U+0130, Ⱥ, Ⱦ, ...) -- the lowered text has the same byte offsets as the original, so the scan
// can run the CaseSensitive kernels on a lowered COPY of the text.  This kernel writes that copy: one thread
// per 16-byte granule, SWAR for all-ASCII granules, decode -> table -> re-encode otherwise.  A code point
// whose lowering changes length is handled in one of two ways:
//  * keep = 1 (the automaton holds the needle variants for every such code point, am_build.cpp step 1):
//    it is left unchanged in the copy, where the variant needles match it -- exact, no second pass;
//  * keep = 0: it is overwritten with 0xFF bytes (never part of a valid UTF-8 needle) and counted; the host
//    then falls back to the exact per-code-point walk kernel for that text.
__global__ void __launch_bounds__(256) lower_kernel(DevAutomaton A, const uint8_t* text, uint64_t text_len, uint8_t* out /* same misalignment as text */,
                                                    unsigned int* exceptions, int keep) {
  const uintptr_t addr0 = reinterpret_cast<uintptr_t>(text);
  const uint32_t a0 = (uint32_t)(addr0 & 15);
  const uint4* in16 = reinterpret_cast<const uint4*>(addr0 - a0);
  uint4* out16 = reinterpret_cast<uint4*>(out - a0);
  const uint32_t* in32 = reinterpret_cast<const uint32_t*>(in16);
  const uint64_t nvec = (a0 + text_len + 15) >> 4;
  const uint64_t lo = a0, hi = a0 + text_len;                 // virtual byte range of the text
  for (uint64_t gi = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; gi < nvec; gi += (uint64_t)gridDim.x * blockDim.x) {
    const uint4 q = ld_stream_v4(in16 + gi);
    // ASCII letters of all four words at once; an all-ASCII granule is done (no code point reaches into it)
    uint32_t R[6] = {0u, lower_ascii_word(q.x), lower_ascii_word(q.y), lower_ascii_word(q.z), lower_ascii_word(q.w), 0u};
    if (((q.x | q.y | q.z | q.w) & 0x80808080u) != 0) {
      // Window of bytes [-4, 20) of the granule as six words.  Every code point above ASCII whose lead byte lies in
      // window bytes [1, 20) and that overlaps the granule is decoded (decodeN, Utf8.hs:344-350), lowered, re-encoded
      // (unicode2utf8, :154-160).  The word index is static (unrolled), so the bytes of a code point come out of a word
      // pair by a funnel shift and go back in by a 64-bit mask -- no byte arrays, no local memory.
      const uint32_t prev = gi > 0 ? __ldg(in32 + gi * 4 - 1) : 0u;
      const uint32_t next = gi + 1 < nvec ? __ldg(in32 + (gi + 1) * 4) : 0u;
      const uint32_t W[6] = {prev, q.x, q.y, q.z, q.w, next};
      const uint64_t vbase = (gi << 4) - 4;                   // virtual index of window byte 0 (wraps for gi = 0; only used with j >= 4 there)
#pragma unroll
      for (int wi = 0; wi < 5; wi++) {
        uint32_t mw = W[wi] & (W[wi] << 1) & 0x80808080u;    // bit 7 of every byte 11xxxxxx: a lead byte of a 2..4-byte code point
        if (wi == 0) mw &= 0xFFFFFF00u;                       // window byte 0 cannot reach the granule
        while (mw) {
          const uint32_t k = (uint32_t)(__ffs((int)mw) - 1) >> 3;   // byte of the word
          mw &= mw - 1;
          const uint32_t g = __funnelshift_r(W[wi], W[wi + 1], 8 * k);   // the lead byte and the three that follow
          const uint32_t b0 = g & 0xFFu;
          uint32_t len, cp;
          if (b0 < 0xE0u) { len = 2; cp = ((b0 & 0x1Fu) << 6) | ((g >> 8) & 0x3Fu); }
          else if (b0 < 0xF0u) { len = 3; cp = ((b0 & 0x0Fu) << 12) | (((g >> 8) & 0x3Fu) << 6) | ((g >> 16) & 0x3Fu); }
          else { len = 4; cp = ((b0 & 0x07u) << 18) | (((g >> 8) & 0x3Fu) << 12) | (((g >> 16) & 0x3Fu) << 6) | ((g >> 24) & 0x3Fu); }
          const uint32_t j = 4u * wi + k;                     // window byte of the lead
          if (j + len <= 4u) continue;                        // ends before the granule
          const uint64_t v = vbase + j;
          if ((gi == 0 && j < 4u) || v < lo || v + len > hi) continue;   // outside / truncated by the text: copied through
          const uint32_t l = lower_cp(A, cp);
          if (l == cp) continue;
          uint32_t e, elen;
          if (l < 0x80u) { e = l; elen = 1; }
          else if (l < 0x800u) { e = (0xC0u | (l >> 6)) | ((0x80u | (l & 0x3Fu)) << 8); elen = 2; }
          else if (l < 0x10000u) { e = (0xE0u | (l >> 12)) | ((0x80u | ((l >> 6) & 0x3Fu)) << 8) | ((0x80u | (l & 0x3Fu)) << 16); elen = 3; }
          else { e = (0xF0u | (l >> 18)) | ((0x80u | ((l >> 12) & 0x3Fu)) << 8) | ((0x80u | ((l >> 6) & 0x3Fu)) << 16) | ((0x80u | (l & 0x3Fu)) << 24); elen = 4; }
          if (elen != len) {
            if (keep) continue;                               // stays as it is: the needle variants match it
            if (*reinterpret_cast<volatile unsigned int*>(exceptions) == 0) *exceptions = 1u;   // only zero / non-zero matters
            e = 0xFFFFFFFFu;                                  // marked: never part of a valid UTF-8 needle
          }
          const uint32_t m32 = 0xFFFFFFFFu >> (32u - 8u * len);                   // the code point's bytes, at byte k of word wi
          e &= m32;
          const uint32_t sh = 8u * k;
          R[wi] = (R[wi] & ~(m32 << sh)) | (e << sh);
          R[wi + 1] = (R[wi + 1] & ~__funnelshift_l(m32, 0u, sh)) | __funnelshift_l(e, 0u, sh);   // the part that spills into the next word
        }
      }
    }
    out16[gi] = make_uint4(R[1], R[2], R[3], R[4]);
  }
}

// =====================================================================================================
// key -> am_match
// =====================================================================================================
__global__ void unpack_kernel(const uint64_t* keys, uint64_t n, uint32_t rank_bits, const uint32_t* id_of_rank, am_match* out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t k = keys[i];
  am_match mm;
  mm.end_pos = k >> rank_bits;
  mm.needle_id = __ldg(id_of_rank + (uint32_t)(k & ((1ull << rank_bits) - 1)));
  mm.reserved = 0;
  out[i] = mm;
}

// =====================================================================================================
// launchers
// =====================================================================================================
template <bool IC, int MODE>
static cudaError_t launch_walk_t(const DevAutomaton& A, const ScanArgs& a, cudaStream_t st) {
  if (a.text_len <= a.report_begin) return cudaSuccess;
  static std::atomic<uint64_t> attr_done{0};   // per device (am_options.device: one process may use several GPUs)
  {
    cudaError_t e = ensure_dynamic_smem(walk_kernel<IC, MODE>, (int)(sizeof(WalkSmem) + WALK_HOT_BYTES_BIG), attr_done);
    if (e != cudaSuccess) return e;
  }
  const uint64_t span = a.text_len - a.report_begin;
  // small automaton: one big CTA per SM with most rows in shared memory; large: two CTAs, rows mostly via L1/L2
  const uint64_t row_bytes = 4ull << A.cdfa_shift;
  const bool small = (uint64_t)A.cdfa_states * row_bytes <= WALK_SMALL_AUTOMATON_BYTES;
  const int threads = small ? 1024 : 512;
  const uint32_t hot_budget = small ? WALK_HOT_BYTES_BIG : WALK_HOT_BYTES_SMALL;
  // segment: long enough to amortise the halo, short enough to fill the GPU
  uint64_t seg = (uint64_t)A.halo * 8; if (seg < 256) seg = 256;
  const uint64_t want = (uint64_t)sm_count() * 2048;  // segments wanted: two per resident thread
  while (seg > 64 && seg > (uint64_t)A.halo * 2 && (span + seg - 1) / seg < want) seg >>= 1;
  seg = (seg + 15) & ~15ull;
  const uint64_t nseg = (span + seg - 1) / seg;
  const uint64_t npairs = (nseg + 1) / 2;              // a thread walks two segments in lockstep
  uint64_t blocks = (npairs + threads - 1) / threads;
  const uint64_t max_blocks = (uint64_t)sm_count() * (small ? 1 : 2);
  if (blocks > max_blocks) blocks = max_blocks;
  // rows staged in shared memory: as many of the shallowest (BFS-first) states as fit, a whole number of 16-byte units
  uint32_t hot = (uint32_t)std::min<uint64_t>(A.cdfa_states, (uint64_t)hot_budget / row_bytes);
  if (A.cdfa_shift == 1) hot &= ~1u;
  g_kernel_launches++;
  walk_kernel<IC, MODE><<<(unsigned)blocks, threads, sizeof(WalkSmem) + hot_budget, st>>>(A, a, seg, nseg, hot);
  return cudaGetLastError();
}

cudaError_t launch_walk(const DevAutomaton& A, const ScanArgs& a, int mode, cudaStream_t st) {
  if (A.ignore_case) {
    if (mode == MODE_COUNT) return launch_walk_t<true, MODE_COUNT>(A, a, st);
    if (mode == MODE_ANY) return launch_walk_t<true, MODE_ANY>(A, a, st);
    return launch_walk_t<true, MODE_EMIT>(A, a, st);
  }
  if (mode == MODE_COUNT) return launch_walk_t<false, MODE_COUNT>(A, a, st);
  if (mode == MODE_ANY) return launch_walk_t<false, MODE_ANY>(A, a, st);
  return launch_walk_t<false, MODE_EMIT>(A, a, st);
}

cudaError_t launch_lower(const DevAutomaton& A, const uint8_t* text, uint64_t text_len, uint8_t* out, unsigned int* exceptions, bool keep, cudaStream_t st) {
  if (text_len == 0) return cudaSuccess;
  const uint64_t nvec = ((reinterpret_cast<uintptr_t>(text) & 15) + text_len + 15) >> 4;
  const unsigned blocks = (unsigned)std::min<uint64_t>((nvec + 255) / 256, (uint64_t)sm_count() * 16);
  g_kernel_launches++;
  lower_kernel<<<blocks, 256, 0, st>>>(A, text, text_len, out, exceptions, keep ? 1 : 0);
  return cudaGetLastError();
}

// =====================================================================================================
// segmented emission of the filter kernel -> ordered key list
// =====================================================================================================
// bases[s] = sum of min(count, seg_cap) over s' < s, for s = 0 .. num_segs (the last one is the total): one CUB scan.
struct SegClamp {
  uint32_t cap;
  __host__ __device__ unsigned long long operator()(uint32_t c) const { return c < cap ? c : cap; }
};

// One warp per segment: rank sort of its (distinct) keys straight into their final place.
constexpr int SEG_SORT_WARPS = 8;
__global__ void __launch_bounds__(SEG_SORT_WARPS * 32) seg_sort_kernel(const uint64_t* seg_keys, const uint32_t* seg_counts, const uint64_t* bases, uint64_t num_segs,
                                                                       uint32_t seg_cap, uint64_t* out, am_match* matches, uint64_t matches_cap,
                                                                       uint32_t rank_bits, const uint32_t* id_of_rank) {
  __shared__ unsigned long long sk[SEG_SORT_WARPS][SEG_CAP_MAX];
  auto place = [&](uint64_t at, unsigned long long k) {         // final position of key k (and, fused, its am_match record)
    out[at] = k;
    if (matches && at < matches_cap) {
      am_match mm;
      mm.end_pos = k >> rank_bits; mm.needle_id = __ldg(id_of_rank + (uint32_t)(k & ((1ull << rank_bits) - 1))); mm.reserved = 0;
      matches[at] = mm;
    }
  };
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint64_t s = (uint64_t)blockIdx.x * SEG_SORT_WARPS + warp; s < num_segs; s += (uint64_t)gridDim.x * SEG_SORT_WARPS) {
    uint32_t c = seg_counts[s]; if (c > seg_cap) c = seg_cap;
    if (c == 0) continue;
    const uint64_t* src = seg_keys + s * seg_cap;
    const uint64_t base = bases[s];
    if (c <= 32) {                                             // the common case: one key per lane, ranks by shuffles
      const unsigned long long k = lane < c ? src[lane] : ~0ull;
      uint32_t rank = 0;
      for (uint32_t j = 0; j < c; j++) rank += __shfl_sync(0xFFFFFFFFu, k, j) < k;
      if (lane < c) place(base + rank, k);
    } else {
      __syncwarp();
      for (uint32_t i = lane; i < c; i += 32) sk[warp][i] = src[i];
      __syncwarp();
      for (uint32_t i = lane; i < c; i += 32) {
        const unsigned long long k = sk[warp][i];
        uint32_t rank = 0;
        for (uint32_t j = 0; j < c; j++) rank += sk[warp][j] < k;
        place(base + rank, k);
      }
    }
  }
}

// Some key overflowed its segment: gather the segments and the overflow area into one unordered list for the radix sort.
__global__ void seg_compact_kernel(const uint64_t* seg_keys, const uint32_t* seg_counts, const uint64_t* bases, uint64_t num_segs, uint32_t seg_cap,
                                   const uint64_t* ovf_keys, uint64_t n_ovf, uint64_t stored_total, uint64_t* out) {
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t s = warp; s < num_segs; s += nwarps) {
    uint32_t c = seg_counts[s]; if (c > seg_cap) c = seg_cap;
    for (uint32_t i = lane; i < c; i += 32) out[bases[s] + i] = seg_keys[s * seg_cap + i];
  }
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_ovf; i += (uint64_t)gridDim.x * blockDim.x) out[stored_total + i] = ovf_keys[i];
}

size_t seg_scan_temp_bytes(uint64_t num_segs) {
  size_t bytes = 0;
  auto in = thrust::make_transform_iterator(static_cast<const uint32_t*>(nullptr), SegClamp{0});
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, (unsigned long long*)nullptr, (int64_t)num_segs + 1);
  return bytes;
}
// seg_counts holds num_segs + 1 entries (the last one zero), bases num_segs + 1: bases[num_segs] is the number of stored keys.
cudaError_t launch_seg_scan(void* temp, size_t temp_bytes, const uint32_t* seg_counts, uint64_t num_segs, uint32_t seg_cap, uint64_t* bases, cudaStream_t st) {
  auto in = thrust::make_transform_iterator(seg_counts, SegClamp{seg_cap});
  g_kernel_launches++;
  return cub::DeviceScan::ExclusiveSum(temp, temp_bytes, in, reinterpret_cast<unsigned long long*>(bases), (int64_t)num_segs + 1, st);
}
cudaError_t launch_seg_sort(const uint64_t* seg_keys, const uint32_t* seg_counts, const uint64_t* bases, uint64_t num_segs, uint32_t seg_cap, uint64_t* out,
                            am_match* matches, uint64_t matches_cap, uint32_t rank_bits, const uint32_t* id_of_rank, cudaStream_t st) {
  const unsigned blocks = (unsigned)std::min<uint64_t>((num_segs + SEG_SORT_WARPS - 1) / SEG_SORT_WARPS, (uint64_t)sm_count() * 8);
  g_kernel_launches++;
  seg_sort_kernel<<<blocks, SEG_SORT_WARPS * 32, 0, st>>>(seg_keys, seg_counts, bases, num_segs, seg_cap, out, matches, matches_cap, rank_bits, id_of_rank);
  return cudaGetLastError();
}
cudaError_t launch_seg_compact(const uint64_t* seg_keys, const uint32_t* seg_counts, const uint64_t* bases, uint64_t num_segs, uint32_t seg_cap,
                               const uint64_t* ovf_keys, uint64_t n_ovf, uint64_t stored_total, uint64_t* out, cudaStream_t st) {
  const unsigned blocks = (unsigned)std::min<uint64_t>((num_segs * 32 + 255) / 256 + (n_ovf + 255) / 256, (uint64_t)sm_count() * 8);
  g_kernel_launches++;
  seg_compact_kernel<<<blocks ? blocks : 1, 256, 0, st>>>(seg_keys, seg_counts, bases, num_segs, seg_cap, ovf_keys, n_ovf, stored_total, out);
  return cudaGetLastError();
}

// =====================================================================================================
// small helpers of the sharded calls and of containsAll
// =====================================================================================================
// The value a rank contributes to the sharded calls' all-gather: the sum of `n` scan counters, with bit 63 raised when the
// filter scan listed more survivors than its list holds (its counters are then incomplete and every rank repeats the round).
__global__ void shard_total_kernel(const unsigned long long* in, int first, int n, const unsigned long long* surv_count, unsigned long long surv_cap, unsigned long long* out) {
  unsigned long long s = 0;
  for (int i = 0; i < n; i++) s += in[first + i];
  if (surv_cap && *surv_count > surv_cap) s |= 1ull << 63;
  *out = s;
}
cudaError_t launch_shard_total(const unsigned long long* d_scalars, int first, int n, const unsigned long long* surv_count, unsigned long long surv_cap,
                               unsigned long long* d_out, cudaStream_t st) {
  g_kernel_launches++;
  shard_total_kernel<<<1, 1, 0, st>>>(d_scalars, first, n, surv_count, surv_cap, d_out);
  return cudaGetLastError();
}

// Searcher.containsAll (Searcher.hs:173-187): OR the needle ranks of a sorted key list into a bit set ...
__global__ void mark_seen_kernel(const uint64_t* keys, uint64_t n, uint32_t rank_bits, uint32_t* seen) {
  const uint64_t mask = (1ull << rank_bits) - 1;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t r = (uint32_t)(keys[i] & mask);
    const uint32_t bit = 1u << (r & 31);
    if (!(seen[r >> 5] & bit)) atomicOr(seen + (r >> 5), bit);
  }
}
// ... and count the needles that have not been seen yet.
__global__ void count_missing_kernel(const uint32_t* seen, uint32_t num_needles, unsigned int* missing) {
  __shared__ unsigned int red[32];
  const uint32_t words = (num_needles + 31) / 32;
  unsigned int local = 0;
  for (uint32_t w = threadIdx.x; w < words; w += blockDim.x) local += __popc(seen[w]);
  for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xFFFFFFFFu, local, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int s = 0;
    for (unsigned i = 0; i < blockDim.x / 32; i++) s += red[i];
    *missing = num_needles - s;
  }
}
cudaError_t launch_mark_seen(const uint64_t* keys, uint64_t n, uint32_t rank_bits, uint32_t* seen, uint32_t num_needles, unsigned int* d_missing, cudaStream_t st) {
  const unsigned blocks = (unsigned)std::min<uint64_t>((n + 255) / 256, (uint64_t)sm_count() * 8);
  g_kernel_launches += 2;
  if (n) mark_seen_kernel<<<blocks, 256, 0, st>>>(keys, n, rank_bits, seen);
  count_missing_kernel<<<1, 1024, 0, st>>>(seen, num_needles, d_missing);
  return cudaGetLastError();
}

size_t sort_temp_bytes(uint64_t n, int end_bit) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr, (int64_t)n, 0, end_bit);
  return bytes;
}

cudaError_t sort_keys(void* temp, size_t temp_bytes, const uint64_t* in, uint64_t* out, uint64_t n, int end_bit, cudaStream_t st) {
  return cub::DeviceRadixSort::SortKeys(temp, temp_bytes, in, out, (int64_t)n, 0, end_bit, st);
}

cudaError_t launch_unpack(const uint64_t* keys, uint64_t n, uint32_t rank_bits, const uint32_t* id_of_rank, am_match* out, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  g_kernel_launches++;
  unpack_kernel<<<blocks, 256, 0, st>>>(keys, n, rank_bits, id_of_rank, out);
  return cudaGetLastError();
}

}  // namespace am
