// am_replacer.cu -- Replacer.build / run / runWithLimit on the device.
//
// Semantics: src/Data/Text/AhoCorasick/Replacer.hs.  Pair i has priority -i (:101-111).  `runWithLimit`
// (:203-242) loops: scan the whole text, keep only the matches of the best priority below the threshold
// (`prependMatch` :252-260), sort them (:241), drop the ones that start inside an earlier kept match
// (`removeOverlap` :191-198), splice the replacement in (`replace` :163-180), lower the threshold, rescan.
// One pass therefore replaces the occurrences of exactly ONE needle (priorities are distinct).
//
// Device formulation of a pass (the text stays in HBM, ping-ponging between two buffers):
//   1. full scan with the scan kernels of am_kernels.cu -> sorted (end_pos, rank) keys
//   2. atomicMin over the keys' needle ids above the previous pass's needle -> this pass's needle
//   3. stream-compact that needle's matches (already ordered by start), compute their start offsets
//      (IgnoreCase: `skipCodePointsBackwards`, Utf8.hs:256-276, over `lenCodePoints` of the ORIGINAL needle)
//   4. `replacementLength` (:183-187) over ALL of them, before overlap removal (:240)
//   5. removeOverlap: cluster heads (no overlap with the predecessor) are always kept; one thread walks
//      each cluster greedily
//   6. exclusive scan of the length deltas, then a tile-parallel splice copy into the other buffer
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>

#include <cstring>

#include "am_api_internal.h"

using namespace am;

struct am_replacer {
  am_automaton* automaton = nullptr;     // over the needles as `build` stores them: lowered iff built with IgnoreCase (:105-107)
  int built_cs = AM_CASE_SENSITIVE;
  uint64_t n = 0;
  std::vector<uint32_t> len_bytes, len_cps, repl_off;   // Payload of needle i: lengths of the ORIGINAL needle (:111-113)
  std::vector<uint8_t> repl_bytes;
  uint8_t* d_repl = nullptr;
  void* d_idinfo[2] = {nullptr, nullptr};  // per case mode: IdInfo of every needle id (rank in that image, replacement, payload lengths)
  std::mutex mu;
  std::vector<void*> scratch_pool;         // idle ReplacerScratch sets (one per concurrent run)
  bool has_empty = false;
  bool stored_len_differs = false;       // some stored (lowered) needle has another byte length than the original
};

namespace {

struct DevBuf {
  void* p = nullptr; size_t cap = 0;
  int ensure(size_t bytes) {
    if (cap >= bytes) return AM_OK;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    if (cudaMalloc(&p, want) != cudaSuccess) { cudaGetLastError(); return fail(AM_E_OOM, "cudaMalloc(replacer scratch)"); }
    cap = want; return AM_OK;
  }
  ~DevBuf() { if (p) cudaFree(p); }
  template <class T> T* as() { return static_cast<T*>(p); }
};

// Device scratch of one Replacer run, kept with the replacer between runs (a run on a 2 GiB text holds ~7 GiB of tiles, lists
// and buffers: allocating and freeing them per call costs more than the passes).
struct ReplacerScratch {
  DevBuf contig, tiles, tile_len, tile_base[2], tile_flag, touched, new_len, sel, starts, ends, keep, kidx, kstart, kend, kdelta, kshift, cubtmp, scal, spare, keys_c, rkeys;
};

struct RankIs {
  uint64_t mask; uint32_t rank;
  __host__ __device__ bool operator()(const uint64_t& k) const { return (uint32_t)(k & mask) == rank; }
};

__global__ void best_needle_kernel(const uint64_t* keys, uint64_t n, uint64_t mask, const uint32_t* id_of_rank, long long prev_id, unsigned int* best) {
  unsigned int local = 0xFFFFFFFFu;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t id = __ldg(id_of_rank + (uint32_t)(keys[i] & mask));
    if ((long long)id > prev_id && id < local) local = id;
  }
  for (int o = 16; o > 0; o >>= 1) { unsigned int v = __shfl_down_sync(0xFFFFFFFFu, local, o); if (v < local) local = v; }
  if ((threadIdx.x & 31) == 0 && local != 0xFFFFFFFFu) atomicMin(best, local);
}

// start / length of every match of the pass's needle + sum of the length deltas (replacementLength)
__global__ void starts_kernel(const uint64_t* sel, uint64_t n, uint32_t rank_bits, const uint8_t* text, int ignore_case,
                              uint32_t len_bytes, uint32_t len_cps, long long repl_len, uint64_t* start, uint64_t* end,
                              long long* delta_sum, int* error) {
  long long local = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t pos = sel[i] >> rank_bits;
    uint64_t s;
    if (!ignore_case) {
      if (pos < len_bytes) { *error = 1; s = 0; }            // (only a replacer switched to the other case mode after build can get here)
      else s = pos - len_bytes;                              // makeMatch CaseSensitive (:268-269)
    } else {
      // makeMatch IgnoreCase (:271-274): skipCodePointsBackwards haystack (pos - 1) (lenc - 1)
      long long idx = (long long)pos - 1, k = (long long)len_cps - 1;
      for (;;) {
        if (idx >= 0 && (text[idx] & 0xC0) == 0x80) { idx--; continue; }
        if (k == 0 || idx < 0) break;
        idx--; k--;
      }
      if (idx < 0) { *error = 1; idx = 0; }
      s = (uint64_t)idx;
    }
    start[i] = s; end[i] = pos;
    local += repl_len - (long long)(pos - s);
  }
  for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xFFFFFFFFu, local, o);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd((unsigned long long*)delta_sum, (unsigned long long)local);
}

// removeOverlap (:191-198).  A match that does not overlap its predecessor is kept no matter what happened
// before it, so it heads an independent cluster; inside a cluster the greedy rule is applied serially.
__global__ void overlap_kernel(const uint64_t* start, const uint64_t* end, uint64_t n, uint8_t* keep) {
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (uint64_t)gridDim.x * blockDim.x) {
    if (j > 0 && start[j] < end[j - 1]) continue;            // not a cluster head
    keep[j] = 1;
    uint64_t last_end = end[j];
    for (uint64_t k = j + 1; k < n; k++) {
      if (start[k] >= end[k - 1]) break;                      // next cluster
      if (start[k] >= last_end) { keep[k] = 1; last_end = end[k]; } else keep[k] = 0;
    }
  }
}

__global__ void gather_kept_kernel(const uint64_t* start, const uint64_t* end, const uint8_t* keep, const uint64_t* kept_index /*exclusive scan of keep*/,
                                   uint64_t n, long long repl_len, uint64_t* k_start, uint64_t* k_end, long long* k_delta) {
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (uint64_t)gridDim.x * blockDim.x) {
    if (!keep[j]) continue;
    const uint64_t k = kept_index[j];
    k_start[k] = start[j]; k_end[k] = end[j]; k_delta[k] = repl_len - (long long)(end[j] - start[j]);
  }
}

// `replace` (:163-180) as a tile-parallel splice: each CTA owns SPLICE_TILE source bytes, finds the kept
// matches that intersect it and copies the gaps between them, shifted by the prefix sum of the deltas.
// The CTA that owns a match's first byte (or, for an empty match, its position) writes the replacement.
constexpr int SPLICE_TILE = 1 << 16;

// CTA-cooperative copy of n bytes with arbitrary relative misalignment: 16-byte aligned stores; the source is read
// as aligned 16-byte vectors (vector i + 1 is the next thread's vector i: served by L1) and re-aligned with funnel
// shifts.  May read up to 31 bytes beyond s + n (every text buffer carries 64 bytes of slack).
template <int WS>
__device__ __forceinline__ uint4 splice_realign(const uint4& a, const uint4& b, uint32_t bs) {
  const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  return make_uint4(__funnelshift_r(w[WS], w[WS + 1], bs), __funnelshift_r(w[WS + 1], w[WS + 2], bs),
                    __funnelshift_r(w[WS + 2], w[WS + 3], bs), __funnelshift_r(w[WS + 3], w[WS + 4], bs));
}
template <int WS>
__device__ __forceinline__ void splice_copy_vec(uint4* dv, const uint4* sv, uint64_t nvec, uint32_t bs) {
#pragma unroll 2
  for (uint64_t i = threadIdx.x; i < nvec; i += blockDim.x) dv[i] = splice_realign<WS>(__ldg(sv + i), __ldg(sv + i + 1), bs);
}
__device__ __forceinline__ void splice_copy(uint8_t* d, const uint8_t* s, uint64_t n) {
  // head: bytes until d is 16-byte aligned
  uint64_t head = (16 - (reinterpret_cast<uintptr_t>(d) & 15)) & 15;
  if (head > n) head = n;
  for (uint64_t x = threadIdx.x; x < head; x += blockDim.x) d[x] = s[x];
  d += head; s += head; n -= head;
  const uint64_t nvec = n >> 4;
  const uint32_t r = (uint32_t)(reinterpret_cast<uintptr_t>(s) & 15);
  const uint4* sv = reinterpret_cast<const uint4*>(s - r);   // aligned vectors; vector i covers source bytes [16 i - r, 16 i - r + 16)
  uint4* dv = reinterpret_cast<uint4*>(d);
  if (r == 0) {
#pragma unroll 2
    for (uint64_t i = threadIdx.x; i < nvec; i += blockDim.x) dv[i] = __ldg(sv + i);
  } else {
    const uint32_t bs = (r & 3) * 8;
    switch (r >> 2) {                                          // CTA-uniform
      case 0: splice_copy_vec<0>(dv, sv, nvec, bs); break;
      case 1: splice_copy_vec<1>(dv, sv, nvec, bs); break;
      case 2: splice_copy_vec<2>(dv, sv, nvec, bs); break;
      default: splice_copy_vec<3>(dv, sv, nvec, bs); break;
    }
  }
  for (uint64_t x = (nvec << 4) + threadIdx.x; x < n; x += blockDim.x) d[x] = s[x];   // tail
}
__global__ void __launch_bounds__(256) splice_kernel(const uint8_t* src, uint64_t src_len, uint8_t* dst, const uint64_t* k_start, const uint64_t* k_end,
                                                     const long long* k_shift /*exclusive prefix of deltas*/, uint64_t K, const uint8_t* repl, uint32_t repl_len,
                                                     long long total_shift) {
  const uint64_t t0 = (uint64_t)blockIdx.x * SPLICE_TILE;
  const uint64_t t1 = t0 + SPLICE_TILE < src_len ? t0 + SPLICE_TILE : src_len;
  const bool last_tile = t1 == src_len;
  // first kept match whose END is > t0, or whose start is >= t0 (covers empty matches at t0)
  uint64_t lo = 0, hi = K;
  while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (k_end[mid] > t0 || k_start[mid] >= t0) hi = mid; else lo = mid + 1; }
  uint64_t k = lo;
  uint64_t cur = t0;                                        // next source byte to place
  // a match that began in an earlier tile and extends into this one swallows our first bytes
  if (k < K && k_start[k] < t0) { cur = k_end[k] < t1 ? k_end[k] : t1; k++; }
  for (;;) {
    const bool have = k < K && (k_start[k] < t1 || (last_tile && k_start[k] == t1));
    const uint64_t gap_end = have ? k_start[k] : t1;
    if (gap_end > cur) {                                    // copy source [cur, gap_end)
      const long long shift = k < K ? k_shift[k] : total_shift;
      splice_copy(dst + ((long long)cur + shift), src + cur, gap_end - cur);
    }
    if (!have) break;
    const long long at = (long long)k_start[k] + k_shift[k];
    for (uint32_t r = threadIdx.x; r < repl_len; r += blockDim.x) dst[at + r] = repl[r];
    cur = k_end[k] < t1 ? k_end[k] : t1;
    if (k_end[k] > t1) break;                               // the rest of the tile is inside this match
    k++;
  }
}


// ---- incremental passes -----------------------------------------------------------------
// After a pass has replaced the kept occurrences of one needle, the matches of the NEW text are
//   (a) the old matches that do not touch a replaced span, moved by the sum of the length deltas to their left, and
//   (b) matches that contain at least one byte of a replacement, or straddle the junction a deletion left behind;
//       they lie within max_len - 1 bytes of an edit.
// So instead of rescanning the whole text (`go p $ replace ...`, Replacer.hs:242) the sorted match list is carried from
// pass to pass: (a) is a filter + shift over the list, (b) a walk of the byte automaton over one small window per edit.
// Matches of needles at or above the threshold can never be used again (`pMatch < threshold`, :253) and are dropped.
// CaseSensitive, no empty needle (spans are `lenBytes` long and non-empty); other replacers rescan.

// Append `key` for the calling lanes (any subset of a warp): one atomic per warp.
__device__ __forceinline__ void append_key(uint64_t* out, unsigned long long* counter, uint64_t cap, uint64_t key) {
  const unsigned active = __activemask();
  const unsigned lane = threadIdx.x & 31;
  const int leader = __ffs(active) - 1;
  unsigned long long base = 0;
  if ((int)lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(active));
  base = __shfl_sync(active, base, leader);
  const unsigned long long slot = base + __popc(active & ((1u << lane) - 1u));
  if (slot < cap) out[slot] = key;
}

// (a): keep the matches of still-eligible needles that touch no replaced span [k_start, k_end), shifted.
__global__ void carry_matches_kernel(const uint64_t* keys, uint64_t n, uint32_t rank_bits, const uint32_t* id_of_rank, const uint32_t* len_of_rank,
                                     uint32_t pass_id, const uint64_t* k_start, const uint64_t* k_end, const long long* k_shift, uint64_t K,
                                     long long total_shift, uint64_t* out, unsigned long long* counter, uint64_t cap) {
  const uint64_t mask = (1ull << rank_bits) - 1;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t key = keys[i];
    const uint32_t rank = (uint32_t)(key & mask);
    if (__ldg(id_of_rank + rank) <= pass_id) continue;        // at or above the new threshold
    const uint64_t e = key >> rank_bits, s = e - __ldg(len_of_rank + rank);
    uint64_t lo = 0, hi = K;                                  // first edit that ends after the match starts
    while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (k_end[mid] > s) hi = mid; else lo = mid + 1; }
    if (lo < K && k_start[lo] < e) continue;                  // touches a replaced span
    const long long shift = lo < K ? k_shift[lo] : total_shift;
    append_key(out, counter, cap, ((uint64_t)((long long)e + shift) << rank_bits) | rank);
  }
}

// (b): one thread per edit walks [new_start - (L - 1), new_start + repl_len + (L - 1)) of the NEW text from the root
// state and reports the matches that touch its edit; a match that touches several edits is reported by the last one.
__global__ void rescan_edits_kernel(DevAutomaton A, const uint8_t* text, uint64_t text_len, const uint64_t* k_start, const long long* k_shift, uint64_t K,
                                    uint32_t repl_len, uint32_t pass_id, uint64_t* out, unsigned long long* counter, uint64_t cap) {
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < K; j += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t ns = (uint64_t)((long long)k_start[j] + k_shift[j]);          // the replacement occupies [ns, ns + repl_len)
    const bool has_next = j + 1 < K;
    const uint64_t ns2 = has_next ? (uint64_t)((long long)k_start[j + 1] + k_shift[j + 1]) : 0;
    const uint64_t halo = A.max_len - 1;
    uint64_t p = ns > halo ? ns - halo : 0;
    uint64_t end = ns + repl_len + halo; if (end > text_len) end = text_len;
    uint32_t state = 0;
    for (; p < end; p++) {
      const uint32_t t = ac_step(A, state, __ldg(text + p));
      state = t & ID_MASK;
      if (!(t & OUT_FLAG)) continue;
      const uint64_t e = p + 1;
      for (uint32_t c = __ldg(A.first_out + state); c != NONE; c = __ldg(A.next_out + c)) {
        const uint32_t olo = __ldg(A.own_off + c), ohi = __ldg(A.own_off + c + 1);
        for (uint32_t o = olo; o < ohi; o++) {
          const uint32_t rank = __ldg(A.own_rank + o);
          if (__ldg(A.id_of_rank + rank) <= pass_id) continue;
          const uint64_t s = e - __ldg(A.len_of_rank + rank);
          // touches edit j: holds a replacement byte, or (deletion) spans the junction
          const bool mine = repl_len ? (s < ns + repl_len && e > ns) : (s < ns && e > ns);
          if (!mine) continue;
          if (has_next && (repl_len ? (s < ns2 + repl_len && e > ns2) : (s < ns2 && e > ns2))) continue;   // the next edit reports it
          append_key(out, counter, cap, (e << A.rank_bits) | rank);
        }
      }
    }
  }
}

}  // namespace

// =====================================================================================================================
// Passes that cost O(edits), not O(text): the text lives in TILES
// =====================================================================================================================
// A pass of `runWithLimit` replaces the kept occurrences of ONE needle: a few thousand edits in a multi-GiB text.  Copying
// the whole text per pass (`replace`, :163-180, is a Text.concat) made the passes of config C4 cost 2.9 ms each, all of it
// the copy.  Here the text is cut into tiles of TILE_FILL bytes that sit in slots of TILE_CAP bytes, so a tile can grow or
// shrink in place: a pass rewrites only the tiles its edits touch, the per-tile lengths are prefix-summed again
// (tile_base), and every kernel that reads text does so through a TextView that maps a text position to (tile, offset).
// The match list is carried from pass to pass in text coordinates as before: matches that touch no edit are shifted,
// the neighbourhood of every edit is rescanned on the rewritten tiles.  A tile that would outgrow its slot sends that one
// pass down the contiguous path (materialise, splice_kernel, re-tile).  All kernels of a pass are queued back to back
// -- the needle of the pass, its match count and the number of kept edits stay on the device (PassScalars) -- and the host
// waits ONCE per pass.
constexpr uint32_t TILE_FILL = 4096, TILE_CAP = 8192;

struct TextView {
  const uint8_t* base;          // the contiguous text, or the tile slots (tile t at base + t * TILE_CAP)
  uint64_t len;
  const uint32_t* tile_len;     // tiled: bytes in tile t ...
  const uint64_t* tile_base;    // ... and the text position of its first byte (num_tiles + 1 entries, the last one = len)
  uint32_t num_tiles, tiled;
};
struct Cursor { uint64_t pos; uint32_t t, off; };

__device__ __forceinline__ void cur_set(const TextView& v, Cursor& c, uint64_t pos) {   // pos < v.len
  c.pos = pos;
  if (v.tiled) {
    uint32_t lo = 0, hi = v.num_tiles;                        // the last tile whose base is <= pos: it is not empty and holds pos
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(v.tile_base + mid) <= pos) lo = mid; else hi = mid; }
    c.t = lo; c.off = (uint32_t)(pos - __ldg(v.tile_base + lo));
  }
}
__device__ __forceinline__ uint32_t cur_byte(const TextView& v, const Cursor& c) {
  return v.tiled ? __ldg(v.base + (size_t)c.t * TILE_CAP + c.off) : __ldg(v.base + c.pos);
}
__device__ __forceinline__ void cur_next(const TextView& v, Cursor& c) {                 // may step to v.len (then the cursor must not be read)
  c.pos++;
  if (v.tiled) {
    c.off++;
    while (c.off >= __ldg(v.tile_len + c.t) && c.t + 1 < v.num_tiles) { c.t++; c.off = 0; }
  }
}
__device__ __forceinline__ void cur_prev(const TextView& v, Cursor& c) {                 // pos > 0
  c.pos--;
  if (v.tiled) {
    if (c.off > 0) c.off--;
    else { do { c.t--; } while (__ldg(v.tile_len + c.t) == 0); c.off = __ldg(v.tile_len + c.t) - 1; }
  }
}
// skipCodePointsBackwards (Utf8.hs:256-276) from byte `pos` (inside a code point) back by n more code points; -1 where the
// reference calls `error`.
__device__ __forceinline__ long long skip_back(const TextView& v, uint64_t pos, long long n) {
  Cursor c; cur_set(v, c, pos);
  for (;;) {
    if ((cur_byte(v, c) & 0xC0u) == 0x80u) { if (c.pos == 0) return -1; cur_prev(v, c); continue; }
    if (n == 0) return (long long)c.pos;
    if (c.pos == 0) return -1;
    cur_prev(v, c); n--;
  }
}

struct IdInfo { uint32_t rank, repl_off, repl_len, len_bytes, len_cps; };
struct PassScalars {
  unsigned int best;            // needle id of this pass (NONE: no match below the threshold)
  int error;                    // skipCodePointsBackwards ran off the text
  unsigned int overflow;        // some tile would outgrow its slot: nothing was rewritten
  unsigned int ntouched;
  long long delta_sum;          // replacementLength - text length over ALL matches of the needle, before removeOverlap (:240)
  unsigned long long nsel;      // matches of the needle
  unsigned long long nkept;     // after removeOverlap
  long long total_shift;        // sum of the kept deltas
  unsigned long long n_new;     // matches found around the edits of the rewritten text
  unsigned long long n_carried; // matches carried over
  IdInfo id;                    // of `best`
};

__global__ void pass_begin_kernel(PassScalars* sc) { PassScalars z; memset(&z, 0, sizeof z); z.best = NONE; *sc = z; }
__global__ void pass_resolve_kernel(PassScalars* sc, const IdInfo* info) { if (sc->best != NONE) sc->id = info[sc->best]; }

struct RankIsDev {
  uint64_t mask; const PassScalars* sc;
  __device__ bool operator()(const uint64_t& k) const { return sc->best != NONE && (uint32_t)(k & mask) == sc->id.rank; }
};

// start / length of every match of the pass's needle + sum of the length deltas (replacementLength); `n` lives on the device
__global__ void starts_view_kernel(const uint64_t* sel, const PassScalars* sc, uint32_t rank_bits, TextView v, int ignore_case, uint64_t* start, uint64_t* end,
                                   PassScalars* out) {
  const uint64_t n = sc->nsel;
  const uint32_t len_bytes = sc->id.len_bytes, len_cps = sc->id.len_cps;
  const long long repl_len = sc->id.repl_len;
  long long local = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t pos = sel[i] >> rank_bits;
    uint64_t s;
    if (!ignore_case) {
      if (pos < len_bytes) { out->error = 1; s = 0; }
      else s = pos - len_bytes;                              // makeMatch CaseSensitive (:268-269)
    } else {
      const long long idx = skip_back(v, pos - 1, (long long)len_cps - 1);   // makeMatch IgnoreCase (:271-274)
      if (idx < 0) { out->error = 1; s = 0; } else s = (uint64_t)idx;
    }
    start[i] = s; end[i] = pos;
    local += repl_len - (long long)(pos - s);
  }
  for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xFFFFFFFFu, local, o);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd((unsigned long long*)&out->delta_sum, (unsigned long long)local);
}

// removeOverlap (:191-198) with the count on the device.  Cluster heads are always kept; one WARP walks a cluster: its
// lanes look 32 matches ahead at a time (a periodic text makes the whole list one cluster).
__global__ void overlap_dev_kernel(const uint64_t* start, const uint64_t* end, const PassScalars* sc, uint8_t* keep) {
  const uint64_t n = sc->nsel;
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  // every warp takes the cluster heads among "its" 32-match slices
  for (uint64_t base = warp * 32; base < n; base += nwarps * 32) {
    const uint64_t j = base + lane;
    const bool head = j < n && (j == 0 || start[j] >= end[j - 1]);
    unsigned heads = __ballot_sync(0xFFFFFFFFu, head);
    while (heads) {
      const int hl = __ffs(heads) - 1;
      heads &= heads - 1;
      uint64_t k = base + hl;                                 // walk this cluster with the whole warp
      if (lane == 0) keep[k] = 1;
      uint64_t last_end = end[k];
      k++;
      for (;;) {
        // the next kept match: the first k' >= k with start >= last_end, unless the cluster ends before it
        const uint64_t kk = k + lane;
        const bool in = kk < n;
        const uint64_t s = in ? start[kk] : 0, e_prev = in ? end[kk - 1] : 0;
        const unsigned brk = __ballot_sync(0xFFFFFFFFu, !in || s >= e_prev);     // cluster boundary (or the end of the list) at kk
        const unsigned ok = __ballot_sync(0xFFFFFFFFu, in && s >= last_end);
        const int b = brk ? __ffs(brk) - 1 : 32, o = ok ? __ffs(ok) - 1 : 32;
        if (o < b) {                                          // a kept match inside the cluster
          if (lane == 0) keep[k + o] = 1;
          last_end = end[k + o];
          k += o + 1;
        } else if (b < 32) break;                             // the cluster ends first
        else k += 32;                                         // neither in these 32: look further
      }
    }
  }
}

__global__ void gather_kept_dev_kernel(const uint64_t* start, const uint64_t* end, const uint8_t* keep, const uint64_t* kept_index, uint64_t n_bound,
                                       PassScalars* sc, uint64_t* k_start, uint64_t* k_end, long long* k_delta) {
  const uint64_t n = sc->nsel;
  const long long repl_len = sc->id.repl_len;
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (uint64_t)gridDim.x * blockDim.x) {
    if (!keep[j]) continue;
    const uint64_t k = kept_index[j];
    k_start[k] = start[j]; k_end[k] = end[j]; k_delta[k] = repl_len - (long long)(end[j] - start[j]);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) sc->nkept = kept_index[n_bound];
}
__global__ void total_shift_kernel(const long long* k_shift, uint64_t n_bound, PassScalars* sc) { sc->total_shift = k_shift[n_bound]; }

// (a) of the carried list, on the text BEFORE the pass's edits: the matches of still-eligible needles that touch no
// replaced span, shifted; every other key becomes DROPPED and an order-preserving compaction (DeviceSelect) follows, so
// the carried list stays sorted and no pass sorts it again.  IgnoreCase: the span of a match starts len_cps code points
// before its end; it is found by walking back only when an edit lies close enough to matter.
constexpr uint64_t DROPPED = ~0ull;
struct NotDropped { __host__ __device__ bool operator()(const uint64_t& k) const { return k != DROPPED; } };
__global__ void carry_view_kernel(const uint64_t* keys, uint64_t n, uint32_t rank_bits, const uint32_t* id_of_rank, const uint32_t* len_of_rank, const IdInfo* info,
                                  int ignore_case, TextView v, const uint64_t* k_start, const uint64_t* k_end, const long long* k_shift, uint64_t n_bound,
                                  PassScalars* sc, uint64_t* out) {
  const uint64_t mask = (1ull << rank_bits) - 1;
  const uint64_t K = sc->nkept;
  const uint32_t pass_id = sc->best;
  const long long total_shift = k_shift[n_bound];
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t key = keys[i];
    const uint32_t rank = (uint32_t)(key & mask);
    const uint32_t id = __ldg(id_of_rank + rank);
    uint64_t res = DROPPED;
    if (id > pass_id) {                                        // (else: at or above the new threshold, `pMatch < threshold`, :253)
      const uint64_t e = key >> rank_bits;
      uint64_t s;
      if (!ignore_case) s = e - __ldg(len_of_rank + rank);
      else { const uint64_t span = 4ull * __ldg(&info[id].len_cps); s = e > span ? e - span : 0; }   // earliest possible start
      uint64_t lo = 0, hi = K;                                // first edit that ends after the match starts
      while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (k_end[mid] > s) hi = mid; else lo = mid + 1; }
      bool touched = lo < K && k_start[lo] < e;
      if (touched && ignore_case) {
        const long long es = skip_back(v, e - 1, (long long)__ldg(&info[id].len_cps) - 1);   // the exact start
        if (es < 0) sc->error = 1;
        else {
          s = (uint64_t)es;
          while (lo < K && k_end[lo] <= s) lo++;
          touched = lo < K && k_start[lo] < e;
        }
      }
      if (!touched) res = ((uint64_t)((long long)e + (lo < K ? k_shift[lo] : total_shift)) << rank_bits) | rank;
    }
    out[i] = res;
  }
}
// Merge the (sorted) carried list A with the (sorted, small) list B of rescanned matches: the keys are distinct, so the
// place of a key is its own index plus the number of keys of the other list below it.
__global__ void merge_kernel(const uint64_t* A, uint64_t nA, const uint64_t* B, uint64_t nB, uint64_t* out) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nA + nB; i += (uint64_t)gridDim.x * blockDim.x) {
    const bool fromA = i < nA;
    const uint64_t key = fromA ? A[i] : B[i - nA];
    const uint64_t* o = fromA ? B : A;
    uint64_t lo = 0, hi = fromA ? nB : nA;
    while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (o[mid] < key) lo = mid + 1; else hi = mid; }
    out[(fromA ? i : i - nA) + lo] = key;
  }
}

// ---- tiles -----------------------------------------------------------------------------------------------------------------
__global__ void tileify_kernel(const uint8_t* src, uint64_t len, uint8_t* tiles, uint32_t* tile_len, uint32_t num_tiles) {
  // 16-byte copies where the source allows it (tile t starts at src + t * 4096: as aligned as src itself)
  const bool al = (reinterpret_cast<uintptr_t>(src) & 15) == 0;
  for (uint32_t t = blockIdx.x; t < num_tiles; t += gridDim.x) {
    const uint64_t b = (uint64_t)t * TILE_FILL;
    const uint32_t n = (uint32_t)(len - b < TILE_FILL ? len - b : TILE_FILL);
    uint8_t* d = tiles + (size_t)t * TILE_CAP;
    if (al) {
      const uint32_t nv = n >> 4;
      for (uint32_t i = threadIdx.x; i < nv; i += blockDim.x) reinterpret_cast<uint4*>(d)[i] = __ldg(reinterpret_cast<const uint4*>(src + b) + i);
      for (uint32_t i = (nv << 4) + threadIdx.x; i < n; i += blockDim.x) d[i] = src[b + i];
    } else {
      for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) d[i] = src[b + i];
    }
    if (threadIdx.x == 0) tile_len[t] = n;
  }
}
__global__ void materialize_kernel(const uint8_t* tiles, const uint32_t* tile_len, const uint64_t* tile_base, uint32_t num_tiles, uint8_t* dst) {
  for (uint32_t t = blockIdx.x; t < num_tiles; t += gridDim.x) {
    const uint32_t n = tile_len[t];
    const uint8_t* s = tiles + (size_t)t * TILE_CAP;
    uint8_t* d = dst + tile_base[t];
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) d[i] = s[i];
  }
}
// The tiles an edit touches (the tile of its first byte .. the tile of its last byte), each listed once.
__global__ void tile_mark_kernel(const PassScalars* sc, const uint64_t* k_start, const uint64_t* k_end, TextView v, uint32_t* tile_flag, uint32_t* touched, PassScalars* out) {
  const uint64_t K = sc->nkept;
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < K; j += (uint64_t)gridDim.x * blockDim.x) {
    Cursor a, b;
    cur_set(v, a, k_start[j]);
    cur_set(v, b, k_end[j] - 1);                               // (spans are not empty: no empty needle in this form)
    for (uint32_t t = a.t; t <= b.t; t++)
      if (atomicExch(tile_flag + t, 1u) == 0u) touched[atomicAdd(&out->ntouched, 1u)] = t;
  }
}
// New length of every touched tile; a tile that would outgrow its slot raises `overflow` and the pass rewrites nothing.
// One thread per touched tile walks the (few) edits that intersect it.
__device__ __forceinline__ uint64_t first_edit_after(const uint64_t* k_end, uint64_t K, uint64_t t0) {
  uint64_t lo = 0, hi = K;
  while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (k_end[mid] > t0) hi = mid; else lo = mid + 1; }
  return lo;
}
__global__ void tile_plan_kernel(PassScalars* sc, const uint32_t* touched, const uint64_t* k_start, const uint64_t* k_end, TextView v, uint32_t* new_len) {
  const uint64_t K = sc->nkept;
  const uint32_t rl = sc->id.repl_len, nt = sc->ntouched;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nt; i += gridDim.x * blockDim.x) {
    const uint32_t t = touched[i];
    const uint64_t t0 = v.tile_base[t], t1 = t0 + v.tile_len[t];
    long long len = (long long)v.tile_len[t];
    for (uint64_t k = first_edit_after(k_end, K, t0); k < K && k_start[k] < t1; k++) {
      const uint64_t a = k_start[k] > t0 ? k_start[k] : t0, b = k_end[k] < t1 ? k_end[k] : t1;
      len -= (long long)(b - a);
      if (k_start[k] >= t0) len += rl;                         // the tile that holds an edit's first byte writes its replacement
    }
    new_len[i] = (uint32_t)len;
    if (len > (long long)TILE_CAP) sc->overflow = 1;
  }
}
// Rewrite the touched tiles in place: the tile is staged in shared memory, the gaps between its edits are copied back
// shifted, the replacements written in (`replace`, :163-180, one tile at a time).
__global__ void __launch_bounds__(128) tile_splice_kernel(const PassScalars* sc, const uint32_t* touched, const uint32_t* new_len, const uint64_t* k_start, const uint64_t* k_end,
                                                          uint8_t* tiles, uint32_t* tile_len, const uint64_t* tile_base, uint32_t* tile_flag, const uint8_t* repl_pool) {
  __shared__ __align__(16) uint8_t in[TILE_CAP];
  if (sc->overflow) {                                          // (the flags still have to be cleared for the next pass)
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < sc->ntouched; i += gridDim.x * blockDim.x) tile_flag[touched[i]] = 0;
    return;
  }
  const uint64_t K = sc->nkept;
  const uint32_t rl = sc->id.repl_len, nt = sc->ntouched;
  const uint8_t* repl = repl_pool + sc->id.repl_off;
  for (uint32_t i = blockIdx.x; i < nt; i += gridDim.x) {
    const uint32_t t = touched[i];
    const uint32_t n = tile_len[t];
    const uint64_t t0 = tile_base[t], t1 = t0 + n;
    uint8_t* tile = tiles + (size_t)t * TILE_CAP;
    __syncthreads();
    for (uint32_t x = threadIdx.x; x < (n + 15) / 16; x += blockDim.x) reinterpret_cast<uint4*>(in)[x] = reinterpret_cast<const uint4*>(tile)[x];
    __syncthreads();
    uint64_t k = first_edit_after(k_end, K, t0);
    uint64_t cur = t0;                                         // next source byte to place
    uint32_t out = 0;
    if (k < K && k_start[k] < t0) { cur = k_end[k] < t1 ? k_end[k] : t1; k++; }   // an edit that began in an earlier tile swallows our first bytes
    for (;;) {
      const bool have = k < K && k_start[k] < t1;
      const uint64_t gap_end = have ? k_start[k] : t1;
      const uint32_t g = (uint32_t)(gap_end - cur), from = (uint32_t)(cur - t0);
      for (uint32_t x = threadIdx.x; x < g; x += blockDim.x) tile[out + x] = in[from + x];
      out += g;
      if (!have) break;
      for (uint32_t x = threadIdx.x; x < rl; x += blockDim.x) tile[out + x] = repl[x];
      out += rl;
      cur = k_end[k] < t1 ? k_end[k] : t1;
      if (k_end[k] >= t1) break;                               // the rest of the tile is inside this edit
      k++;
    }
    if (threadIdx.x == 0) { tile_len[t] = new_len[i]; tile_flag[t] = 0; }
  }
}

// (b) of the carried list, on the REWRITTEN text: one thread per edit walks [new_start - halo, new_start + repl_len + halo) from
// the root state and reports the matches that touch its edit; a match that touches several edits is reported by the last
// one.  IgnoreCase walks lowered code points (decode, `Char.toLower` table, re-encode; a code point whose lower case has
// another UTF-8 length is fed as it is: the automaton holds the variants) and finds a match's start by walking back.
__global__ void rescan_view_kernel(DevAutomaton A, TextView v, const IdInfo* info, int ignore_case, const uint64_t* k_start, const long long* k_shift,
                                   PassScalars* sc, uint64_t* out, uint64_t cap) {   // out: the pass's list of rescanned matches (unordered), counted in sc->n_new
  if (sc->overflow || sc->best == NONE) return;
  const uint64_t K = sc->nkept;
  const uint32_t repl_len = sc->id.repl_len, pass_id = sc->best;
  const uint64_t text_len = v.tiled ? v.tile_base[v.num_tiles] : v.len;   // (a pass is queued before the host knows the new length)
  const uint64_t halo = A.halo;
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < K; j += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t ns = (uint64_t)((long long)k_start[j] + k_shift[j]);          // the replacement occupies [ns, ns + repl_len)
    const bool has_next = j + 1 < K;
    const uint64_t ns2 = has_next ? (uint64_t)((long long)k_start[j + 1] + k_shift[j + 1]) : 0;
    uint64_t p = ns > halo ? ns - halo : 0;
    uint64_t end = ns + repl_len + halo; if (end > text_len) end = text_len;
    if (p >= end) continue;
    Cursor c; cur_set(v, c, p);
    uint32_t state = 0, cp = 0, rem = 0, raw = 0, nraw = 0;
    bool synced = !(ignore_case && p > 0);
    for (; p < end; p++, cur_next(v, c)) {
      const uint32_t byte = cur_byte(v, c);
      uint32_t t;
      if (!ignore_case) {
        t = ac_step(A, state, byte);
      } else {
        if (!synced) { if ((byte & 0xC0u) == 0x80u) continue; synced = true; }   // start on a code point boundary
        if (rem == 0 && byte < 0x80u) {
          t = ac_step(A, state, byte + ((byte - 'A' < 26u) ? 0x20u : 0u));
        } else {
          if (rem == 0) { raw = 0; nraw = 0; cp = byte < 0xE0u ? byte & 0x1Fu : byte < 0xF0u ? byte & 0x0Fu : byte & 0x07u; rem = byte < 0xC0u ? 0u : byte < 0xE0u ? 1u : byte < 0xF0u ? 2u : 3u; if (byte < 0xC0u) cp = byte; }
          else { cp = (cp << 6) | (byte & 0x3Fu); rem--; }
          raw |= byte << (8 * nraw); nraw++;
          if (rem != 0) continue;                              // the code point is not complete yet
          uint32_t l = lower_cp(A, cp);
          const uint32_t ln = l < 0x80u ? 1u : l < 0x800u ? 2u : l < 0x10000u ? 3u : 4u;
          uint32_t enc = raw;
          if (l != cp && ln == nraw) {
            enc = ln == 1 ? l : ln == 2 ? (0xC0u | (l >> 6)) | ((0x80u | (l & 0x3Fu)) << 8)
                : ln == 3 ? (0xE0u | (l >> 12)) | ((0x80u | ((l >> 6) & 0x3Fu)) << 8) | ((0x80u | (l & 0x3Fu)) << 16)
                          : (0xF0u | (l >> 18)) | ((0x80u | ((l >> 12) & 0x3Fu)) << 8) | ((0x80u | ((l >> 6) & 0x3Fu)) << 16) | ((0x80u | (l & 0x3Fu)) << 24);
          }
          t = state;
          for (uint32_t k = 0; k < nraw; k++) { t = ac_step(A, t & ID_MASK, (enc >> (8 * k)) & 0xFFu); }
        }
      }
      state = t & ID_MASK;
      if (!(t & OUT_FLAG)) continue;
      const uint64_t e = p + 1;
      for (uint32_t ch = __ldg(A.first_out + state); ch != NONE; ch = __ldg(A.next_out + ch)) {
        const uint32_t olo = __ldg(A.own_off + ch), ohi = __ldg(A.own_off + ch + 1);
        for (uint32_t o = olo; o < ohi; o++) {
          const uint32_t rank = __ldg(A.own_rank + o);
          const uint32_t id = __ldg(A.id_of_rank + rank);
          if (id <= pass_id) continue;
          uint64_t s;
          if (!ignore_case) s = e - __ldg(A.len_of_rank + rank);
          else { const long long es = skip_back(v, e - 1, (long long)__ldg(&info[id].len_cps) - 1); if (es < 0) { sc->error = 1; continue; } s = (uint64_t)es; }
          // touches edit j: holds a replacement byte, or (deletion) spans the junction
          const bool mine = repl_len ? (s < ns + repl_len && e > ns) : (s < ns && e > ns);
          if (!mine) continue;
          if (has_next && (repl_len ? (s < ns2 + repl_len && e > ns2) : (s < ns2 && e > ns2))) continue;   // the next edit reports it
          append_key(out, reinterpret_cast<unsigned long long*>(&sc->n_new), cap, (e << A.rank_bits) | rank);
        }
      }
    }
  }
}

// Replacer.runWithLimit (:203-242) on a device-resident text, the reference's literal pass structure: every pass scans the
// whole text (or, CaseSensitive, carries the match list) and splices into the other ping-pong buffer.  Used for replacers
// the tiled form below does not take (an empty needle), and with AM_REPLACER_RESCAN=1 as the A/B reference.
// `d_in` is only read; the result is left in a buffer of the library's: *d_out (cudaMalloc'ed, ownership passes to the
// caller), *out_len.
static int replacer_core_classic(const am_replacer* r, const Image* a, int cs, const uint8_t* d_in, uint64_t len, uint64_t max_len, cudaStream_t st,
                         uint8_t** d_out, uint64_t* out_len, int* exceeded) {
  *d_out = nullptr; *out_len = 0; *exceeded = 0; g_last_passes = 0; g_last_rescans = 0;
  Workspace* ws = acquire_ws(a); if (!ws) return fail(AM_E_OOM, "workspace");
  DevBuf text_a, text_b, sel, starts, ends, keep, kidx, kstart, kend, kdelta, kshift, cubtmp, scal;
  int rc;
  auto done = [&](int code) { release_ws(a, ws); return code; };
  if ((rc = scal.ensure(64))) return done(rc);
  const uint8_t* cur = d_in;                       // the current text; text_a / text_b are the library's ping-pong buffers
  DevBuf* nxt = &text_a;
  const uint32_t rank_bits = a->host.rank_bits;
  const uint64_t mask = (1ull << rank_bits) - 1;
  long long prev_id = -1;                          // threshold 1 keeps every priority (:211)
  const long long last_id = (long long)r->n - 1;   // minPriority = 1 - numNeedles (:217)
  struct Scalars { unsigned int best; int error; long long delta_sum; unsigned long long nsel; unsigned long long nkept; };
  Scalars* d_s = scal.as<Scalars>();
  Scalars h_s;

  // carry the match list from pass to pass instead of rescanning (see carry_matches_kernel); AM_REPLACER_RESCAN=1 keeps
  // the reference's literal pass structure (a full scan per pass) for A/B runs
  const char* env_rescan = std::getenv("AM_REPLACER_RESCAN");   // read per call so that tests can A/B both forms
  const bool force_rescan = env_rescan && std::atoi(env_rescan) != 0;
  const bool incremental = cs == AM_CASE_SENSITIVE && !r->has_empty && !r->stored_len_differs && !force_rescan;
  bool have_list = false;
  uint64_t n = 0;
  for (;;) {
    // ---- 1. the matches of the current text: a full scan, or the list carried over from the previous pass ---------
    if (!have_list) {
      am_dev_text t{cur, len, 0, 0};
      if ((rc = find_all_sorted(a, ws, t, st, &n))) return done(rc);
      g_last_rescans++;
    }
    g_last_passes++;
    if (n == 0) break;                             // (_, []) -> Just haystack (:230)
    // ---- 2. the best priority below the threshold ------------------------------------------------------
    h_s = Scalars{0xFFFFFFFFu, 0, 0, 0, 0};
    cudaMemcpyAsync(d_s, &h_s, sizeof h_s, cudaMemcpyHostToDevice, st);
    {
      unsigned blocks = (unsigned)std::min<uint64_t>((n + 255) / 256, 148 * 8);
      g_kernel_launches++;
      best_needle_kernel<<<blocks, 256, 0, st>>>(ws->keys_b, n, mask, a->dev.id_of_rank, prev_id, &d_s->best);
    }
    cudaMemcpyAsync(&h_s, d_s, sizeof h_s, cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return done(cuda_fail(e, "best needle"));
    if (h_s.best == 0xFFFFFFFFu) break;            // nothing below the threshold
    const uint32_t id = h_s.best;
    const uint32_t rank = a->host.rank_of_id[id];
    const uint32_t rl = r->repl_off[id + 1] - r->repl_off[id];
    // ---- 3. this needle's matches, in order ------------------------------------------------------------
    if ((rc = sel.ensure(n * 8))) return done(rc);
    size_t tb = 0;
    RankIs pred{mask, rank};
    cub::DeviceSelect::If(nullptr, tb, ws->keys_b, sel.as<uint64_t>(), &d_s->nsel, (int64_t)n, pred, st);
    if ((rc = cubtmp.ensure(tb))) return done(rc);
    cub::DeviceSelect::If(cubtmp.p, tb, ws->keys_b, sel.as<uint64_t>(), &d_s->nsel, (int64_t)n, pred, st);
    cudaMemcpyAsync(&h_s, d_s, sizeof h_s, cudaMemcpyDeviceToHost, st);
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return done(cuda_fail(e, "select"));
    const uint64_t ns = h_s.nsel;
    if ((rc = starts.ensure(ns * 8)) || (rc = ends.ensure(ns * 8)) || (rc = keep.ensure(ns + 8)) || (rc = kidx.ensure(ns * 8 + 8)) ||
        (rc = kstart.ensure(ns * 8)) || (rc = kend.ensure(ns * 8)) || (rc = kdelta.ensure(ns * 8)) || (rc = kshift.ensure(ns * 8 + 8)))
      return done(rc);
    {
      unsigned blocks = (unsigned)std::min<uint64_t>((ns + 127) / 128, 148 * 8);
      g_kernel_launches++;
      starts_kernel<<<blocks, 128, 0, st>>>(sel.as<uint64_t>(), ns, rank_bits, cur, cs == AM_IGNORE_CASE, r->len_bytes[id], r->len_cps[id],
                                            (long long)rl, starts.as<uint64_t>(), ends.as<uint64_t>(), &d_s->delta_sum, &d_s->error);
    }
    // ---- 4. replacementLength over the un-deoverlapped matches (:240) -----------------------------------
    cudaMemcpyAsync(&h_s, d_s, sizeof h_s, cudaMemcpyDeviceToHost, st);
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return done(cuda_fail(e, "starts"));
    if (h_s.error) return done(fail(AM_E_BADARG, "Invalid use of skipCodePointsBackwards (text is not valid UTF-8?)"));
    if (max_len != UINT64_MAX) {
      const long long would = (long long)len + h_s.delta_sum;
      if (would > 0 && (uint64_t)would > max_len) { *exceeded = 1; return done(AM_OK); }
    }
    // ---- 5. removeOverlap ----------------------------------------------------------------------------------
    cudaMemsetAsync(keep.p, 0, ns + 8, st);
    {
      unsigned blocks = (unsigned)std::min<uint64_t>((ns + 127) / 128, 148 * 8);
      g_kernel_launches++;
      overlap_kernel<<<blocks, 128, 0, st>>>(starts.as<uint64_t>(), ends.as<uint64_t>(), ns, keep.as<uint8_t>());
    }
    tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, keep.as<uint8_t>(), kidx.as<uint64_t>(), (int64_t)ns + 1, st);
    if ((rc = cubtmp.ensure(tb))) return done(rc);
    cub::DeviceScan::ExclusiveSum(cubtmp.p, tb, keep.as<uint8_t>(), kidx.as<uint64_t>(), (int64_t)ns + 1, st);   // kidx[ns] = #kept
    cudaMemcpyAsync(&h_s.nkept, kidx.as<uint64_t>() + ns, 8, cudaMemcpyDeviceToHost, st);
    {
      unsigned blocks = (unsigned)std::min<uint64_t>((ns + 255) / 256, 148 * 8);
      g_kernel_launches++;
      gather_kept_kernel<<<blocks, 256, 0, st>>>(starts.as<uint64_t>(), ends.as<uint64_t>(), keep.as<uint8_t>(), kidx.as<uint64_t>(), ns, (long long)rl,
                                                 kstart.as<uint64_t>(), kend.as<uint64_t>(), kdelta.as<long long>());
    }
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return done(cuda_fail(e, "overlap"));
    const uint64_t K = h_s.nkept;
    // ---- 6. splice -------------------------------------------------------------------------------------------
    tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, kdelta.as<long long>(), kshift.as<long long>(), (int64_t)K, st);
    if ((rc = cubtmp.ensure(tb))) return done(rc);
    cub::DeviceScan::ExclusiveSum(cubtmp.p, tb, kdelta.as<long long>(), kshift.as<long long>(), (int64_t)K, st);
    long long last_shift = 0, last_delta = 0;
    if (K) {
      cudaMemcpyAsync(&last_shift, kshift.as<long long>() + (K - 1), 8, cudaMemcpyDeviceToHost, st);
      cudaMemcpyAsync(&last_delta, kdelta.as<long long>() + (K - 1), 8, cudaMemcpyDeviceToHost, st);
    }
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return done(cuda_fail(e, "scan"));
    const long long total_shift = last_shift + last_delta;
    const uint64_t new_len = (uint64_t)((long long)len + total_shift);
    if ((rc = nxt->ensure(new_len + 64))) return done(rc);
    {
      uint64_t tiles = (len + SPLICE_TILE - 1) / SPLICE_TILE;
      if (tiles == 0) tiles = 1;                                  // empty text with an empty-needle match
      g_kernel_launches++;
      splice_kernel<<<(unsigned)tiles, 256, 0, st>>>(cur, len, nxt->as<uint8_t>(), kstart.as<uint64_t>(), kend.as<uint64_t>(), kshift.as<long long>(), K,
                                                     r->d_repl + r->repl_off[id], rl, total_shift);
      if ((e = cudaGetLastError()) != cudaSuccess) return done(cuda_fail(e, "splice launch"));
    }
    cur = nxt->as<uint8_t>();
    nxt = nxt == &text_a ? &text_b : &text_a;
    len = new_len;
    if ((long long)id == last_id) break;           // p == minPriority: no needle is left (:241)
    prev_id = id;                                   // go p (:242)
    // ---- 7. the next pass's matches without a rescan ----------------------------------------------------------
    have_list = false;
    if (incremental) {
      const uint64_t cap = ws->keys_a_bytes / 8;
      unsigned long long* d_n = reinterpret_cast<unsigned long long*>(ws->d_scalars);
      cudaMemsetAsync(d_n, 0, 8, st);
      {
        unsigned blocks = (unsigned)std::min<uint64_t>((n + 255) / 256, 148 * 8);
        g_kernel_launches++;
        carry_matches_kernel<<<blocks, 256, 0, st>>>(ws->keys_b, n, rank_bits, a->dev.id_of_rank, a->dev.len_of_rank, id, kstart.as<uint64_t>(), kend.as<uint64_t>(),
                                                    kshift.as<long long>(), K, total_shift, ws->keys_a, d_n, cap);
      }
      if (K) {
        unsigned blocks = (unsigned)std::min<uint64_t>((K + 63) / 64, 148 * 16);
        g_kernel_launches++;
        rescan_edits_kernel<<<blocks, 64, 0, st>>>(a->dev, cur, len, kstart.as<uint64_t>(), kshift.as<long long>(), K, rl, id, ws->keys_a, d_n, cap);
      }
      unsigned long long n_new = 0;
      cudaMemcpyAsync(&n_new, d_n, 8, cudaMemcpyDeviceToHost, st);
      if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return done(cuda_fail(e, "carry matches"));
      if (n_new <= cap) {                          // (more than the key buffers hold: fall back to a rescan, which resizes them)
        n = n_new;
        if (n) {
          const int end_bit = std::min(64, bitlen(len) + (int)rank_bits);
          size_t stb = sort_temp_bytes(n, end_bit);
          if ((rc = ws->need_sort_temp(stb))) return done(rc);
          if ((e = sort_keys(ws->sort_temp, stb, ws->keys_a, ws->keys_b, n, end_bit, st)) != cudaSuccess) return done(cuda_fail(e, "radix sort"));
        }
        have_list = true;
      }
    }
  }
  cudaError_t e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return done(cuda_fail(e, "replacer"));
  // hand the result buffer over: the ping-pong buffer that holds it, or a copy of the (untouched) input
  DevBuf* holder = cur == text_a.p ? &text_a : cur == text_b.p ? &text_b : nullptr;
  if (holder) { *d_out = holder->as<uint8_t>(); holder->p = nullptr; holder->cap = 0; }
  else {
    void* p = nullptr;
    if (cudaMalloc(&p, len + 64) != cudaSuccess) { cudaGetLastError(); return done(fail(AM_E_OOM, "cudaMalloc(result)")); }
    if (len && (e = cudaMemcpyAsync(p, cur, len, cudaMemcpyDeviceToDevice, st)) != cudaSuccess) { cudaFree(p); return done(cuda_fail(e, "result copy")); }
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) { cudaFree(p); return done(cuda_fail(e, "result copy")); }
    *d_out = static_cast<uint8_t*>(p);
  }
  *out_len = len;
  return done(AM_OK);
}

// IdInfo of every needle id for one case mode (the rank of a needle depends on the image), built on first use.
static int replacer_idinfo(const am_replacer* cr, const Image* a, int cs, const IdInfo** out) {
  am_replacer* r = const_cast<am_replacer*>(cr);
  std::lock_guard<std::mutex> g(r->mu);
  if (!r->d_idinfo[cs]) {
    std::vector<IdInfo> h(r->n ? r->n : 1);
    for (uint64_t i = 0; i < r->n; i++) h[i] = IdInfo{a->host.rank_of_id[i], r->repl_off[i], r->repl_off[i + 1] - r->repl_off[i], r->len_bytes[i], r->len_cps[i]};
    void* p = nullptr;
    if (cudaMalloc(&p, h.size() * sizeof(IdInfo)) != cudaSuccess) { cudaGetLastError(); return fail(AM_E_OOM, "cudaMalloc(needle info)"); }
    cudaError_t e = cudaMemcpy(p, h.data(), h.size() * sizeof(IdInfo), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(p); return cuda_fail(e, "needle info"); }
    r->d_idinfo[cs] = p;
  }
  *out = static_cast<const IdInfo*>(r->d_idinfo[cs]);
  return AM_OK;
}

// Replacer.runWithLimit (:203-242) with the match list carried between the passes and the text in tiles (see above).
static int replacer_core_tiled(const am_replacer* r, const Image* a, int cs, const uint8_t* d_in, uint64_t len, uint64_t max_len, cudaStream_t st,
                               uint8_t** d_out, uint64_t* out_len, int* exceeded) {
  *d_out = nullptr; *out_len = 0; *exceeded = 0; g_last_passes = 0; g_last_rescans = 0; g_last_replacer_ms = 0.f; g_last_replacer_bytes = 0;
  Workspace* ws = acquire_ws(a); if (!ws) return fail(AM_E_OOM, "workspace");
  ReplacerScratch* box = nullptr;
  {
    am_replacer* mr = const_cast<am_replacer*>(r);
    std::lock_guard<std::mutex> g(mr->mu);
    if (!mr->scratch_pool.empty()) { box = static_cast<ReplacerScratch*>(mr->scratch_pool.back()); mr->scratch_pool.pop_back(); }
  }
  if (!box) box = new ReplacerScratch();
  ReplacerScratch& S = *box;
  DevBuf &contig = S.contig, &tiles = S.tiles, &tile_len = S.tile_len, (&tile_base)[2] = S.tile_base, &tile_flag = S.tile_flag, &touched = S.touched, &new_len = S.new_len,
         &sel = S.sel, &starts = S.starts, &ends = S.ends, &keep = S.keep, &kidx = S.kidx, &kstart = S.kstart, &kend = S.kend, &kdelta = S.kdelta, &kshift = S.kshift,
         &cubtmp = S.cubtmp, &scal = S.scal, &spare = S.spare, &keys_c = S.keys_c, &rkeys = S.rkeys;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  const bool prof = profiling_enabled();
  int rc;
  auto done = [&](int code) {
    if (ev0) { cudaEventDestroy(ev0); cudaEventDestroy(ev1); }
    {
      am_replacer* mr = const_cast<am_replacer*>(r);
      std::lock_guard<std::mutex> g(mr->mu);
      mr->scratch_pool.push_back(box);
    }
    release_ws(a, ws); return code;
  };
  const IdInfo* d_info = nullptr;
  if ((rc = replacer_idinfo(r, a, cs, &d_info)) || (rc = scal.ensure(sizeof(PassScalars) + 64))) return done(rc);
  if (prof) { cudaEventCreate(&ev0); cudaEventCreate(&ev1); cudaEventRecord(ev0, st); }
  PassScalars* d_sc = scal.as<PassScalars>();
  PassScalars* h_sc = reinterpret_cast<PassScalars*>(ws->h_scalars + 256);   // pinned
  const uint32_t rank_bits = a->host.rank_bits;
  const uint64_t mask = (1ull << rank_bits) - 1;
  const int ic = cs == AM_IGNORE_CASE;
  long long prev_id = -1;                          // threshold 1 keeps every priority (:211)
  const long long last_id = (long long)r->n - 1;   // minPriority = 1 - numNeedles (:217)
  uint64_t bytes_moved = 0;

  const uint8_t* cur = d_in;                       // the contiguous text (when !tiled)
  bool tiled = false;
  uint32_t T = 0; int tb_cur = 0;                  // tiles: count, which tile_base buffer is current
  auto view = [&]() -> TextView {
    if (tiled) return TextView{tiles.as<uint8_t>(), len, tile_len.as<uint32_t>(), tile_base[tb_cur].as<uint64_t>(), T, 1u};
    return TextView{cur, len, nullptr, nullptr, 0u, 0u};
  };
  auto scan_tiles = [&](int into) -> int {         // tile_base[into] = exclusive prefix of tile_len (T + 1 entries; tile_len[T] = 0)
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, tile_len.as<uint32_t>(), tile_base[into].as<unsigned long long>(), (int64_t)T + 1, st);
    { int rc2 = cubtmp.ensure(tb); if (rc2) return rc2; }
    cudaError_t e2 = cub::DeviceScan::ExclusiveSum(cubtmp.p, tb, tile_len.as<uint32_t>(), tile_base[into].as<unsigned long long>(), (int64_t)T + 1, st);
    return e2 == cudaSuccess ? AM_OK : cuda_fail(e2, "tile scan");
  };
  auto tileify = [&]() -> int {                    // contiguous `cur` -> tiles
    T = (uint32_t)((len + TILE_FILL - 1) / TILE_FILL); if (T == 0) T = 1;
    int rc2;
    if ((rc2 = tiles.ensure((size_t)T * TILE_CAP + 64)) || (rc2 = tile_len.ensure(((size_t)T + 1) * 4)) || (rc2 = tile_base[0].ensure(((size_t)T + 1) * 8)) ||
        (rc2 = tile_base[1].ensure(((size_t)T + 1) * 8)) || (rc2 = tile_flag.ensure((size_t)T * 4)) || (rc2 = touched.ensure((size_t)T * 4)) || (rc2 = new_len.ensure((size_t)T * 4)))
      return rc2;
    cudaMemsetAsync(tile_flag.p, 0, (size_t)T * 4, st);
    cudaMemsetAsync(tile_len.as<uint32_t>() + T, 0, 4, st);
    g_kernel_launches++;
    tileify_kernel<<<std::min<uint32_t>(T, 148 * 16), 128, 0, st>>>(cur, len, tiles.as<uint8_t>(), tile_len.as<uint32_t>(), T);
    tb_cur = 0;
    bytes_moved += 2 * len;
    return scan_tiles(0);
  };
  auto materialize = [&](DevBuf* into) -> int {    // tiles -> a contiguous buffer of the library's
    int rc2 = into->ensure(len + 64);
    if (rc2) return rc2;
    g_kernel_launches++;
    materialize_kernel<<<std::min<uint32_t>(T, 148 * 16), 128, 0, st>>>(tiles.as<uint8_t>(), tile_len.as<uint32_t>(), tile_base[tb_cur].as<uint64_t>(), T, into->as<uint8_t>());
    bytes_moved += 2 * len;
    return cudaGetLastError() == cudaSuccess ? AM_OK : cuda_fail(cudaGetLastError(), "materialize");
  };

  bool have_list = false;
  uint64_t n = 0;
  // (S.spare: second contiguous buffer of the overflow path; the match list ping-pongs between ws->keys_b and S.keys_c; S.rkeys:
  // the rescanned matches of a pass)
  uint64_t* list = nullptr;                        // the current (sorted) match list
  uint64_t* other = nullptr;
  for (;;) {
    // ---- 1. the matches of the current text: a full scan of the (contiguous) text, or the carried list ------------------------
    if (!have_list) {
      if (tiled) {                                 // (only after a pass whose match list outgrew the key buffers)
        if ((rc = materialize(&contig))) return done(rc);
        cur = contig.as<uint8_t>(); tiled = false;
      }
      am_dev_text t{cur, len, 0, 0};
      if ((rc = find_all_sorted(a, ws, t, st, &n))) return done(rc);
      g_last_rescans++;
      bytes_moved += len;
      if ((rc = keys_c.ensure(ws->keys_b_bytes))) return done(rc);
      list = ws->keys_b; other = keys_c.as<uint64_t>();
    }
    g_last_passes++;
    if (n == 0) break;                             // (_, []) -> Just haystack (:230)
    const uint64_t list_cap = std::min<uint64_t>(ws->keys_b_bytes, keys_c.cap) / 8;
    // ---- 2. the needle of the pass and its matches: first host round trip (two scalars) ---------------------------------------------
    if ((rc = sel.ensure(n * 8))) return done(rc);
    const unsigned gn = (unsigned)std::min<uint64_t>((n + 255) / 256, 148 * 8);
    g_kernel_launches += 4;
    pass_begin_kernel<<<1, 1, 0, st>>>(d_sc);
    best_needle_kernel<<<gn, 256, 0, st>>>(list, n, mask, a->dev.id_of_rank, prev_id, &d_sc->best);
    pass_resolve_kernel<<<1, 1, 0, st>>>(d_sc, d_info);
    {
      size_t tb = 0;
      RankIsDev pred{mask, d_sc};
      cub::DeviceSelect::If(nullptr, tb, list, sel.as<uint64_t>(), &d_sc->nsel, (int64_t)n, pred, st);
      if ((rc = cubtmp.ensure(tb))) return done(rc);
      cub::DeviceSelect::If(cubtmp.p, tb, list, sel.as<uint64_t>(), &d_sc->nsel, (int64_t)n, pred, st);
    }
    cudaMemcpyAsync(h_sc, d_sc, sizeof(PassScalars), cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return done(cuda_fail(e, "replacer pass"));
    if (h_sc->best == NONE) break;                 // nothing below the threshold
    const uint32_t id = h_sc->best;
    const uint64_t ns = h_sc->nsel;
    bytes_moved += n * 8 * 2;
    // ---- 3..7: the rest of the pass is queued; the second round trip reads its scalars --------------------------------------------------
    if ((rc = starts.ensure(ns * 8)) || (rc = ends.ensure(ns * 8)) || (rc = keep.ensure(ns + 16)) || (rc = kidx.ensure((ns + 1) * 8)) ||
        (rc = kstart.ensure(ns * 8 + 8)) || (rc = kend.ensure(ns * 8 + 8)) || (rc = kdelta.ensure((ns + 1) * 8)) || (rc = kshift.ensure((ns + 1) * 8)))
      return done(rc);
    const uint64_t rcap = std::max<uint64_t>(1 << 16, 4 * ns * (a->host.halo_bytes + 2));   // rescanned matches: a few per edit at most (overflow: the next pass scans)
    if ((rc = rkeys.ensure(std::min<uint64_t>(rcap, list_cap) * 8))) return done(rc);
    const uint64_t rkeys_cap = rkeys.cap / 8;
    const unsigned gs = (unsigned)std::min<uint64_t>((ns + 127) / 128, 148 * 8);
    {
      size_t tb2 = 0, tb3 = 0, tb4 = 0;
      cub::DeviceScan::ExclusiveSum(nullptr, tb2, keep.as<uint8_t>(), kidx.as<uint64_t>(), (int64_t)ns + 1, st);
      cub::DeviceScan::ExclusiveSum(nullptr, tb3, kdelta.as<long long>(), kshift.as<long long>(), (int64_t)ns + 1, st);
      cub::DeviceSelect::If(nullptr, tb4, other, other, &d_sc->n_carried, (int64_t)n, NotDropped(), st);
      if ((rc = cubtmp.ensure(std::max(std::max(tb2, tb3), tb4)))) return done(rc);
      TextView v = view();
      g_kernel_launches += 6;
      starts_view_kernel<<<gs, 128, 0, st>>>(sel.as<uint64_t>(), d_sc, rank_bits, v, ic, starts.as<uint64_t>(), ends.as<uint64_t>(), d_sc);
      cudaMemsetAsync(keep.p, 0, ns + 16, st);
      cudaMemsetAsync(kdelta.p, 0, (ns + 1) * 8, st);
      overlap_dev_kernel<<<gs, 256, 0, st>>>(starts.as<uint64_t>(), ends.as<uint64_t>(), d_sc, keep.as<uint8_t>());
      cub::DeviceScan::ExclusiveSum(cubtmp.p, tb2, keep.as<uint8_t>(), kidx.as<uint64_t>(), (int64_t)ns + 1, st);               // kidx[ns] = #kept
      gather_kept_dev_kernel<<<gs, 256, 0, st>>>(starts.as<uint64_t>(), ends.as<uint64_t>(), keep.as<uint8_t>(), kidx.as<uint64_t>(), ns, d_sc,
                                                 kstart.as<uint64_t>(), kend.as<uint64_t>(), kdelta.as<long long>());
      cub::DeviceScan::ExclusiveSum(cubtmp.p, tb3, kdelta.as<long long>(), kshift.as<long long>(), (int64_t)ns + 1, st);          // kshift[ns] = total shift
      total_shift_kernel<<<1, 1, 0, st>>>(kshift.as<long long>(), ns, d_sc);
      // the next pass's list, part (a): on the text as it still is.  Transformed in place order (into sel's space: n keys),
      // then compacted into the other list buffer -- still sorted.
      carry_view_kernel<<<gn, 256, 0, st>>>(list, n, rank_bits, a->dev.id_of_rank, a->dev.len_of_rank, d_info, ic, v, kstart.as<uint64_t>(), kend.as<uint64_t>(),
                                            kshift.as<long long>(), ns, d_sc, sel.as<uint64_t>());
      cub::DeviceSelect::If(cubtmp.p, tb4, sel.as<uint64_t>(), other, &d_sc->n_carried, (int64_t)n, NotDropped(), st);
    }
    bytes_moved += n * 8 * 4;
    // the text: rewrite the touched tiles in place
    if (!tiled) { if ((rc = tileify())) return done(rc); tiled = true; }
    {
      TextView v = view();
      g_kernel_launches += 4;
      tile_mark_kernel<<<gs, 128, 0, st>>>(d_sc, kstart.as<uint64_t>(), kend.as<uint64_t>(), v, tile_flag.as<uint32_t>(), touched.as<uint32_t>(), d_sc);
      tile_plan_kernel<<<gs, 128, 0, st>>>(d_sc, touched.as<uint32_t>(), kstart.as<uint64_t>(), kend.as<uint64_t>(), v, new_len.as<uint32_t>());
      const unsigned gt = (unsigned)std::min<uint64_t>(std::max<uint64_t>(1, std::min<uint64_t>(2 * ns + 1, T)), 148 * 16);
      tile_splice_kernel<<<gt, 128, 0, st>>>(d_sc, touched.as<uint32_t>(), new_len.as<uint32_t>(), kstart.as<uint64_t>(), kend.as<uint64_t>(), tiles.as<uint8_t>(),
                                             tile_len.as<uint32_t>(), tile_base[tb_cur].as<uint64_t>(), tile_flag.as<uint32_t>(), r->d_repl);
      if ((rc = scan_tiles(tb_cur ^ 1))) return done(rc);
      // part (b): the neighbourhood of every edit, on the rewritten tiles (the host does not know the new length yet: the kernel
      // takes it from the new prefix sums)
      TextView vn{tiles.as<uint8_t>(), 0, tile_len.as<uint32_t>(), tile_base[tb_cur ^ 1].as<uint64_t>(), T, 1u};
      const unsigned gk = (unsigned)std::min<uint64_t>((ns + 63) / 64, 148 * 16);
      rescan_view_kernel<<<gk, 64, 0, st>>>(a->dev, vn, d_info, ic, kstart.as<uint64_t>(), kshift.as<long long>(), d_sc, rkeys.as<uint64_t>(), rkeys_cap);
    }
    // ---- the second host round trip of the pass -------------------------------------------------------------------------------------------
    cudaMemcpyAsync(h_sc, d_sc, sizeof(PassScalars), cudaMemcpyDeviceToHost, st);
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return done(cuda_fail(e, "replacer pass"));
    if (h_sc->error) return done(fail(AM_E_BADARG, "Invalid use of skipCodePointsBackwards (text is not valid UTF-8?)"));
    if (max_len != UINT64_MAX) {                   // replacementLength over the un-deoverlapped matches (:240)
      const long long would = (long long)len + h_sc->delta_sum;
      if (would > 0 && (uint64_t)would > max_len) { *exceeded = 1; return done(AM_OK); }
    }
    const uint64_t K = h_sc->nkept;
    const uint64_t new_len_total = (uint64_t)((long long)len + h_sc->total_shift);
    bytes_moved += (uint64_t)h_sc->ntouched * 2 * TILE_FILL + K * (2 * a->host.halo_bytes + h_sc->id.repl_len);
    have_list = false;
    if (h_sc->overflow) {
      // a tile would have outgrown its slot: this pass splices the contiguous way, the next one cuts new tiles
      if ((rc = materialize(&contig))) return done(rc);
      if ((rc = spare.ensure(new_len_total + 64))) return done(rc);
      uint64_t tl = (len + SPLICE_TILE - 1) / SPLICE_TILE; if (tl == 0) tl = 1;
      g_kernel_launches++;
      splice_kernel<<<(unsigned)tl, 256, 0, st>>>(contig.as<uint8_t>(), len, spare.as<uint8_t>(), kstart.as<uint64_t>(), kend.as<uint64_t>(), kshift.as<long long>(), K,
                                                  r->d_repl + h_sc->id.repl_off, h_sc->id.repl_len, h_sc->total_shift);
      if ((e = cudaGetLastError()) != cudaSuccess) return done(cuda_fail(e, "splice launch"));
      std::swap(contig.p, spare.p); std::swap(contig.cap, spare.cap);
      cur = contig.as<uint8_t>(); tiled = false;
      bytes_moved += len + new_len_total;
      len = new_len_total;                         // (the carried list lacks part (b): the next pass scans)
    } else {
      tb_cur ^= 1;
      len = new_len_total;
      const uint64_t nA = h_sc->n_carried, nB = h_sc->n_new;
      if (nB <= rkeys_cap && nA + nB <= list_cap) {
        if (nB == 0) std::swap(list, other);       // the compacted carried list IS the next pass's list
        else {
          // sort the few rescanned matches, merge them into the carried list (into the buffer the old list occupied)
          const int end_bit = std::min(64, bitlen(len) + (int)rank_bits);
          size_t stb = sort_temp_bytes(nB, end_bit);
          if ((rc = ws->need_sort_temp(stb)) || (rc = sel.ensure(nB * 8))) return done(rc);
          if ((e = sort_keys(ws->sort_temp, stb, rkeys.as<uint64_t>(), sel.as<uint64_t>(), nB, end_bit, st)) != cudaSuccess) return done(cuda_fail(e, "radix sort"));
          g_kernel_launches++;
          merge_kernel<<<(unsigned)std::min<uint64_t>((nA + nB + 255) / 256, 148 * 8), 256, 0, st>>>(other, nA, sel.as<uint64_t>(), nB, list);
          bytes_moved += (nA + nB) * 8 * 2;
        }
        n = nA + nB;
        have_list = true;
      }                                            // (else: more matches than the buffers hold -- the next pass scans, which resizes them)
    }
    if ((long long)id == last_id) break;           // p == minPriority: no needle is left (:241)
    prev_id = id;                                   // go p (:242)
  }
  // ---- the result: a contiguous buffer of the library's ----------------------------------------------------------------------------------
  DevBuf result;
  if (tiled) { if ((rc = materialize(&result))) return done(rc); }
  else if (cur == contig.p && contig.p) { std::swap(result.p, contig.p); std::swap(result.cap, contig.cap); }
  else {
    if ((rc = result.ensure(len + 64))) return done(rc);
    if (len && cudaMemcpyAsync(result.p, cur, len, cudaMemcpyDeviceToDevice, st) != cudaSuccess) return done(cuda_fail(cudaGetLastError(), "result copy"));
    bytes_moved += 2 * len;
  }
  if (prof) cudaEventRecord(ev1, st);
  cudaError_t e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return done(cuda_fail(e, "replacer"));
  if (prof) cudaEventElapsedTime(&g_last_replacer_ms, ev0, ev1);
  g_last_replacer_bytes = bytes_moved;
  *d_out = result.as<uint8_t>(); result.p = nullptr; result.cap = 0;
  *out_len = len;
  return done(AM_OK);
}

// The form a run takes: tiles + carried match list unless the replacer holds an empty needle (its matches have no span to
// carry), its stored needles and payload lengths disagree in a CaseSensitive run (a replacer switched to the other case
// mode after `build`), a needle is longer than a quarter tile, or AM_REPLACER_RESCAN=1 asks for the reference's literal form.
static int replacer_core(const am_replacer* r, const Image* a, int cs, const uint8_t* d_in, uint64_t len, uint64_t max_len, cudaStream_t st,
                         uint8_t** d_out, uint64_t* out_len, int* exceeded) {
  const char* env_rescan = std::getenv("AM_REPLACER_RESCAN");   // read per call so that tests can A/B both forms
  const bool force_rescan = env_rescan && std::atoi(env_rescan) != 0;
  const bool ok = !r->has_empty && !(cs == AM_CASE_SENSITIVE && r->stored_len_differs) && a->host.max_len <= TILE_FILL / 4 && a->host.halo_bytes <= TILE_FILL / 4;
  if (ok && !force_rescan) return replacer_core_tiled(r, a, cs, d_in, len, max_len, st, d_out, out_len, exceeded);
  return replacer_core_classic(r, a, cs, d_in, len, max_len, st, d_out, out_len, exceeded);
}

extern "C" {

void am_replacer_free(am_replacer* r);

// The stored form of a Replacer: the needles exactly as its searcher holds them (lowered iff it was built with
// IgnoreCase, :105-107) and the Payload lengths of the ORIGINAL needles (:111-113).  `build` derives it; `compose`
// (:120-133), `mapReplacement` (:136-141) and the derived FromJSON instance start from it and lower nothing.
static int replacer_from_stored(const am_u8slice* stored, const uint32_t* len_bytes, const uint32_t* len_cps, const am_u8slice* repls, size_t n,
                                int prepare_cs, const am_lower_table* lower, const am_options* opts, am_replacer** out) {
  am_replacer* r = new am_replacer();
  r->built_cs = prepare_cs; r->n = n;
  r->repl_off.assign(n + 1, 0);
  for (size_t i = 0; i < n; i++) {
    if (stored[i].len < 0 || repls[i].len < 0 || stored[i].off < 0 || repls[i].off < 0 || (stored[i].len && !stored[i].ptr) || (repls[i].len && !repls[i].ptr)) { delete r; return fail(AM_E_BADARG, "bad slice"); }
    r->len_bytes.push_back(len_bytes[i]);
    r->len_cps.push_back(len_cps[i]);
    if (stored[i].len == 0) r->has_empty = true;
    if ((uint64_t)stored[i].len != len_bytes[i]) r->stored_len_differs = true;
    r->repl_bytes.insert(r->repl_bytes.end(), repls[i].ptr + repls[i].off, repls[i].ptr + repls[i].off + repls[i].len);
    r->repl_off[i + 1] = (uint32_t)r->repl_bytes.size();
  }
  int rc = am_automaton_build(stored, n, lower, opts, &r->automaton);
  if (!rc && !(prepare_cs == AM_IGNORE_CASE && r->has_empty)) rc = am_automaton_prepare(r->automaton, prepare_cs);
  if (rc) { am_replacer_free(r); return rc; }
  if (r->automaton->device >= 0) {
    DeviceGuard g;
    g.enter(r->automaton->device);
    if (cudaMalloc((void**)&r->d_repl, r->repl_bytes.size() + 16) != cudaSuccess) { cudaGetLastError(); am_replacer_free(r); return fail(AM_E_OOM, "cudaMalloc(replacements)"); }
    if (!r->repl_bytes.empty()) cudaMemcpy(r->d_repl, r->repl_bytes.data(), r->repl_bytes.size(), cudaMemcpyHostToDevice);
  }
  *out = r;
  return AM_OK;
}

int am_replacer_build_stored(const am_u8slice* stored, const uint32_t* len_bytes, const uint32_t* len_cps, const am_u8slice* repls, size_t n,
                             int cs, const am_lower_table* lower, const am_options* opts, am_replacer** out) {
  if (!out) return fail(AM_E_BADARG, "out is null");
  *out = nullptr;
  if (n > 0 && (!stored || !repls || !len_bytes || !len_cps)) return fail(AM_E_BADARG, "null argument");
  if (cs != AM_CASE_SENSITIVE && cs != AM_IGNORE_CASE) return fail(AM_E_BADARG, "unknown case sensitivity");
  if (cs == AM_IGNORE_CASE && !lower) return fail(AM_E_BADARG, "IgnoreCase needs the Char.toLower table");
  return replacer_from_stored(stored, len_bytes, len_cps, repls, n, cs, lower, opts, out);
}

int am_replacer_build(const am_u8slice* needles, const am_u8slice* repls, size_t n, int cs, const am_lower_table* lower,
                      const am_options* opts, am_replacer** out) {
  if (!out) return fail(AM_E_BADARG, "out is null");
  *out = nullptr;
  if (n > 0 && (!needles || !repls)) return fail(AM_E_BADARG, "needles / replacements is null");
  if (cs != AM_CASE_SENSITIVE && cs != AM_IGNORE_CASE) return fail(AM_E_BADARG, "unknown case sensitivity");
  if (cs == AM_IGNORE_CASE && !lower) return fail(AM_E_BADARG, "IgnoreCase needs the Char.toLower table");
  LowerTable lt;
  int rc = build_lower_table(cs == AM_IGNORE_CASE ? lower : nullptr, &lt);
  if (rc) return fail(rc, "bad lower table");
  std::vector<std::vector<uint8_t>> built(n);
  std::vector<am_u8slice> slices(n);
  std::vector<uint32_t> lb(n), lc(n);
  for (size_t i = 0; i < n; i++) {
    if (needles[i].len < 0 || needles[i].off < 0 || (needles[i].len && !needles[i].ptr)) return fail(AM_E_BADARG, "bad slice");
    const uint8_t* d = needles[i].ptr + needles[i].off;
    uint32_t cps = 0;
    for (int64_t k = 0; k < needles[i].len; k++) cps += (d[k] & 0xC0) != 0x80;
    lb[i] = (uint32_t)needles[i].len;                             // needleLengthBytes of the ORIGINAL needle (:112)
    lc[i] = cps;                                                  // needleLengthCodePoints (:113)
    if (cs == AM_IGNORE_CASE) lower_utf8_host(lt, d, needles[i].len, &built[i]);   // Utf8.lowerUtf8 needle (:107)
    else built[i].assign(d, d + needles[i].len);
    slices[i] = am_u8slice{built[i].data(), 0, (int64_t)built[i].size()};
  }
  if (cs == AM_IGNORE_CASE)
    for (size_t i = 0; i < n; i++)
      if (needles[i].len == 0)
        return fail(AM_E_UNSUPPORTED, "empty needle in an IgnoreCase replacer: the reference's skipCodePointsBackwards diverges on it");
  return replacer_from_stored(slices.data(), lb.data(), lc.data(), repls, n, cs, lower, opts, out);
}

void am_replacer_free(am_replacer* r) {
  if (!r) return;
  if (r->d_repl) {
    DeviceGuard g;
    if (r->automaton && r->automaton->device >= 0) g.enter(r->automaton->device);
    cudaFree(r->d_repl);
    for (void* p : r->d_idinfo) if (p) cudaFree(p);
    for (void* b : r->scratch_pool) delete static_cast<ReplacerScratch*>(b);
  }
  if (r->automaton) am_automaton_free(r->automaton);
  delete r;
}

static int replacer_enter(const am_replacer* r, int cs, Image** im, DeviceGuard* g) {
  if (!r) return fail(AM_E_BADARG, "replacer is null");
  int rc = get_image(r->automaton, cs, im);
  if (rc) return rc;
  if (cs == AM_IGNORE_CASE && r->has_empty)
    return fail(AM_E_UNSUPPORTED, "empty needle in an IgnoreCase replacer: the reference's skipCodePointsBackwards diverges on it");
  return check_ready(*im, g);
}

int am_replacer_run_dev(const am_replacer* r, int cs, const void* dev_text, uint64_t text_len, uint64_t max_len, void* stream, void** dev_out,
                        uint64_t* out_len, int* exceeded) {
  AM_NVTX("am_replacer_run_dev");
  if (!dev_out || !out_len || !exceeded) return fail(AM_E_BADARG, "null argument");
  if (text_len > 0 && !dev_text) return fail(AM_E_BADARG, "dev_text is null");
  Image* im = nullptr; DeviceGuard guard;
  int rc = replacer_enter(r, cs, &im, &guard); if (rc) return rc;
  uint8_t* d = nullptr;
  rc = replacer_core(r, im, cs, static_cast<const uint8_t*>(dev_text), text_len, max_len, static_cast<cudaStream_t>(stream), &d, out_len, exceeded);
  *dev_out = d;
  return rc;
}

int am_replacer_run(const am_replacer* r, int cs, const am_u8slice* hay, uint64_t max_len, uint8_t** out, uint64_t* out_len, int* exceeded) {
  AM_NVTX("am_replacer_run");
  if (!out || !out_len || !exceeded || !hay) return fail(AM_E_BADARG, "null argument");
  if (hay->len < 0 || hay->off < 0 || (hay->len > 0 && !hay->ptr)) return fail(AM_E_BADARG, "bad text slice");
  Image* im = nullptr; DeviceGuard guard;
  int rc = replacer_enter(r, cs, &im, &guard); if (rc) return rc;
  *out = nullptr; *out_len = 0; *exceeded = 0;
  const uint64_t len = (uint64_t)hay->len;
  DevBuf in;
  if ((rc = in.ensure(len + 64))) return rc;
  cudaStream_t st = 0;
  if (len) {
    cudaError_t e = cudaMemcpyAsync(in.p, hay->ptr + hay->off, len, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return cuda_fail(e, "H2D text");
  }
  uint8_t* d = nullptr; uint64_t n = 0;
  rc = replacer_core(r, im, cs, in.as<uint8_t>(), len, max_len, st, &d, &n, exceeded);
  if (rc || *exceeded) { if (d) cudaFree(d); return rc; }
  uint8_t* host = static_cast<uint8_t*>(std::malloc(n ? n : 1));
  if (!host) { cudaFree(d); return fail(AM_E_OOM, "malloc(result)"); }
  if (n) {
    cudaError_t e = cudaMemcpy(host, d, n, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { std::free(host); cudaFree(d); return cuda_fail(e, "D2H result"); }
  }
  cudaFree(d);
  *out = host; *out_len = n;
  return AM_OK;
}

}  // extern "C"
