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

// Replacer.runWithLimit (:203-242) on a device-resident text.  `d_in` is only read; the result is left in a buffer of the
// library's: *d_out (cudaMalloc'ed, ownership passes to the caller), *out_len.
static int replacer_core(const am_replacer* r, const Image* a, int cs, const uint8_t* d_in, uint64_t len, uint64_t max_len, cudaStream_t st,
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
