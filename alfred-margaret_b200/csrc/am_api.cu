// am_api.cu -- the C ABI of libam_b200 (include/am_b200.h).
//
// Host orchestration only: automaton upload, workspaces, launch order, result readback.
// There is deliberately NO CPU fallback: without an sm_100 device every compute entry point
// fails with AM_E_NODEVICE.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "am_api_internal.h"

namespace am {

thread_local std::string g_last_error;
thread_local uint64_t g_last_passes = 0, g_last_rescans = 0;
thread_local float g_last_replacer_ms = 0.f;
thread_local uint64_t g_last_replacer_bytes = 0;
std::atomic<uint64_t> g_kernel_launches{0};
static std::atomic<int> g_profile{0};
// profiling: the event pair (owned by a pooled workspace) around the scan kernels of this thread's most recent scan
thread_local cudaEvent_t g_ev0 = nullptr, g_ev1 = nullptr;
thread_local int g_ev_device = -1;
bool profiling_enabled() { return g_profile.load(std::memory_order_relaxed) != 0; }

int fail(int code, const std::string& msg) { g_last_error = msg; return code; }
int cuda_fail(cudaError_t e, const char* what) {
  g_last_error = std::string(what) + ": " + cudaGetErrorString(e);
  return AM_E_CUDA;
}

static int usable_device_count() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  int ok = 0;
  for (int d = 0; d < n; d++) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) ok++;
  }
  return ok;
}

// ---- device memory helpers ---------------------------------------------------------------------------
template <class T>
static int upload(Image* a, const std::vector<T>& v, const T** out) {
  void* p = nullptr;
  size_t bytes = std::max<size_t>(v.size() * sizeof(T), 16);
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(automaton)");
  a->dev_allocs.push_back(p);
  if (!v.empty()) {
    e = cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpy(automaton)");
  }
  *out = static_cast<const T*>(p);
  return AM_OK;
}

Workspace::~Workspace() {
  for (void* p : {(void*)d_scalars, (void*)keys_a, (void*)keys_b, sort_temp, (void*)text, (void*)matches, (void*)aux_a, (void*)aux_b, (void*)seg_counts, (void*)seg_bases, (void*)seen_bits, (void*)surv, (void*)surv_counts})
    if (p) cudaFree(p);
  if (ev0) cudaEventDestroy(ev0);
  if (ev1) cudaEventDestroy(ev1);
  if (h_scalars) cudaFreeHost(h_scalars);
  if (copy_stream) cudaStreamDestroy(copy_stream);
  if (scan_stream) cudaStreamDestroy(scan_stream);
  for (cudaEvent_t ev : copy_done) if (ev) cudaEventDestroy(ev);
}

static int ws_grow(void** p, size_t* cap, size_t need, const char* what) {
  if (*cap >= need) return AM_OK;
  if (*p) cudaFree(*p);
  *p = nullptr; *cap = 0;
  size_t want = need + need / 8 + 256;
  cudaError_t e = cudaMalloc(p, want);
  if (e != cudaSuccess) { cudaGetLastError(); g_last_error = std::string("cudaMalloc(") + what + ") of " + std::to_string(want) + " bytes failed"; return AM_E_OOM; }
  *cap = want;
  return AM_OK;
}
int Workspace::need_keys(uint64_t n) {
  int rc = ws_grow((void**)&keys_a, &keys_a_bytes, n * 8, "keys");
  if (rc) return rc;
  return ws_grow((void**)&keys_b, &keys_b_bytes, n * 8, "keys");
}
int Workspace::need_sort_temp(size_t b) { return ws_grow(&sort_temp, &sort_temp_bytes_, b, "sort temp"); }
int Workspace::need_text(uint64_t n) { return ws_grow((void**)&text, &text_bytes, n + 64, "text"); }
int Workspace::need_matches(uint64_t n) { return ws_grow((void**)&matches, &matches_bytes, n * sizeof(am_match), "matches"); }
int Workspace::need_aux(uint64_t a_bytes, uint64_t b_bytes) {
  int rc = ws_grow((void**)&aux_a, &aux_a_bytes, a_bytes, "aux");
  if (rc) return rc;
  return ws_grow((void**)&aux_b, &aux_b_bytes, b_bytes, "aux");
}

int Workspace::need_segs(uint64_t n) {
  int rc = ws_grow((void**)&seg_counts, &seg_counts_bytes, n * 4 + 16, "segment counters");
  if (rc) return rc;
  return ws_grow((void**)&seg_bases, &seg_bases_bytes, n * 8 + 16, "segment bases");
}

int Workspace::need_surv(uint64_t entries) { return ws_grow((void**)&surv, &surv_bytes, entries * 16, "survivor list"); }
int Workspace::need_seen(uint64_t bits) { return ws_grow((void**)&seen_bits, &seen_bytes, (bits + 31) / 32 * 4 + 16, "needle bit set"); }

Image::~Image() {
  DeviceGuard g;
  if (device >= 0) g.enter(device);
  for (Workspace* w : ws_pool) delete w;
  for (void* p : dev_allocs) cudaFree(p);
}

constexpr size_t SCALARS_BYTES = 1024;
constexpr uint32_t SURV_REGIONS_MAX = 1024;   // per-CTA regions of the survivor list (one per SM in use)   // d_scalars / h_scalars: [0..64) scan counters, [64..) the gathered per-rank counts of the sharded calls

Workspace* acquire_ws(const Image* ca) {
  Image* a = const_cast<Image*>(ca);
  {
    std::lock_guard<std::mutex> g(a->ws_mutex);
    if (!a->ws_pool.empty()) { Workspace* w = a->ws_pool.back(); a->ws_pool.pop_back(); return w; }
  }
  Workspace* w = new Workspace();
  if (cudaMalloc((void**)&w->d_scalars, SCALARS_BYTES) != cudaSuccess || cudaMallocHost((void**)&w->h_scalars, SCALARS_BYTES) != cudaSuccess ||
      cudaMalloc((void**)&w->surv_counts, SURV_REGIONS_MAX * 8) != cudaSuccess) {
    cudaGetLastError(); delete w; return nullptr;
  }
  return w;
}
void release_ws(const Image* ca, Workspace* w) {
  Image* a = const_cast<Image*>(ca);
  std::lock_guard<std::mutex> g(a->ws_mutex);
  a->ws_pool.push_back(w);
}

int check_ready(const Image* a, DeviceGuard* g) {
  if (!a) return fail(AM_E_BADARG, "automaton is null");
  if (a->device < 0) return fail(AM_E_NODEVICE, "automaton was built without a device (host image only); there is no CPU fallback");
  cudaError_t e = g->enter(a->device);
  if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
  return AM_OK;
}


// Launch one scan in `mode`; COUNT/EMIT totals land in ws->d_scalars[0], the ANY flag in d_scalars[8..].
int launch_scan(const Image* a, Workspace* ws, const am_dev_text& t, int mode, cudaStream_t st) {
  if (t.report_begin > t.text_len) return fail(AM_E_BADARG, "report_begin > text_len");
  if (t.text_len > 0 && !t.dev_text) return fail(AM_E_BADARG, "dev_text is null");
  if (bitlen(t.text_len + t.pos_base) + (int)a->host.rank_bits > 64) return fail(AM_E_UNSUPPORTED, "position and needle rank do not fit a 64-bit sort key");
  ScanArgs sa;
  sa.text = static_cast<const uint8_t*>(t.dev_text);
  sa.text_len = t.text_len; sa.report_begin = t.report_begin; sa.pos_base = t.pos_base;
  sa.d_count = reinterpret_cast<unsigned long long*>(ws->d_scalars);
  sa.d_flag = reinterpret_cast<int*>(ws->d_scalars + 8);
  sa.d_keys = ws->keys_a; sa.cap = ws->keys_a_bytes / 8;
  static const uint32_t dbg = []() { const char* e = std::getenv("AM_DEBUG_FLAGS"); return e ? (uint32_t)std::atoi(e) : 0u; }();
  static const uint32_t force_list = []() { const char* e = std::getenv("AM_FILTER_LIST"); return e ? (uint32_t)std::atoi(e) : 0u; }();
  sa.force_list = force_list;
  sa.debug = dbg; sa.krow = 4u * (uint32_t)filter_copies(a->dev.q, a->dev.t2_exact != 0);
  cudaError_t e = cudaMemsetAsync(ws->d_scalars, 0, 16, st);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync");
  // filter scan: the survivor list between filter_kernel and verify_kernel.  1 / 256 of the positions to begin with (C2: 1 / 2000
  // survive); a scan that produces more is repeated with the list grown to what it asked for (scan_overflowed).
  sa.surv = nullptr; sa.surv_counts = ws->surv_counts; sa.surv_count = reinterpret_cast<unsigned long long*>(ws->d_scalars + 48); sa.surv_cap_cta = 0; sa.any_mode = mode == MODE_ANY;
  sa.surv_regions = std::min<uint32_t>((uint32_t)sm_count(), SURV_REGIONS_MAX);
  ws->last_span = t.text_len - std::min(t.report_begin, t.text_len);
  Image* am = const_cast<Image*>(a);
  const unsigned scan_no = am->scans.fetch_add(1, std::memory_order_relaxed);
  const bool use_filter = a->kernel_kind == 2 && !ws->force_walk && !(am->walk_streak.load(std::memory_order_relaxed) >= 2 && (scan_no & 7u) != 0);
  ws->force_walk = false;
  if (use_filter) {
    const uint64_t span = ws->last_span;
    int rc = ws->need_surv(std::max<uint64_t>(1u << 16, span / 256));
    if (rc) return rc;
    sa.surv = ws->surv; sa.surv_cap_cta = ws->surv_bytes / 16 / sa.surv_regions;
    if ((e = cudaMemsetAsync(ws->surv_counts, 0, (size_t)sa.surv_regions * 8, st)) != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync");
  }
  if ((e = cudaMemsetAsync(ws->d_scalars + 48, 0, 16, st)) != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync");   // the survivor counts
  // EMIT on the filter kernel: keys go into per-segment slots of keys_a (first half) + an overflow area (second half)
  ws->emit_segmented = false;
  sa.seg_counts = nullptr; sa.seg_shift = SEG_SHIFT; sa.seg_cap = 0; sa.ovf_base = 0; sa.ovf_cap = 0;
  sa.ic_one_pass = 0;
  if (mode == MODE_EMIT && use_filter && t.text_len > t.report_begin) {
    const uint64_t num_segs = ((t.text_len - t.report_begin) + (1ull << SEG_SHIFT) - 1) >> SEG_SHIFT;
    uint64_t per = (sa.cap / 2) / num_segs;
    uint32_t seg_cap = 0;
    if (per >= 16) { seg_cap = 16; while (seg_cap * 2 <= per && seg_cap < SEG_CAP_MAX) seg_cap *= 2; }
    int rc = ws->need_segs(num_segs);
    if (rc) return rc;
    e = cudaMemsetAsync(ws->seg_counts, 0, (num_segs + 1) * 4, st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync");
    ws->num_segs = num_segs; ws->seg_cap = seg_cap; ws->ovf_base = num_segs * seg_cap; ws->ovf_cap = sa.cap - ws->ovf_base;
    sa.seg_counts = ws->seg_counts; sa.seg_cap = seg_cap; sa.ovf_base = ws->ovf_base; sa.ovf_cap = ws->ovf_cap;
    ws->emit_segmented = true;   // (cleared again below if the scan has to take the walk kernel)
  }
  const bool prof = profiling_enabled();
  if (prof) {
    if (!ws->ev0) { cudaEventCreate(&ws->ev0); cudaEventCreate(&ws->ev1); }
    g_ev0 = ws->ev0; g_ev1 = ws->ev1; g_ev_device = a->device;
    cudaEventRecord(g_ev0, st);
  }
  ws->last_kernel = use_filter ? 2 : 1;
  ws->last_inline = use_filter && !a->dev.ignore_case && a->dev.q >= 1 && a->dev.q <= 4 && !sa.force_list;   // (the predicate of launch_filter)
  if (use_filter && a->host.case_sensitivity == AM_IGNORE_CASE && t.text_len > 0) {
    // runLower on the filter kernel.  ONE pass over the original text: the probe and the second level work on folded
    // bytes (every byte | 0x20; the cells hold every case variant of the needles' first code points: no `Char.toLower`
    // anywhere near the hot loop), the few survivors are verified on code points lowered
    // on the fly.  Code points whose lower case has another UTF-8 length are matched by the needle variants the
    // automaton holds for them (am_build.cpp step 1).  Only an automaton that could not take the variants
    // (ic_copy_exact == false) scans a lowered COPY of the text in which such code points are marked, and falls back to
    // the exact per-code-point walk when the text holds one.
    static const bool one_pass_off = []() { const char* v = std::getenv("AM_IC_ONE_PASS"); return v && std::atoi(v) == 0; }();
    const bool keep = a->host.ic_copy_exact;
    if (keep && a->host.ic_fold_ok && !one_pass_off) {
      sa.ic_one_pass = 1;
      e = launch_filter(a->dev, sa, mode, st);
    } else {
      const uint32_t a0 = (uint32_t)(reinterpret_cast<uintptr_t>(t.dev_text) & 15);
      int rc = ws->need_aux(t.text_len + 64, 0);
      if (rc) return rc;
      unsigned int* d_exc = reinterpret_cast<unsigned int*>(ws->d_scalars + 16);
      unsigned int exceptions = 0;
      e = keep ? cudaSuccess : cudaMemsetAsync(d_exc, 0, 4, st);
      if (e == cudaSuccess) e = launch_lower(a->dev, sa.text, t.text_len, ws->aux_a + a0, d_exc, keep, st);
      if (!keep) {
        if (e == cudaSuccess) e = cudaMemcpyAsync(ws->h_scalars + 16, d_exc, 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e == cudaSuccess) exceptions = *reinterpret_cast<unsigned int*>(ws->h_scalars + 16);
      }
      if (e != cudaSuccess) return cuda_fail(e, "lowering pass");
      if (exceptions == 0) {
        sa.text = ws->aux_a + a0;
        e = launch_filter(a->dev, sa, mode, st);
      } else {
        ws->emit_segmented = false; ws->last_kernel = 1;
        e = launch_walk(a->dev, sa, mode, st);
      }
    }
  } else {
    e = use_filter ? launch_filter(a->dev, sa, mode, st) : launch_walk(a->dev, sa, mode, st);
  }
  if (prof) cudaEventRecord(g_ev1, st);
  if (e != cudaSuccess) return cuda_fail(e, "kernel launch");
  return AM_OK;
}

// Scan in COUNT / ANY mode and read the counters back; a scan whose survivor list was too short is repeated.
int scan_sync(const Image* a, Workspace* ws, const am_dev_text& t, int mode, cudaStream_t st);

// After read_scalars: did the last filter scan produce more survivors than its list holds?  Then its results are incomplete:
// grow the list to what the scan asked for and tell the caller to repeat it.
int scan_overflowed(const Image* a, Workspace* ws, bool* again) {
  *again = false;
  Image* am = const_cast<Image*>(a);
  if (ws->last_kernel != 2) return AM_OK;
  const uint64_t want = *reinterpret_cast<const uint64_t*>(ws->h_scalars + 48), total = *reinterpret_cast<const uint64_t*>(ws->h_scalars + 56);
  if (total * 16 > ws->last_span && ws->last_span >= (1u << 16)) {
    // a survivor flood: this text defeats the filter.  Hand the scan over to the per-segment walk (and remember it).
    am->walk_streak.fetch_add(1, std::memory_order_relaxed);
    if (!ws->last_inline) { ws->force_walk = true; *again = true; }   // (the inline form has verified everything itself: its results stand)
    return AM_OK;
  }
  am->walk_streak.store(0, std::memory_order_relaxed);
  const uint64_t regions = std::min<uint32_t>((uint32_t)sm_count(), SURV_REGIONS_MAX);
  if (ws->last_inline || ws->surv == nullptr || want <= ws->surv_bytes / 16 / regions * regions) return AM_OK;
  *again = true;                                   // some CTA's region was too short: grow the list to regions * (the fullest region) and a quarter
  return ws->need_surv(want + want / 4);
}

int scan_sync(const Image* a, Workspace* ws, const am_dev_text& t, int mode, cudaStream_t st) {
  for (int attempt = 0;; attempt++) {
    int rc = launch_scan(a, ws, t, mode, st);
    if (!rc) rc = read_scalars(ws, st);
    if (rc) return rc;
    if (mode == MODE_ANY && *reinterpret_cast<int*>(ws->h_scalars + 8) != 0) return AM_OK;   // a match is a match, however many survivors went unlisted
    bool again = false;
    if ((rc = scan_overflowed(a, ws, &again))) return rc;
    if (!again) return AM_OK;
    if (attempt >= 3) return fail(AM_E_INTERNAL, "survivor count kept growing");
  }
}

int read_scalars(Workspace* ws, cudaStream_t st, size_t bytes) {
  if (bytes < 64) bytes = 64;
  cudaError_t e = cudaMemcpyAsync(ws->h_scalars, ws->d_scalars, bytes, cudaMemcpyDeviceToHost, st);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyAsync(scalars)");
  e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return cuda_fail(e, "scan kernel");
  return AM_OK;
}

// Scan in EMIT mode and order the keys; on return ws->keys_b[0..*n) holds the sorted keys.
//  * filter kernel: the keys arrive in per-segment slots; bases = prefix sums of the fills; one local rank sort per segment
//    puts every key into its final place.  If some segment overflowed, segments + overflow area are compacted and
//    radix-sorted instead; if even the overflow area was too small the buffers grow and the scan runs again.
//  * walk kernel: global append + radix sort.
// Two phases, so that a caller can queue more work (the sharded calls: the all-gather of the counts) before the one host
// round trip:
//   emit_enqueue  launches the scan and -- for segmented emission -- the segment scan and the segment sort; afterwards
//                 d_scalars[0..8) = keys produced (global append) or keys that overflowed their segment (segmented) and
//                 d_scalars[8..16) = keys stored in the segments (segmented); their sum is the match count either way
//   emit_finish   reads the counters and handles the slow paths (overflow: compaction + radix sort; buffers too small: rescan)
int emit_enqueue(const Image* a, Workspace* ws, const am_dev_text& t, cudaStream_t st, am_match* matches, uint64_t matches_cap) {
  int rc = launch_scan(a, ws, t, MODE_EMIT, st);
  if (rc) return rc;
  if (ws->emit_segmented) {
    const size_t stb = seg_scan_temp_bytes(ws->num_segs);
    if ((rc = ws->need_sort_temp(stb))) return rc;
    cudaError_t e = launch_seg_scan(ws->sort_temp, stb, ws->seg_counts, ws->num_segs, ws->seg_cap, ws->seg_bases, st);
    if (e == cudaSuccess)   // d_scalars[8..16): number of keys stored in the segments
      e = cudaMemcpyAsync(ws->d_scalars + 8, ws->seg_bases + ws->num_segs, 8, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return cuda_fail(e, "segment scan");
    // queued before the host knows whether a segment overflowed (the common case: none did, and then this was the last
    // kernel of the call -- one host round trip in total); after an overflow its output is simply overwritten
    e = launch_seg_sort(ws->keys_a, ws->seg_counts, ws->seg_bases, ws->num_segs, ws->seg_cap, ws->keys_b, matches, matches_cap, a->host.rank_bits, a->dev.id_of_rank, st);
    if (e != cudaSuccess) return cuda_fail(e, "segment sort");
  }
  return AM_OK;
}

// The host has synchronised `st` and ws->h_scalars holds the counters of emit_enqueue.
int emit_finish(const Image* a, Workspace* ws, const am_dev_text& t, cudaStream_t st, uint64_t* n, am_match* matches, uint64_t matches_cap, bool* unpacked) {
  if (unpacked) *unpacked = false;
  const int end_bit = std::min(64, bitlen(t.text_len + t.pos_base) + (int)a->host.rank_bits);
  int rc = AM_OK;
  for (int attempt = 0;; attempt++) {
    const uint64_t cap = ws->keys_a_bytes / 8;
    bool again = false;
    if ((rc = scan_overflowed(a, ws, &again))) return rc;
    if (again) {                                   // the survivor list was too short: the keys are incomplete; scan again (the list has grown)
      if (attempt >= 3) return fail(AM_E_INTERNAL, "survivor count kept growing");
      if ((rc = emit_enqueue(a, ws, t, st, matches, matches_cap))) return rc;
      if ((rc = read_scalars(ws, st))) return rc;
      continue;
    }
    if (ws->emit_segmented) {
      const uint64_t n_ovf = *reinterpret_cast<uint64_t*>(ws->h_scalars), stored = *reinterpret_cast<uint64_t*>(ws->h_scalars + 8);
      *n = stored + n_ovf;
      if (n_ovf == 0) {
        if (unpacked && matches) *unpacked = true;
        return AM_OK;
      }
      if (n_ovf <= ws->ovf_cap && *n <= ws->keys_b_bytes / 8) {
        cudaError_t e = launch_seg_compact(ws->keys_a, ws->seg_counts, ws->seg_bases, ws->num_segs, ws->seg_cap, ws->keys_a + ws->ovf_base, n_ovf, stored, ws->keys_b, st);
        if (e != cudaSuccess) return cuda_fail(e, "segment compaction");
        size_t tb = sort_temp_bytes(*n, end_bit);
        if ((rc = ws->need_sort_temp(tb))) return rc;
        e = sort_keys(ws->sort_temp, tb, ws->keys_b, ws->keys_a, *n, end_bit, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(ws->keys_b, ws->keys_a, *n * 8, cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return cuda_fail(e, "radix sort");
        return AM_OK;
      }
      if (attempt >= 2) return fail(AM_E_INTERNAL, "match count kept growing");
      rc = ws->need_keys(2 * *n + 1024);   // half of the buffer is the overflow area: now it holds every key
    } else {
      *n = *reinterpret_cast<uint64_t*>(ws->h_scalars);
      if (*n <= cap) {
        if (*n == 0) return AM_OK;
        size_t tb = sort_temp_bytes(*n, end_bit);
        if ((rc = ws->need_sort_temp(tb))) return rc;
        cudaError_t e = sort_keys(ws->sort_temp, tb, ws->keys_a, ws->keys_b, *n, end_bit, st);
        if (e != cudaSuccess) return cuda_fail(e, "radix sort");
        return AM_OK;
      }
      if (attempt >= 2) return fail(AM_E_INTERNAL, "match count kept growing");
      rc = ws->need_keys(*n);              // exact size is now known: rescan
    }
    if (rc) return rc;
    if ((rc = emit_enqueue(a, ws, t, st, matches, matches_cap))) return rc;
    if ((rc = read_scalars(ws, st))) return rc;
  }
}

int find_all_sorted(const Image* a, Workspace* ws, const am_dev_text& t, cudaStream_t st, uint64_t* n, am_match* matches, uint64_t matches_cap, bool* unpacked) {
  const uint64_t span = t.text_len - std::min(t.report_begin, t.text_len);
  int rc = ws->need_keys(std::max<uint64_t>(1u << 16, span / 512));
  if (rc) return rc;
  if ((rc = emit_enqueue(a, ws, t, st, matches, matches_cap))) return rc;
  if ((rc = read_scalars(ws, st))) return rc;
  return emit_finish(a, ws, t, st, n, matches, matches_cap, unpacked);
}

// ---- L1 text substrate helpers (host) -----------------------------------------------------------------------------
static inline int dec(const uint8_t* d, int64_t i, int64_t n, uint32_t* cp) {  // decodeN, Utf8.hs:344-350
  uint32_t c0 = d[i];
  if (c0 < 0xC0) { *cp = c0; return 1; }
  uint32_t c1 = i + 1 < n ? d[i + 1] : 0;
  if (c0 < 0xE0) { *cp = ((c0 & 0x1F) << 6) | (c1 & 0x3F); return 2; }
  uint32_t c2 = i + 2 < n ? d[i + 2] : 0;
  if (c0 < 0xF0) { *cp = ((c0 & 0xF) << 12) | ((c1 & 0x3F) << 6) | (c2 & 0x3F); return 3; }
  uint32_t c3 = i + 3 < n ? d[i + 3] : 0;
  *cp = ((c0 & 7) << 18) | ((c1 & 0x3F) << 12) | ((c2 & 0x3F) << 6) | (c3 & 0x3F);
  return 4;
}
static inline int enc(uint32_t c, uint8_t* o) {  // unicode2utf8, Utf8.hs:154-160
  if (c < 0x80) { o[0] = (uint8_t)c; return 1; }
  if (c < 0x800) { o[0] = 0xC0 | (c >> 6); o[1] = 0x80 | (c & 0x3F); return 2; }
  if (c < 0x10000) { o[0] = 0xE0 | (c >> 12); o[1] = 0x80 | ((c >> 6) & 0x3F); o[2] = 0x80 | (c & 0x3F); return 3; }
  o[0] = 0xF0 | (c >> 18); o[1] = 0x80 | ((c >> 12) & 0x3F); o[2] = 0x80 | ((c >> 6) & 0x3F); o[3] = 0x80 | (c & 0x3F); return 4;
}

int lower_utf8_host(const LowerTable& lt, const uint8_t* in, int64_t len, std::vector<uint8_t>* out) {
  out->clear(); out->reserve((size_t)len + 8);
  int64_t i = 0;
  while (i < len) {
    uint32_t cp; int k = dec(in, i, len, &cp);
    uint8_t b[4]; int m = enc(lt.lower(cp), b);
    out->insert(out->end(), b, b + m);
    i += k;
  }
  return AM_OK;
}


// Host model of filter_kernel's two filter levels (am_filter.cu: fk_probe16 / fk_probe16_s2, fk_phase_a) on the host image,
// with the very cell / hash functions the kernel uses (am_internal.h).  IgnoreCase: `d` is the ORIGINAL text; both levels
// see it folded (every byte | 0x20), as in the kernel.
void host_filter_model(const HostAutomaton& H, const uint8_t* d, uint64_t n, uint32_t align, uint8_t* out_flags) {
  const bool ic = H.case_sensitivity == AM_IGNORE_CASE;
  auto byte_at = [&](uint64_t i) -> uint32_t {   // the kernel sees arbitrary bytes beyond the text: any value may only ADD candidates
    const uint32_t b = i < n ? d[i] : 0u;
    return ic ? (b | 0x20u) : b;
  };
  auto gram4 = [&](uint64_t i) -> uint32_t { return byte_at(i) | byte_at(i + 1) << 8 | byte_at(i + 2) << 16 | byte_at(i + 3) << 24; };
  const bool exact = H.t2_exact != 0;
  const uint32_t q = H.q;
  const int copies = filter_copies(q, exact);
  const int rowbits = filter_rowbits(copies);
  const uint32_t qmask = qgram_mask(q);
  for (uint64_t i = 0; i < n; i++) {
    uint32_t level1;
    if (q > 4) {
      // long q-gram: one cell per needle, two bits of its word, probed at every position
      uint32_t row, by, bt;
      long_cell((uint64_t)gram4(i) | ((uint64_t)gram4(i + 4) << 32), q, &row, &by, &bt);   // (long_cell reads bytes 0..3 and q-4..q-1)
      level1 = (H.filter[row] >> by) & (H.filter[row] >> bt) & 1u;
    } else if (filter_is_s2(q)) {
      // stride-2 cells: an even virtual position p tests cell A of its own 4-gram, the odd position p + 1 cell B; the row is
      // hashed from the 3 bytes the two share (for the odd position i these are text[i .. i+3), for the even one text[i+1 .. i+4))
      const bool odd = ((align + i) & 1) != 0;
      const uint32_t row = ((odd ? gram4(i) : gram4(i + 1)) * HASH_MUL_S2) >> (32 - rowbits);
      const uint32_t priv = odd ? byte_at(i + 3) : byte_at(i);                 // text[p + 4] resp. text[p]
      level1 = (H.filter[(size_t)row * copies] >> (31u - (priv & 31u))) & 1u;
    } else {
      uint32_t row, bit;
      filter_cell(gram4(i) & qmask, &row, &bit);
      level1 = (H.filter[(size_t)row * copies] >> bit) & 1u;
    }
    const uint32_t g = gram4(i) & qmask;
    uint32_t level2 = 0;
    if (exact) {
      uint32_t hb = t2_bucket(g);
      for (;;) {
        const uint32_t* slot = H.filter2.data() + 4 * (size_t)hb;
        uint32_t aux = 0; bool hit = false;
        if (slot[0] == g) { aux = slot[1]; hit = true; } else if (slot[2] == g) { aux = slot[3]; hit = true; }
        if (hit) { level2 = (aux & T2_AUX_ANY) ? 1u : (byte_at(i + q) == (aux & 0xFFu)); break; }
        if (!(slot[3] & T2_AUX_OVERFLOW)) break;
        hb = (hb + 1) & ((1u << T2_LOG2_BUCKETS) - 1);
      }
    } else if (q > 4) {
      const uint32_t hi = gram4(i + 4) & (q >= 8 ? 0xFFFFFFFFu : 0xFFFFu);
      const uint32_t b0 = gq_hash(g, hi) >> (32 - H.gbits_log2);
      level2 = (H.gbits[b0 >> 5] >> (b0 & 31)) & 1u;
    } else if (q == 4) {
      const uint32_t nb = byte_at(i + 4);
      const uint32_t ba = t2a_bit(g), bb = t2b_bit(g, nb), bc = t2c_bit(g, nb);
      level2 = ((H.filter2[T2A_WORD0 + (ba >> 5)] >> (ba & 31)) | ((H.filter2[T2B_WORD0 + (bb >> 5)] >> (bb & 31)) & (H.filter2[T2C_WORD0 + (bc >> 5)] >> (bc & 31)))) & 1u;
    } else {
      const uint32_t b2 = filter2_bit(g);
      level2 = (H.filter2[b2 >> 5] >> (b2 & 31)) & 1u;
    }
    uint32_t level3 = 1;                                       // (images without a third level pass everything)
    if (q == 4 && !exact && H.gbits_log2 != 0) {
      level3 = 0;
      const uint32_t hi = gram4(i + 4);
      for (uint32_t L = 4; L <= 8; L++) {
        const uint32_t b3 = gp_hash(g, L == 4 ? 0u : hi & (L == 8 ? 0xFFFFFFFFu : ((1u << (8 * (L - 4))) - 1u)), L) >> (32 - H.gbits_log2);
        level3 |= (H.gbits[b3 >> 5] >> (b3 & 31)) & 1u;
      }
    }
    out_flags[i] = (uint8_t)(level1 | (level2 << 1) | (level3 << 2));
  }
}

}  // namespace am

using namespace am;

// =====================================================================================================
// handle -> image of one case mode
// =====================================================================================================
namespace am {

// Build the host image of `cs`, pick its kernel and upload it to the handle's device.
static int build_image(const am_automaton* a, int cs, Image** out) {
  Image* im = new Image();
  const size_t n = a->needle_off.size() - 1;
  std::vector<am_u8slice> slices(n);
  for (size_t i = 0; i < n; i++) slices[i] = am_u8slice{a->needle_pool.data(), (int64_t)a->needle_off[i], (int64_t)(a->needle_off[i + 1] - a->needle_off[i])};
  am_lower_table lt{a->lower_pairs.data(), a->lower_pairs.size()};
  std::string err;
  int rc = build_host_automaton(slices.data(), n, cs, cs == AM_IGNORE_CASE ? &lt : nullptr, &im->host, &err);
  if (rc != AM_OK) { delete im; return fail(rc, err); }
  HostAutomaton& H = im->host;
  // The q-gram filter keeps its bitmaps in shared memory; beyond ~64 k distinct q-grams of the shortest form (q = 4) they
  // saturate and the per-segment walk is the better kernel.  Longer q-grams (needle sets whose shortest needle has >= 6
  // bytes) keep the filter selective far beyond that.  force_kernel overrides the heuristic.
  const int force = a->force_kernel;
  const bool filter_ok = H.q > 0;
  const bool filter_good = filter_ok && (H.q > 4 || H.filter_keys <= 65536);
  im->kernel_kind = (force == 2 && filter_ok) || (force != 1 && filter_good) ? 2 : 1;
  if (force == 2 && im->kernel_kind != 2) { delete im; return fail(AM_E_UNSUPPORTED, "filter kernel not applicable to this needle set"); }
  if (a->device < 0) { im->device = -1; *out = im; return AM_OK; }   // host image only (tests / introspection)

  DeviceGuard g;
  cudaError_t e = g.enter(a->device);
  if (e != cudaSuccess) { delete im; return cuda_fail(e, "cudaSetDevice"); }
  im->device = a->device;
  DevAutomaton& D = im->dev;
  std::memset(&D, 0, sizeof D);
  if (H.filter.empty()) { H.filter.assign(FILTER_WORDS, 0); }
  if (H.filter2.empty()) { H.filter2.assign(T2_WORDS, 0); }
  if (H.jump.empty()) { H.jump.assign(16, JumpSlot{0, NONE, 0, 0}); H.jump_mask = 15; }
  if (H.tails.empty()) H.tails.assign(4, 0);
  if (H.gbits.empty()) H.gbits.assign(4, 0);
  if ((rc = upload(im, H.dense, &D.dense)) || (rc = upload(im, H.fail, &D.fail)) || (rc = upload(im, H.edges, &D.edges)) ||
      (rc = upload(im, H.jump, &D.jump)) || (rc = upload(im, H.tails, &D.tails)) || (rc = upload(im, H.filter, &D.filter)) || (rc = upload(im, H.filter2, &D.filter2)) || (rc = upload(im, H.gbits, &D.gbits)) ||
      (rc = upload(im, H.own_off, &D.own_off)) || (rc = upload(im, H.own_rank, &D.own_rank)) ||
      (rc = upload(im, H.first_out, &D.first_out)) || (rc = upload(im, H.next_out, &D.next_out)) ||
      (rc = upload(im, H.chain_count, &D.chain_count)) || (rc = upload(im, H.id_of_rank, &D.id_of_rank)) ||
      (rc = upload(im, H.len_of_rank, &D.len_of_rank)) || (rc = upload(im, H.lower.stage1, &D.lower1)) ||
      (rc = upload(im, H.lower.stage2, &D.lower2)) || (rc = upload(im, H.cdfa, &D.cdfa)) ||
      (rc = upload(im, std::vector<uint8_t>(H.cls, H.cls + 256), &D.cls))) {
    delete im;
    return rc;
  }
  D.dense_states = H.dense_states; D.edge_mask = H.edge_mask; D.jump_mask = H.jump_mask;
  D.q = H.q; D.qmask = qgram_mask(H.q); D.min_len = H.min_len; D.max_len = H.max_len; D.rank_bits = H.rank_bits;
  D.num_states = H.num_states; D.num_needles = H.num_needles;
  D.ignore_case = cs == AM_IGNORE_CASE; D.halo = (uint32_t)H.halo_bytes;
  D.t2_exact = H.t2_exact; D.t2_empty_key = H.t2_empty_key; D.gbits_shift = 32 - H.gbits_log2;
  D.cdfa_states = H.cdfa_states; D.cdfa_shift = H.cdfa_shift;
  *out = im;
  return AM_OK;
}

int get_image(const am_automaton* ca, int cs, Image** out) {
  if (!ca) return fail(AM_E_BADARG, "automaton is null");
  if (cs != AM_CASE_SENSITIVE && cs != AM_IGNORE_CASE) return fail(AM_E_BADARG, "unknown case sensitivity");
  am_automaton* a = const_cast<am_automaton*>(ca);
  std::lock_guard<std::mutex> g(a->mu);
  if (!a->img[cs]) {
    if (cs == AM_IGNORE_CASE && !a->has_lower)
      return fail(AM_E_BADARG, "IgnoreCase needs the Char.toLower table (it may be empty): pass it to am_automaton_build");
    Image* im = nullptr;
    int rc = build_image(a, cs, &im);
    if (rc) return rc;
    a->img[cs] = im;
  }
  *out = a->img[cs];
  return AM_OK;
}

}  // namespace am

// Every compute entry point starts with this: image of the case mode, on its device (restored when `g` dies).
#define AM_ENTER(a, cs)                                   \
  AM_NVTX(__func__);                                      \
  Image* im = nullptr;                                    \
  DeviceGuard guard;                                      \
  { int rc_ = get_image((a), (cs), &im); if (rc_) return rc_; rc_ = check_ready(im, &guard); if (rc_) return rc_; }

// =====================================================================================================
// C ABI
// =====================================================================================================
extern "C" {

const char* am_last_error(void) { return g_last_error.c_str(); }
size_t am_last_error_copy(char* buf, size_t cap) {
  const size_t n = g_last_error.size();
  if (buf && cap) { const size_t k = std::min(n, cap - 1); std::memcpy(buf, g_last_error.data(), k); buf[k] = 0; }
  return n;
}
int am_abi_version(void) { return AM_ABI_VERSION; }
int am_device_count(void) { return usable_device_count(); }

int am_automaton_build(const am_u8slice* needles, size_t n, const am_lower_table* lower, const am_options* opts, am_automaton** out) {
  if (!out) return fail(AM_E_BADARG, "out is null");
  *out = nullptr;
  if (n > 0 && !needles) return fail(AM_E_BADARG, "needles is null");
  if (n >= (1ull << 31)) return fail(AM_E_BADARG, "too many needles");
  if (lower && lower->n > 0 && !lower->pairs) return fail(AM_E_BADARG, "bad lower table");
  const int want_dev = opts ? opts->device : -1;
  const int force = opts ? opts->force_kernel : 0;
  if (force < 0 || force > 2) return fail(AM_E_BADARG, "unknown force_kernel");
  am_automaton* a = new am_automaton();
  a->force_kernel = force;
  a->needle_off.assign(n + 1, 0);
  uint64_t total = 0;
  for (size_t i = 0; i < n; i++) {
    if (needles[i].len < 0 || needles[i].off < 0 || (needles[i].len > 0 && !needles[i].ptr)) { delete a; return fail(AM_E_BADARG, "bad needle slice"); }
    if ((uint64_t)needles[i].len >= (1ull << 24)) { delete a; return fail(AM_E_BADARG, "needle longer than 16 MiB"); }
    total += (uint64_t)needles[i].len;
    a->needle_off[i + 1] = total;
  }
  a->needle_pool.resize(total + 1);
  for (size_t i = 0; i < n; i++)
    if (needles[i].len) std::memcpy(a->needle_pool.data() + a->needle_off[i], needles[i].ptr + needles[i].off, (size_t)needles[i].len);
  if (lower) {
    a->has_lower = true;
    for (size_t i = 0; i < lower->n; i++) {
      if (lower->pairs[i].from_cp >= 0x110000 || lower->pairs[i].to_cp >= 0x110000) { delete a; return fail(AM_E_BADARG, "bad lower table"); }
      a->lower_pairs.push_back(lower->pairs[i]);
    }
  }
  if (want_dev == -2) { a->device = -1; *out = a; return AM_OK; }   // host images only (tests / introspection)
  if (usable_device_count() == 0) { delete a; return fail(AM_E_NODEVICE, "no sm_100 CUDA device; libam_b200 has no CPU fallback"); }
  int dev = want_dev;
  if (dev < 0) { if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; } }
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess || major != 10) {
    cudaGetLastError(); delete a;
    return fail(AM_E_NODEVICE, "device " + std::to_string(dev) + " is not an sm_100 device");
  }
  a->device = dev;
  *out = a;
  return AM_OK;
}

void am_automaton_free(am_automaton* a) { delete a; }

int am_automaton_prepare(const am_automaton* a, int cs) {
  Image* im = nullptr;
  return get_image(a, cs, &im);
}

int am_automaton_info(const am_automaton* a, int cs, uint64_t* num_states, uint64_t* max_needle_bytes, uint64_t* halo_bytes, int* kernel_kind) {
  Image* im = nullptr;
  int rc = get_image(a, cs, &im);
  if (rc) return rc;
  if (num_states) *num_states = im->host.num_states;
  if (max_needle_bytes) *max_needle_bytes = im->host.max_len;
  if (halo_bytes) *halo_bytes = im->host.halo_bytes;
  if (kernel_kind) *kernel_kind = im->kernel_kind;
  return AM_OK;
}

static int check_slice(const am_u8slice* s) {
  if (!s) return fail(AM_E_BADARG, "text slice is null");
  if (s->len < 0 || s->off < 0 || (s->len > 0 && !s->ptr)) return fail(AM_E_BADARG, "bad text slice");
  return AM_OK;
}
static int check_dev_text(const am_dev_text* t) {
  if (!t) return fail(AM_E_BADARG, "device text is null");
  return AM_OK;
}

// Host model of filter_kernel's two filter levels on the host image (am_filter_model.h: the very functions the kernel uses).
int am_debug_host_filter(const am_automaton* a, int cs, const am_u8slice* text, uint32_t align, uint8_t* out_flags) {
  int rc = check_slice(text); if (rc) return rc;
  if (!out_flags) return fail(AM_E_BADARG, "null argument");
  Image* im = nullptr;
  if ((rc = get_image(a, cs, &im))) return rc;
  const HostAutomaton& H = im->host;
  if (H.q == 0 || H.filter.empty()) return fail(AM_E_UNSUPPORTED, "this automaton has no q-gram filter");
  host_filter_model(H, text->ptr + text->off, (uint64_t)text->len, align, out_flags);
  return AM_OK;
}

// ---- device-resident entry points -------------------------------------------------------------------------
int am_count_matches_dev(const am_automaton* a, int cs, const am_dev_text* t, void* stream, uint64_t* out_count) {
  int rc = check_dev_text(t); if (rc) return rc;
  if (!out_count) return fail(AM_E_BADARG, "out_count is null");
  AM_ENTER(a, cs);
  Workspace* ws = acquire_ws(im); if (!ws) return fail(AM_E_OOM, "workspace");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  rc = scan_sync(im, ws, *t, MODE_COUNT, st);
  if (!rc) *out_count = *reinterpret_cast<uint64_t*>(ws->h_scalars);
  release_ws(im, ws);
  return rc;
}

int am_contains_any_dev(const am_automaton* a, int cs, const am_dev_text* t, void* stream, int* out_bool) {
  int rc = check_dev_text(t); if (rc) return rc;
  if (!out_bool) return fail(AM_E_BADARG, "out_bool is null");
  AM_ENTER(a, cs);
  Workspace* ws = acquire_ws(im); if (!ws) return fail(AM_E_OOM, "workspace");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // window by window (1 GiB of end positions each, with the halo before it): the fold's `Done` -- the scan stops at the first
  // window that holds a match
  *out_bool = 0;
  const uint64_t W = 1ull << 30, halo = im->host.halo_bytes;
  for (uint64_t b = std::min(t->report_begin, t->text_len); !rc && b < t->text_len && !*out_bool; b += W) {
    const uint64_t e = std::min(t->text_len, b + W), w = b > halo ? b - halo : 0;
    am_dev_text win{static_cast<const uint8_t*>(t->dev_text) + w, e - w, b - w, t->pos_base + w};
    rc = scan_sync(im, ws, win, MODE_ANY, st);
    if (!rc) *out_bool = *reinterpret_cast<int*>(ws->h_scalars + 8) != 0;
  }
  release_ws(im, ws);
  return rc;
}

// Shared tail of am_find_all_dev / am_find_all_sharded: the host has the counters; finish, check the capacity, unpack.
static int finish_find_all_dev(const Image* im, Workspace* ws, const am_dev_text& t, cudaStream_t st, am_match* dev_out, size_t cap, uint64_t* n_found) {
  uint64_t n = 0;
  bool unpacked = false;
  int rc = emit_finish(im, ws, t, st, &n, dev_out, dev_out ? cap : 0, &unpacked);
  if (rc) return rc;
  *n_found = n;
  if (n > cap) return fail(AM_E_OVERFLOW, "output buffer too small");
  if (n > 0 && !unpacked) {
    if (!dev_out) return fail(AM_E_BADARG, "dev_out is null");
    cudaError_t e = launch_unpack(ws->keys_b, n, im->host.rank_bits, im->dev.id_of_rank, dev_out, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return cuda_fail(e, "unpack");
  }
  return AM_OK;
}

int am_find_all_dev(const am_automaton* a, int cs, const am_dev_text* t, void* stream, am_match* dev_out, size_t cap, uint64_t* n_found) {
  int rc = check_dev_text(t); if (rc) return rc;
  if (!n_found) return fail(AM_E_BADARG, "n_found is null");
  AM_ENTER(a, cs);
  Workspace* ws = acquire_ws(im); if (!ws) return fail(AM_E_OOM, "workspace");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uint64_t span = t->text_len - std::min(t->report_begin, t->text_len);
  rc = ws->need_keys(std::max<uint64_t>(1u << 16, span / 512));
  if (!rc) rc = emit_enqueue(im, ws, *t, st, dev_out, dev_out ? cap : 0);
  if (!rc) rc = read_scalars(ws, st);
  if (!rc) rc = finish_find_all_dev(im, ws, *t, st, dev_out, cap, n_found);
  release_ws(im, ws);
  return rc;
}

// ---- multi-GPU: scan + all-gather of the counts, one host round trip (am_comm.cu holds the communicator) ---------------
// The value a rank gathers carries bit 63 when its filter scan listed more survivors than its list holds (its counters
// are then incomplete).  Every rank sees every value, so all of them take the same decision: repeat the round (the rank
// concerned has grown its list) until no flag is raised -- no rank is ever alone in a collective.
constexpr uint64_t SHARD_RETRY = 1ull << 63;
static bool shard_round_done(const Image* a, const am_comm* c, Workspace* ws, int* rc) {
  const uint64_t* g = reinterpret_cast<const uint64_t*>(ws->h_scalars + 64);
  bool retry = false;
  for (int i = 0; i < comm_size(c); i++) retry = retry || (g[i] & SHARD_RETRY) != 0;
  bool again = false;
  *rc = scan_overflowed(a, ws, &again);            // grows this rank's list (or hands over to the walk kernel) when it was the one
  return !retry || *rc != AM_OK;
}

int am_count_sharded(const am_automaton* a, int cs, am_comm* c, const am_dev_text* shard, void* stream, am_shard_result* out) {
  int rc = check_dev_text(shard); if (rc) return rc;
  if (!out || !c) return fail(AM_E_BADARG, "null argument");
  AM_ENTER(a, cs);
  if ((rc = comm_check(c, im->device))) return rc;
  Workspace* ws = acquire_ws(im); if (!ws) return fail(AM_E_OOM, "workspace");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  unsigned long long* ds = reinterpret_cast<unsigned long long*>(ws->d_scalars);
  for (int round = 0; !rc; round++) {
    rc = launch_scan(im, ws, *shard, MODE_COUNT, st);
    if (!rc && launch_shard_total(ds, 0, 1, ds + 6, ws->last_kernel == 2 && !ws->last_inline ? ws->surv_bytes / 16 / std::min<uint32_t>((uint32_t)sm_count(), SURV_REGIONS_MAX) * std::min<uint32_t>((uint32_t)sm_count(), SURV_REGIONS_MAX) : 0, ds + 4, st) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "count kernel");
    if (!rc) rc = comm_allgather_u64(c, ws->d_scalars + 32, ws->d_scalars + 64, st);       // this rank's count -> every rank's
    if (!rc) rc = read_scalars(ws, st, 64 + 8 * (size_t)comm_size(c));
    if (rc || shard_round_done(im, c, ws, &rc)) break;
    if (round >= 3) rc = fail(AM_E_INTERNAL, "survivor count kept growing");
  }
  if (!rc) comm_offsets(c, reinterpret_cast<const uint64_t*>(ws->h_scalars + 64), out);
  release_ws(im, ws);
  return rc;
}

int am_contains_any_sharded(const am_automaton* a, int cs, am_comm* c, const am_dev_text* shard, void* stream, int* out_bool) {
  int rc = check_dev_text(shard); if (rc) return rc;
  if (!out_bool || !c) return fail(AM_E_BADARG, "null argument");
  AM_ENTER(a, cs);
  if ((rc = comm_check(c, im->device))) return rc;
  Workspace* ws = acquire_ws(im); if (!ws) return fail(AM_E_OOM, "workspace");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  unsigned long long* ds = reinterpret_cast<unsigned long long*>(ws->d_scalars);
  for (int round = 0; !rc; round++) {
    rc = launch_scan(im, ws, *shard, MODE_ANY, st);                                        // d_scalars[8..16): this shard's flag (0 / 1)
    if (!rc && launch_shard_total(ds, 1, 1, ds + 6, ws->last_kernel == 2 && !ws->last_inline ? ws->surv_bytes / 16 / std::min<uint32_t>((uint32_t)sm_count(), SURV_REGIONS_MAX) * std::min<uint32_t>((uint32_t)sm_count(), SURV_REGIONS_MAX) : 0, ds + 4, st) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "count kernel");
    if (!rc) rc = comm_allgather_u64(c, ws->d_scalars + 32, ws->d_scalars + 64, st);
    if (!rc) rc = read_scalars(ws, st, 64 + 8 * (size_t)comm_size(c));
    if (rc || shard_round_done(im, c, ws, &rc)) break;
    if (round >= 3) rc = fail(AM_E_INTERNAL, "survivor count kept growing");
  }
  if (!rc) {
    am_shard_result r;
    comm_offsets(c, reinterpret_cast<const uint64_t*>(ws->h_scalars + 64), &r);
    *out_bool = r.total != 0;
  }
  release_ws(im, ws);
  return rc;
}

int am_find_all_sharded(const am_automaton* a, int cs, am_comm* c, const am_dev_text* shard, void* stream, am_match* dev_out, size_t cap, am_shard_result* out) {
  int rc = check_dev_text(shard); if (rc) return rc;
  if (!out || !c) return fail(AM_E_BADARG, "null argument");
  AM_ENTER(a, cs);
  if ((rc = comm_check(c, im->device))) return rc;
  Workspace* ws = acquire_ws(im); if (!ws) return fail(AM_E_OOM, "workspace");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uint64_t span = shard->text_len - std::min(shard->report_begin, shard->text_len);
  rc = ws->need_keys(std::max<uint64_t>(1u << 16, span / 512));
  unsigned long long* ds = reinterpret_cast<unsigned long long*>(ws->d_scalars);
  // scan, verification, segment scan and segment sort are queued; the match count is d_scalars[0..8) + d_scalars[8..16) whatever
  // path the ordering takes afterwards, so the all-gather goes on the stream right behind them and the host waits ONCE
  for (int round = 0; !rc; round++) {
    rc = emit_enqueue(im, ws, *shard, st, dev_out, dev_out ? cap : 0);
    if (!rc && launch_shard_total(ds, 0, ws->emit_segmented ? 2 : 1, ds + 6, ws->last_kernel == 2 && !ws->last_inline ? ws->surv_bytes / 16 / std::min<uint32_t>((uint32_t)sm_count(), SURV_REGIONS_MAX) * std::min<uint32_t>((uint32_t)sm_count(), SURV_REGIONS_MAX) : 0, ds + 4, st) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "count kernel");
    if (!rc) rc = comm_allgather_u64(c, ws->d_scalars + 32, ws->d_scalars + 64, st);
    if (!rc) rc = read_scalars(ws, st, 64 + 8 * (size_t)comm_size(c));
    if (rc || shard_round_done(im, c, ws, &rc)) break;
    if (round >= 3) rc = fail(AM_E_INTERNAL, "survivor count kept growing");
  }
  if (!rc) {
    comm_offsets(c, reinterpret_cast<const uint64_t*>(ws->h_scalars + 64), out);
    uint64_t n = 0;
    rc = finish_find_all_dev(im, ws, *shard, st, dev_out, cap, &n);                        // (a rescan for more key space never changes the count)
    if (n != out->n_local && (rc == AM_OK || rc == AM_E_OVERFLOW)) rc = fail(AM_E_INTERNAL, "sharded match count changed between scan and ordering");
  }
  release_ws(im, ws);
  return rc;
}

// ---- host-buffer entry points ---------------------------------------------------------------------------------
static int upload_text(Workspace* ws, const am_u8slice& hay, cudaStream_t st, am_dev_text* t) {
  int rc = ws->need_text((uint64_t)hay.len);
  if (rc) return rc;
  if (hay.len > 0) {
    cudaError_t e = cudaMemcpyAsync(ws->text, hay.ptr + hay.off, (size_t)hay.len, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyAsync(H2D text)");
  }
  t->dev_text = ws->text; t->text_len = (uint64_t)hay.len; t->report_begin = 0; t->pos_base = 0;
  return AM_OK;
}

// COUNT / ANY over a host buffer: the text goes up in chunks on a copy stream while the previous chunk is scanned on a
// second stream (chunk k is scanned with `halo_bytes` of warm-up before it and reports the matches that END inside it,
// exactly like a shard), so the scan hides behind the PCIe transfer -- and `containsAny` stops uploading at the first
// chunk that holds a match: the `Done` of the reference's fold (Searcher.hs:156-164) reaches all the way to the bus.
}  // extern "C" (the chunk loop is a template)

constexpr uint64_t HOST_CHUNK = 64ull << 20;
constexpr uint64_t HOST_REGISTER_MIN = 256ull << 20;

// A host text that is not page-locked (a GHC pinned ByteArray#, malloc, numpy) would make every cudaMemcpyAsync a staged,
// synchronous copy.  Large texts are page-locked in place for the duration of the call; AM_HOST_REGISTER=0 turns it off.
struct HostPin {
  void* base = nullptr;
  ~HostPin() { if (base) cudaHostUnregister(base); }
  void pin(const uint8_t* p, uint64_t len) {
    static const bool off = []() { const char* v = std::getenv("AM_HOST_REGISTER"); return v && std::atoi(v) == 0; }();
    if (off || len < HOST_REGISTER_MIN) return;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return; }
    if (at.type != cudaMemoryTypeUnregistered) return;        // already page-locked (cudaHostAlloc / registered) or managed
    const uintptr_t lo = reinterpret_cast<uintptr_t>(p) & ~uintptr_t(4095), hi = (reinterpret_cast<uintptr_t>(p) + len + 4095) & ~uintptr_t(4095);
    if (cudaHostRegister(reinterpret_cast<void*>(lo), hi - lo, cudaHostRegisterReadOnly) == cudaSuccess) base = reinterpret_cast<void*>(lo);
    else if (cudaGetLastError(), cudaHostRegister(reinterpret_cast<void*>(lo), hi - lo, cudaHostRegisterDefault) == cudaSuccess) base = reinterpret_cast<void*>(lo);
    else cudaGetLastError();                                  // could not pin: the copies are staged by the driver instead
  }
};

// Calls fn(window, stream) for every chunk in order; fn returns < 0 to stop early (no error), 0 to go on, > 0 = error code.
template <class F>
static int for_each_host_chunk(const Image* a, Workspace* ws, const am_u8slice& hay, F fn) {
  const uint64_t len = (uint64_t)hay.len;
  int rc = ws->need_text(len);
  if (rc) return rc;
  if (len <= 2 * HOST_CHUNK) {                                 // small text: one upload, one window
    am_dev_text t;
    rc = upload_text(ws, hay, 0, &t);
    if (!rc) rc = fn(t, (cudaStream_t)0);
    return rc < 0 ? AM_OK : rc;
  }
  if (!ws->copy_stream) {
    if (cudaStreamCreateWithFlags(&ws->copy_stream, cudaStreamNonBlocking) != cudaSuccess || cudaStreamCreateWithFlags(&ws->scan_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ws->copy_done[0], cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&ws->copy_done[1], cudaEventDisableTiming) != cudaSuccess)
      return cuda_fail(cudaGetLastError(), "stream / event creation");
  }
  const uint8_t* src = hay.ptr + hay.off;
  HostPin pin;
  pin.pin(src, len);
  const uint64_t halo = a->host.halo_bytes;
  const uint64_t chunks = (len + HOST_CHUNK - 1) / HOST_CHUNK;
  auto upload = [&](uint64_t k) -> cudaError_t {
    const uint64_t b = k * HOST_CHUNK, e = std::min(len, b + HOST_CHUNK);
    cudaError_t err = cudaMemcpyAsync(ws->text + b, src + b, e - b, cudaMemcpyHostToDevice, ws->copy_stream);
    if (err == cudaSuccess) err = cudaEventRecord(ws->copy_done[k & 1], ws->copy_stream);
    return err;
  };
  cudaError_t e = upload(0);
  for (uint64_t k = 0; k < chunks && e == cudaSuccess; k++) {
    if (k + 1 < chunks) e = upload(k + 1);                     // in flight while chunk k is scanned
    if (e != cudaSuccess) break;
    e = cudaStreamWaitEvent(ws->scan_stream, ws->copy_done[k & 1], 0);
    if (e != cudaSuccess) break;
    const uint64_t b = k * HOST_CHUNK, end = std::min(len, b + HOST_CHUNK), w = b > halo ? b - halo : 0;
    am_dev_text t{ws->text + w, end - w, b - w, w};
    rc = fn(t, ws->scan_stream);                               // fn synchronises scan_stream before it returns (it reads results back):
    if (rc) break;                                             // that also orders the reuse of copy_done[k & 1] by chunk k + 2
  }
  cudaStreamSynchronize(ws->copy_stream);                      // an early exit leaves one upload in flight: the buffer is reused next call
  cudaStreamSynchronize(ws->scan_stream);
  if (e != cudaSuccess) return cuda_fail(e, "pipelined upload");
  return rc < 0 ? AM_OK : rc;
}

static int scan_host_pipelined(const Image* a, Workspace* ws, const am_u8slice& hay, int mode, uint64_t* out_count, int* out_any) {
  if (out_count) *out_count = 0;
  if (out_any) *out_any = 0;
  return for_each_host_chunk(a, ws, hay, [&](const am_dev_text& t, cudaStream_t st) -> int {
    int rc = scan_sync(a, ws, t, mode, st);
    if (rc) return rc;
    if (out_count) *out_count += *reinterpret_cast<uint64_t*>(ws->h_scalars);
    if (out_any && *reinterpret_cast<int*>(ws->h_scalars + 8) != 0) { *out_any = 1; return -1; }
    return 0;
  });
}

extern "C" {

int am_count_matches(const am_automaton* a, int cs, const am_u8slice* hay, uint64_t* out_count) {
  int rc = check_slice(hay); if (rc) return rc;
  if (!out_count) return fail(AM_E_BADARG, "out_count is null");
  AM_ENTER(a, cs);
  Workspace* ws = acquire_ws(im); if (!ws) return fail(AM_E_OOM, "workspace");
  rc = scan_host_pipelined(im, ws, *hay, MODE_COUNT, out_count, nullptr);
  release_ws(im, ws);
  return rc;
}

int am_contains_any(const am_automaton* a, int cs, const am_u8slice* hay, int* out_bool) {
  int rc = check_slice(hay); if (rc) return rc;
  if (!out_bool) return fail(AM_E_BADARG, "out_bool is null");
  AM_ENTER(a, cs);
  Workspace* ws = acquire_ws(im); if (!ws) return fail(AM_E_OOM, "workspace");
  rc = scan_host_pipelined(im, ws, *hay, MODE_ANY, nullptr, out_bool);
  release_ws(im, ws);
  return rc;
}

int am_find_all(const am_automaton* a, int cs, const am_u8slice* hay, am_match* out, size_t cap, uint64_t* n_found) {
  int rc = check_slice(hay); if (rc) return rc;
  if (!n_found) return fail(AM_E_BADARG, "n_found is null");
  AM_ENTER(a, cs);
  Workspace* ws = acquire_ws(im); if (!ws) return fail(AM_E_OOM, "workspace");
  // chunk by chunk (the lists of consecutive chunks concatenate: matches are partitioned by end position); each chunk's
  // records travel back while the next chunk is scanned and the one after that is uploaded
  uint64_t total = 0;
  rc = for_each_host_chunk(im, ws, *hay, [&](const am_dev_text& t, cudaStream_t st) -> int {
    uint64_t n = 0;
    int r = find_all_sorted(im, ws, t, st, &n);
    if (r) return r;
    if (n > 0 && total + n <= cap) {
      if (!out) return fail(AM_E_BADARG, "out is null");
      if ((r = ws->need_matches(n))) return r;
      cudaError_t e = launch_unpack(ws->keys_b, n, im->host.rank_bits, im->dev.id_of_rank, ws->matches, st);
      if (e == cudaSuccess) e = cudaMemcpyAsync(out + total, ws->matches, n * sizeof(am_match), cudaMemcpyDeviceToHost, st);
      if (e != cudaSuccess) return cuda_fail(e, "unpack / D2H");
    }
    total += n;                                              // (keeps counting past `cap`: the caller learns the size it needs)
    return 0;
  });
  if (!rc) {
    cudaError_t e = cudaStreamSynchronize(0);
    if (e != cudaSuccess) rc = cuda_fail(e, "unpack / D2H");
    *n_found = total;
    if (!rc && total > cap) rc = fail(AM_E_OVERFLOW, "output buffer too small");
  }
  release_ws(im, ws);
  return rc;
}

int am_contains_all(const am_automaton* a, int cs, const am_u8slice* hay, int* out_bool) {
  // Searcher.containsAll (Searcher.hs:173-187): the set of outstanding needle ids starts full; every match removes its
  // needle; the answer is "the set is empty" and the fold stops as soon as it is.  Here the set is a bit per needle rank in
  // HBM: after each chunk's scan a kernel ORs in the ranks of the chunk's matches and counts the ranks still missing,
  // and the upload stops at the first chunk after which none is.
  int rc = check_slice(hay); if (rc) return rc;
  if (!out_bool) return fail(AM_E_BADARG, "out_bool is null");
  AM_ENTER(a, cs);
  const uint64_t nn = im->host.num_needles;
  if (nn == 0) { *out_bool = 1; return AM_OK; }
  Workspace* ws = acquire_ws(im); if (!ws) return fail(AM_E_OOM, "workspace");
  rc = ws->need_seen(nn);
  if (!rc && cudaMemsetAsync(ws->seen_bits, 0, (nn + 31) / 32 * 4, 0) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "cudaMemsetAsync");
  if (!rc && cudaStreamSynchronize(0) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "cudaMemsetAsync");
  uint64_t missing = nn;
  if (!rc)
    rc = for_each_host_chunk(im, ws, *hay, [&](const am_dev_text& t, cudaStream_t st) -> int {
      uint64_t n = 0;
      int r = find_all_sorted(im, ws, t, st, &n);
      if (r) return r;
      if (n == 0) return 0;
      unsigned int* d_missing = reinterpret_cast<unsigned int*>(ws->d_scalars + 40);
      cudaError_t e = launch_mark_seen(ws->keys_b, n, im->host.rank_bits, ws->seen_bits, (uint32_t)nn, d_missing, st);
      if (e == cudaSuccess) e = cudaMemcpyAsync(ws->h_scalars + 40, d_missing, 4, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) return cuda_fail(e, "needle bit set");
      missing = *reinterpret_cast<unsigned int*>(ws->h_scalars + 40);
      return missing == 0 ? -1 : 0;                            // every needle seen: stop uploading (Searcher.hs:180-181)
    });
  if (!rc) *out_bool = missing == 0;
  release_ws(im, ws);
  return rc;
}

int am_shard_plan(uint64_t text_len, uint64_t halo_bytes, uint32_t n_shards, uint32_t r, uint64_t* warm_begin, uint64_t* begin, uint64_t* end) {
  if (n_shards == 0 || r >= n_shards || !warm_begin || !begin || !end) return fail(AM_E_BADARG, "bad shard plan arguments");
  // contiguous ranges, 16-byte aligned cut points (the kernels read 16-byte granules)
  auto cut = [&](uint32_t i) -> uint64_t {
    if (i >= n_shards) return text_len;
    unsigned __int128 p = (unsigned __int128)text_len * i / n_shards;
    return (uint64_t)p & ~15ull;
  };
  *begin = cut(r); *end = cut(r + 1);
  *warm_begin = *begin > halo_bytes ? *begin - halo_bytes : 0;
  return AM_OK;
}

int am_lower_utf8(const am_lower_table* lower, const am_u8slice* text, uint8_t* out, size_t cap, uint64_t* out_len) {
  int rc = check_slice(text); if (rc) return rc;
  if (!out_len) return fail(AM_E_BADARG, "out_len is null");
  LowerTable lt;
  if ((rc = build_lower_table(lower, &lt))) return fail(rc, "bad lower table");
  std::vector<uint8_t> v;
  lower_utf8_host(lt, text->ptr + text->off, text->len, &v);
  *out_len = v.size();
  if (v.size() > cap) return fail(AM_E_OVERFLOW, "output buffer too small");
  if (!v.empty()) { if (!out) return fail(AM_E_BADARG, "out is null"); std::memcpy(out, v.data(), v.size()); }
  return AM_OK;
}

int am_skip_code_points_backwards(const am_u8slice* text, int64_t index0, int64_t n0, int64_t* out_index) {
  // Utf8.hs:256-276
  int rc = check_slice(text); if (rc) return rc;
  if (!out_index) return fail(AM_E_BADARG, "out_index is null");
  if (index0 >= text->len) return fail(AM_E_BADARG, "Invalid use of skipCodePointsBackwards");
  const uint8_t* d = text->ptr;
  int64_t index = index0 + text->off, n = n0;
  for (;;) {
    if (index >= 0 && (d[index] & 0xC0) == 0x80) { index--; continue; }
    if (n == 0) {
      if (index < 0) return fail(AM_E_BADARG, "Invalid use of skipCodePointsBackwards");
      *out_index = index - text->off; return AM_OK;
    }
    if (index < 0) return fail(AM_E_BADARG, "Invalid use of skipCodePointsBackwards");
    index--; n--;
  }
}

void am_free(void* p) { std::free(p); }
void am_dev_free(void* p) { if (p) cudaFree(p); }

int am_profile_enable(int on) { g_profile.store(on ? 1 : 0); return AM_OK; }
int am_profile_last_scan_ms(float* ms) {
  if (!ms) return fail(AM_E_BADARG, "ms is null");
  if (!g_ev0 || g_ev_device < 0) return fail(AM_E_BADARG, "no profiled scan on this thread");
  DeviceGuard g;
  cudaError_t e = g.enter(g_ev_device);
  if (e == cudaSuccess) e = cudaEventSynchronize(g_ev1);
  if (e == cudaSuccess) e = cudaEventElapsedTime(ms, g_ev0, g_ev1);
  return e == cudaSuccess ? AM_OK : cuda_fail(e, "event timing");
}
uint64_t am_profile_kernel_launches(void) { return g_kernel_launches.load(); }
uint64_t am_replacer_last_passes(void) { return g_last_passes; }
uint64_t am_replacer_last_rescans(void) { return g_last_rescans; }
int am_replacer_last_profile(float* ms, uint64_t* bytes_moved) {
  if (ms) *ms = g_last_replacer_ms;
  if (bytes_moved) *bytes_moved = g_last_replacer_bytes;
  return AM_OK;
}

// ---- synthetic workloads ----------------------------------------------------------------------------------------------
int am_synth_fill_dev(void* dev_buf, uint64_t len, uint64_t first, uint64_t seed, const uint8_t* alphabet, uint32_t alphabet_len, void* stream) {
  if (usable_device_count() == 0) return fail(AM_E_NODEVICE, "no sm_100 CUDA device");
  if (!alphabet || alphabet_len == 0 || alphabet_len > 256 || (len && !dev_buf)) return fail(AM_E_BADARG, "bad synth arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* d_alpha = nullptr;
  cudaError_t e = cudaMalloc((void**)&d_alpha, 256);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
  e = cudaMemcpyAsync(d_alpha, alphabet, alphabet_len, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = launch_synth_fill(static_cast<uint8_t*>(dev_buf), len, first, seed, d_alpha, alphabet_len, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d_alpha);
  return e == cudaSuccess ? AM_OK : cuda_fail(e, "synth fill");
}

int am_synth_plant_dev(void* dev_buf, uint64_t len, uint64_t first, uint64_t seed, const am_u8slice* needles, size_t n, uint32_t block, void* stream) {
  if (usable_device_count() == 0) return fail(AM_E_NODEVICE, "no sm_100 CUDA device");
  if (!needles || n == 0 || block < 64 || (len && !dev_buf)) return fail(AM_E_BADARG, "bad synth arguments");
  std::vector<uint8_t> bytes; std::vector<uint32_t> off(n + 1, 0);
  for (size_t i = 0; i < n; i++) {
    if (needles[i].len < 0 || needles[i].len > 16) return fail(AM_E_BADARG, "planted needles must be at most 16 bytes");
    bytes.insert(bytes.end(), needles[i].ptr + needles[i].off, needles[i].ptr + needles[i].off + needles[i].len);
    off[i + 1] = (uint32_t)bytes.size();
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* d_b = nullptr; uint32_t* d_o = nullptr;
  cudaError_t e = cudaMalloc((void**)&d_b, bytes.size() + 16);
  if (e == cudaSuccess) e = cudaMalloc((void**)&d_o, off.size() * 4);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_b, bytes.data(), bytes.size(), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_o, off.data(), off.size() * 4, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = launch_synth_plant(static_cast<uint8_t*>(dev_buf), len, first, seed, d_b, d_o, (uint32_t)n, block, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (d_b) cudaFree(d_b);
  if (d_o) cudaFree(d_o);
  return e == cudaSuccess ? AM_OK : cuda_fail(e, "synth plant");
}

}  // extern "C"
