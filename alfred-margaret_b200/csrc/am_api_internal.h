// am_api_internal.h -- handle types and helpers shared by am_api.cu, am_replacer.cu and am_comm.cu.
#pragma once

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>   // header-only; a no-op unless a profiler (nsys, ncu --nvtx) injects itself

#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "am_device.cuh"
#include "am_kernels.h"

namespace am {

// Per-call scratch in HBM, pooled per image so concurrent host threads never share one.
struct Workspace {
  unsigned char* d_scalars = nullptr;   // [0..8) count, [8..12) flag, rest: see launch_scan / the sharded calls
  unsigned char* h_scalars = nullptr;   // pinned mirror
  uint64_t* keys_a = nullptr; size_t keys_a_bytes = 0;
  uint64_t* keys_b = nullptr; size_t keys_b_bytes = 0;
  void* sort_temp = nullptr; size_t sort_temp_bytes_ = 0;
  uint8_t* text = nullptr; size_t text_bytes = 0;
  am_match* matches = nullptr; size_t matches_bytes = 0;
  uint8_t* aux_a = nullptr; size_t aux_a_bytes = 0;   // lowered copy of the text (IgnoreCase automata that cannot take the one-pass form)
  uint8_t* aux_b = nullptr; size_t aux_b_bytes = 0;
  // segmented emission of the filter kernel (set by launch_scan in EMIT mode; emit_segmented = false: global append, e.g. walk kernel)
  uint32_t* seg_counts = nullptr; size_t seg_counts_bytes = 0;
  uint64_t* seg_bases = nullptr; size_t seg_bases_bytes = 0;
  uint32_t* seen_bits = nullptr; size_t seen_bytes = 0;        // containsAll: bit per needle rank
  ulonglong2* surv = nullptr; size_t surv_bytes = 0;           // filter scan: the survivor list {text index, 8 text bytes} between filter_kernel and verify_kernel
  unsigned long long* surv_counts = nullptr;                   // ... and its per-CTA counters (SURV_REGIONS_MAX)
  cudaStream_t copy_stream = nullptr, scan_stream = nullptr;   // host-buffer scans: upload chunk k + 1 while chunk k is scanned
  cudaEvent_t copy_done[2] = {nullptr, nullptr};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;                    // profiling: around the scan kernel(s) of the last launch_scan
  bool emit_segmented = false; uint64_t num_segs = 0; uint32_t seg_cap = 0; uint64_t ovf_base = 0, ovf_cap = 0;
  int last_kernel = 0;                                         // kernel the last launch_scan really ran: 1 = walk, 2 = filter
  uint64_t last_span = 0;                                      // bytes the last launch_scan reported on
  bool last_inline = false;                                    // ... and whether the filter kernel verified its survivors itself (no list to outgrow)
  bool force_walk = false;                                     // the next launch_scan takes the walk kernel (hand-over after a survivor flood)
  ~Workspace();
  int need_keys(uint64_t n);
  int need_sort_temp(size_t bytes);
  int need_text(uint64_t n);
  int need_matches(uint64_t n);
  int need_aux(uint64_t a_bytes, uint64_t b_bytes);
  int need_segs(uint64_t n);
  int need_seen(uint64_t bits);
  int need_surv(uint64_t entries);
};

// One case mode of an automaton: the host image and its copy in HBM.  Immutable once built.
struct Image {
  HostAutomaton host;
  DevAutomaton dev;
  int device = -1;            // -1: host image only
  int kernel_kind = 1;        // 1 = per-segment walk, 2 = q-gram filter + goto verify
  std::vector<void*> dev_allocs;
  std::mutex ws_mutex;
  std::vector<Workspace*> ws_pool;
  // Survivor-rate monitor of the filter scan: texts on which the filter passes more than 1 / 16 of the positions (needles
  // made of the text's most frequent q-grams, "aaaa..." against "aaaa") are handed over to the per-segment walk, whose cost
  // does not depend on the text -- the reference's loop is O(n) on any input (Automaton.hs:489-510).  After two
  // hand-overs in a row the filter is only tried on every eighth scan.
  std::atomic<int> walk_streak{0};
  std::atomic<unsigned> scans{0};
  ~Image();
};

// Remembers the caller's current device and puts it back: the library never leaks a cudaSetDevice into the embedding
// application (a torch or Haskell host with several GPUs).
struct DeviceGuard {
  int prev = -1; bool active = false;
  DeviceGuard() {}
  cudaError_t enter(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
    active = true;
    return prev == dev ? cudaSuccess : cudaSetDevice(dev);
  }
  ~DeviceGuard() { if (active && prev >= 0) { int cur = -1; if (cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev); } }
};

// NVTX range around an ABI entry point (ranges at the C-ABI entry points), closed when the scope ends.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};
#define AM_NVTX(name) ::am::NvtxRange nvtx_range_(name)

extern thread_local std::string g_last_error;
extern thread_local uint64_t g_last_passes, g_last_rescans;
extern thread_local float g_last_replacer_ms;
extern thread_local uint64_t g_last_replacer_bytes;
inline int bitlen(uint64_t x) { int b = 0; while (x) { b++; x >>= 1; } return b; }
int fail(int code, const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);
bool profiling_enabled();
// The image of one case mode of a handle (built and uploaded on first use).
int get_image(const struct ::am_automaton* a, int cs, Image** out);
// Image present on a device?  Enters its device through `g`.
int check_ready(const Image* a, DeviceGuard* g);
Workspace* acquire_ws(const Image* a);
void release_ws(const Image* a, Workspace* w);
int launch_scan(const Image* a, Workspace* ws, const am_dev_text& t, int mode, cudaStream_t st);
int read_scalars(Workspace* ws, cudaStream_t st, size_t bytes = 64);
int scan_overflowed(const Image* a, Workspace* ws, bool* again);
int emit_enqueue(const Image* a, Workspace* ws, const am_dev_text& t, cudaStream_t st, am_match* matches, uint64_t matches_cap);
int emit_finish(const Image* a, Workspace* ws, const am_dev_text& t, cudaStream_t st, uint64_t* n, am_match* matches, uint64_t matches_cap, bool* unpacked);
// `matches` (nullable, device): when the keys could be ordered by the per-segment sort, the am_match records are written
// in the same kernel and *unpacked is set; otherwise the caller unpacks ws->keys_b itself.
int find_all_sorted(const Image* a, Workspace* ws, const am_dev_text& t, cudaStream_t st, uint64_t* n,
                    am_match* matches = nullptr, uint64_t matches_cap = 0, bool* unpacked = nullptr);
void host_filter_model(const HostAutomaton& H, const uint8_t* text, uint64_t n, uint32_t align, uint8_t* out_flags);
// am_comm.cu: the communicator behind am_comm_* (NCCL, loaded at run time)
int comm_check(struct ::am_comm* c, int device);
int comm_size(const struct ::am_comm* c);
int comm_allgather_u64(struct ::am_comm* c, const void* d_send /* 8 bytes */, void* d_recv /* 8 * nranks bytes */, cudaStream_t st);
void comm_offsets(const struct ::am_comm* c, const uint64_t* counts, am_shard_result* out);
int lower_utf8_host(const LowerTable& lt, const uint8_t* in, int64_t len, std::vector<uint8_t>* out);

}  // namespace am

// AcMachine (Automaton.hs:108-123): the needles and (optionally) the host's Char.toLower table; the image of a case
// mode is built when that mode is first used.  `runText` and `runLower` run the SAME machine (:539-553).
struct am_automaton {
  std::vector<uint8_t> needle_pool;             // all needle bytes, back to back
  std::vector<uint64_t> needle_off;             // n + 1 offsets into needle_pool
  std::vector<am_lower_pair> lower_pairs;
  bool has_lower = false;
  int device = -1;                              // resolved device ordinal, -1 = host images only
  int force_kernel = 0;
  std::mutex mu;
  am::Image* img[2] = {nullptr, nullptr};
  ~am_automaton() { delete img[0]; delete img[1]; }
};
