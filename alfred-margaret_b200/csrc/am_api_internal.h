// am_api_internal.h -- handle types and helpers shared by am_api.cu and am_replacer.cu.
#pragma once

#include <cuda_runtime.h>

#include <mutex>
#include <string>
#include <vector>

#include "am_device.cuh"
#include "am_kernels.h"

namespace am {

// Per-call scratch in HBM, pooled per automaton so concurrent host threads never share one.
struct Workspace {
  unsigned char* d_scalars = nullptr;   // [0..8) count, [8..12) flag, rest: replacer scalars
  unsigned char* h_scalars = nullptr;   // pinned mirror
  uint64_t* keys_a = nullptr; size_t keys_a_bytes = 0;
  uint64_t* keys_b = nullptr; size_t keys_b_bytes = 0;
  void* sort_temp = nullptr; size_t sort_temp_bytes_ = 0;
  uint8_t* text = nullptr; size_t text_bytes = 0;
  am_match* matches = nullptr; size_t matches_bytes = 0;
  uint8_t* aux_a = nullptr; size_t aux_a_bytes = 0;   // replacer ping-pong text buffers
  uint8_t* aux_b = nullptr; size_t aux_b_bytes = 0;
  // segmented emission of the filter kernel (set by launch_scan in EMIT mode; emit_segmented = false: global append, e.g. walk kernel)
  uint32_t* seg_counts = nullptr; size_t seg_counts_bytes = 0;
  uint64_t* seg_bases = nullptr; size_t seg_bases_bytes = 0;
  cudaStream_t copy_stream = nullptr, scan_stream = nullptr;   // host-buffer scans: upload chunk k + 1 while chunk k is scanned
  cudaEvent_t copy_done[2] = {nullptr, nullptr};
  bool emit_segmented = false; uint64_t num_segs = 0; uint32_t seg_cap = 0; uint64_t ovf_base = 0, ovf_cap = 0;
  ~Workspace();
  int need_keys(uint64_t n);
  int need_sort_temp(size_t bytes);
  int need_text(uint64_t n);
  int need_matches(uint64_t n);
  int need_aux(uint64_t a_bytes, uint64_t b_bytes);
  int need_segs(uint64_t n);
};

extern thread_local std::string g_last_error;
extern thread_local uint64_t g_last_passes, g_last_rescans;
inline int bitlen(uint64_t x) { int b = 0; while (x) { b++; x >>= 1; } return b; }
int fail(int code, const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);
int check_ready(const struct ::am_automaton* a);
Workspace* acquire_ws(const struct ::am_automaton* a);
void release_ws(const struct ::am_automaton* a, Workspace* w);
int launch_scan(const struct ::am_automaton* a, Workspace* ws, const am_dev_text& t, int mode, cudaStream_t st);
// `matches` (nullable, device): when the keys could be ordered by the per-segment sort, the am_match records are written
// in the same kernel and *unpacked is set; otherwise the caller unpacks ws->keys_b itself.
int find_all_sorted(const struct ::am_automaton* a, Workspace* ws, const am_dev_text& t, cudaStream_t st, uint64_t* n,
                    am_match* matches = nullptr, uint64_t matches_cap = 0, bool* unpacked = nullptr);
int lower_utf8_host(const LowerTable& lt, const uint8_t* in, int64_t len, std::vector<uint8_t>* out);

}  // namespace am

struct am_automaton {
  am::HostAutomaton host;
  am::DevAutomaton dev;
  int device = -1;            // -1: host image only
  int kernel_kind = 1;        // 1 = per-segment walk, 2 = q-gram filter + goto verify
  std::vector<void*> dev_allocs;
  std::mutex ws_mutex;
  std::vector<am::Workspace*> ws_pool;
};
