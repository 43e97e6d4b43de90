/*
 * am_b200.h -- C ABI of libam_b200.so, the B200-native (sm_100a) replacement for the hot
 * path of channable/alfred-margaret:
 *
 *   Data.Text.AhoCorasick.Automaton.build / runText / runLower / runWithCase
 *   Data.Text.AhoCorasick.Searcher.build / containsAny (/ containsAll)
 *   Data.Text.AhoCorasick.Replacer.build / run / runWithLimit
 *
 * The reference is a pure Haskell library with no plugin API; its only FFI precedent is the
 * Rust comparison benchmark
 *   foreign import ccall unsafe "perform_ac" :: CBool -> CSize -> Ptr U8Slice -> Ptr U8Slice -> IO CSize
 * (benchmark/rust-ffi/app/Main.hs:28-29, U8Slice :32-46; libacbench/src/lib.rs:5-11, :24-30).
 * This header follows that convention: plain pointers and sizes, (ptr, off, len) text slices
 * over pinned arrays, the caller owns its memory, every call returns an int status.
 * No torch / CUDA types appear in any signature (streams and device pointers are void*).
 *
 * All `file:line` citations are relative to the reference repository root.
 * INTEGRATION.md shows the Haskell `foreign import` stubs that bind these symbols.
 */
#ifndef AM_B200_H
#define AM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AM_ABI_VERSION 1

/* ---- status codes ---------------------------------------------------------------------- */
enum {
  AM_OK = 0,
  AM_E_BADARG = 1,      /* null pointer, negative length, unknown enum value */
  AM_E_OOM = 2,         /* host or device allocation failed */
  AM_E_CUDA = 3,        /* CUDA runtime error; am_last_error() has the text */
  AM_E_OVERFLOW = 4,    /* caller's output buffer too small; the needed size was written */
  AM_E_NODEVICE = 5,    /* no CUDA device / not an sm_100 device: there is NO CPU fallback */
  AM_E_UNSUPPORTED = 6, /* input outside the reference's contract (e.g. empty needle in an IgnoreCase Replacer) */
  AM_E_INTERNAL = 7
};

/* `data CaseSensitivity = CaseSensitive | IgnoreCase`, src/Data/Text/CaseSensitivity.hs:14-16 */
enum { AM_CASE_SENSITIVE = 0, AM_IGNORE_CASE = 1 };

/* An unpacked `Text u8data off len` (UTF-8 bytes).  Same layout as the reference's
 * `U8Slice` (benchmark/rust-ffi/app/Main.hs:32-46): the text is ptr[off .. off+len). */
typedef struct am_u8slice {
  const uint8_t *ptr;
  int64_t off;
  int64_t len;
} am_u8slice;

/* `Match { matchPos :: CodeUnitIndex, matchValue :: v }`, Automaton.hs:98-105.
 * end_pos is the byte offset one past the last byte of the match, relative to the start of
 * the text slice.  needle_id is the index of the needle in the list given to
 * am_automaton_build; the host maps it to the caller's payload `v`. */
typedef struct am_match {
  uint64_t end_pos;
  uint32_t needle_id;
  uint32_t reserved; /* always 0 */
} am_match;

/* `Data.Char.toLower` (GHC base) as DATA: the simple lower-case mapping differs between GHC
 * versions (Unicode tables), so the host passes the non-identity pairs above ASCII.
 * Utf8.hs:145-151.  ASCII A-Z is lowered by the library itself (Utf8.hs:131-135). */
typedef struct am_lower_pair {
  uint32_t from_cp;
  uint32_t to_cp;
} am_lower_pair;
typedef struct am_lower_table {
  const am_lower_pair *pairs;
  size_t n;
} am_lower_table;

typedef struct am_options {
  int32_t device;         /* CUDA device ordinal; -1 = current device */
  int32_t force_kernel;   /* 0 = auto, 1 = per-segment goto/failure walk, 2 = q-gram filter + goto verify */
  uint64_t reserved[6];
} am_options;

typedef struct am_automaton am_automaton; /* AcMachine + case flag (Automaton.hs:108-123, Searcher.hs:61-66) */
typedef struct am_replacer am_replacer;   /* Replacer (Replacer.hs:78-80) */

const char *am_last_error(void); /* thread-local text of the last failure */
int am_abi_version(void);
/* Number of usable sm_100 devices (0 => every compute entry point returns AM_E_NODEVICE). */
int am_device_count(void);

/* ---- Automaton.build (Automaton.hs:176-200) / Searcher.build (Searcher.hs:110-118) -------
 * Copies the needles.  For AM_IGNORE_CASE the caller has already lower-cased them, as in the
 * reference (Automaton.hs:543-546, Searcher.hs:107-118); `lower` is then required (it may have
 * n == 0).  The returned handle is immutable and may be shared between threads. */
int am_automaton_build(const am_u8slice *needles, size_t n, int case_sensitivity,
                       const am_lower_table *lower, const am_options *opts, am_automaton **out);
void am_automaton_free(am_automaton *a);
/* Introspection (used by tests and by the shard planner). */
int am_automaton_info(const am_automaton *a, uint64_t *num_states, uint64_t *max_needle_bytes,
                      uint64_t *halo_bytes, int *kernel_kind);
/* Introspection (host image only, no device): the q-gram filter of the fast path evaluated on the host for every start
 * position of `text`, exactly as filter_kernel evaluates it for a text whose device address is `align` (0..15) modulo
 * 16.  out_flags[i] bit 0: the shared-memory bitmap passes position i; bit 1: the second level passes it too.  The
 * filter may pass positions that start no needle, never the reverse -- the property the CPU tests check.  Returns
 * AM_E_UNSUPPORTED when the automaton has no filter (empty needle). */
int am_debug_host_filter(const am_automaton *a, am_u8slice text, uint32_t align, uint8_t *out_flags);

/* ---- host-buffer entry points (the drop-in calls; H2D/D2H inside) -------------------------
 * Searcher.containsAny, Searcher.hs:156-164. */
int am_contains_any(const am_automaton *a, am_u8slice hay, int *out_bool);
/* `runText 0 (\n _ -> Step (n + 1))`, benchmark/haskell/app/Main.hs:67-76 (runLower when the
 * automaton was built with AM_IGNORE_CASE). */
int am_count_matches(const am_automaton *a, am_u8slice hay, uint64_t *out_count);
/* All matches, in the order runWithCase (Automaton.hs:442-534) hands them to its fold:
 * end_pos ascending; at one end_pos longest needle first, later-inserted duplicate first
 * (:263, :373-376).  On AM_E_OVERFLOW *n_found holds the required capacity. */
int am_find_all(const am_automaton *a, am_u8slice hay, am_match *out, size_t cap, uint64_t *n_found);
/* Searcher.containsAll, Searcher.hs:173-187 (needle ids = list indices, buildNeedleIdSearcher :167-169). */
int am_contains_all(const am_automaton *a, am_u8slice hay, int *out_bool);

/* ---- device-resident entry points ---------------------------------------------------------
 * `dev_text` points at device memory holding text bytes [0, text_len).  Only matches whose
 * end_pos lies in (report_begin, text_len] are reported, and `pos_base` is added to every
 * reported position: a shard of a larger haystack passes its halo in [0, report_begin).
 * `stream` is a cudaStream_t (NULL = the legacy default stream); calls return after the
 * stream has drained.  am_match / count outputs marked `dev_` live in device memory. */
typedef struct am_dev_text {
  const void *dev_text;
  uint64_t text_len;
  uint64_t report_begin;
  uint64_t pos_base;
} am_dev_text;

int am_count_matches_dev(const am_automaton *a, am_dev_text t, void *stream, uint64_t *out_count);
int am_contains_any_dev(const am_automaton *a, am_dev_text t, void *stream, int *out_bool);
int am_find_all_dev(const am_automaton *a, am_dev_text t, void *stream, am_match *dev_out, size_t cap,
                    uint64_t *n_found);

/* Shard planner for multi-GPU / multi-window scans (pure function, no device needed):
 * shard r of n covers report range (begin, end] and must be resident from warm_begin. */
int am_shard_plan(uint64_t text_len, uint64_t halo_bytes, uint32_t n_shards, uint32_t r,
                  uint64_t *warm_begin, uint64_t *begin, uint64_t *end);

/* ---- Replacer (Replacer.hs) -----------------------------------------------------------------
 * Replacer.build (:97-116): pair i has priority -i.  For AM_IGNORE_CASE the needles are passed
 * ORIGINAL-cased; the library lowers them with `lower` exactly as `Utf8.lowerUtf8` would
 * (:105-107) and keeps the original byte / code point lengths for the payload (:111-113). */
int am_replacer_build(const am_u8slice *needles, const am_u8slice *replacements, size_t n,
                      int case_sensitivity, const am_lower_table *lower, const am_options *opts,
                      am_replacer **out);
void am_replacer_free(am_replacer *r);
/* Replacer.runWithLimit (:203-242).  max_len = UINT64_MAX is `run` (:200-201).  *exceeded = 1 is
 * `Nothing`.  The result buffer is allocated by the library; release it with am_free. */
int am_replacer_run(const am_replacer *r, am_u8slice hay, uint64_t max_len, uint8_t **out,
                    uint64_t *out_len, int *exceeded);
/* Number of scan passes the last am_replacer_run on this thread performed. */
uint64_t am_replacer_last_passes(void);
/* Number of those passes that scanned the whole text: 1 when the match list could be carried from pass to pass
 * (CaseSensitive, no empty needle: only the neighbourhood of each replacement is rescanned), else every pass. */
uint64_t am_replacer_last_rescans(void);
void am_free(void *p);

/* ---- L1 text substrate used by the wrappers (Utf8.hs) ------------------------------------------
 * Host-side helpers so a binding does not need its own copies. */
/* Utf8.lowerUtf8 (:138-140); returns AM_E_OVERFLOW with *out_len = needed size. */
int am_lower_utf8(const am_lower_table *lower, am_u8slice text, uint8_t *out, size_t cap, uint64_t *out_len);
/* Utf8.skipCodePointsBackwards (:256-276); AM_E_BADARG where the reference calls `error`. */
int am_skip_code_points_backwards(am_u8slice text, int64_t index, int64_t n, int64_t *out_index);

/* ---- measurement hooks (bench.py) -------------------------------------------------------------------
 * am_profile_enable(1) makes every scan record CUDA events around its scan kernel(s) on the stream they
 * are launched on; am_profile_last_scan_ms returns the device time of the most recent scan on this
 * thread (kernel only: no sort, no copies).  am_profile_kernel_launches counts every kernel this
 * library has launched in the process (its own kernels, not the CUB sort's). */
int am_profile_enable(int on);
int am_profile_last_scan_ms(float *ms);
uint64_t am_profile_kernel_launches(void);

/* ---- synthetic workload generator (bench / test tooling; BASELINE.json configs) ----------------
 * Counter-based: byte i depends only on (seed, i), so host and device produce identical text. */
int am_synth_fill_dev(void *dev_buf, uint64_t len, uint64_t first_byte_index, uint64_t seed,
                      const uint8_t *alphabet, uint32_t alphabet_len, void *stream);
int am_synth_plant_dev(void *dev_buf, uint64_t len, uint64_t first_byte_index, uint64_t seed,
                       const am_u8slice *needles, size_t n, uint32_t block, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* AM_B200_H */
