/*
 * am_b200.h -- C ABI of libam_b200.so, the B200-native (sm_100a) replacement for the hot
 * path of channable/alfred-margaret:
 *
 *   Data.Text.AhoCorasick.Automaton.build / runText / runLower / runWithCase
 *   Data.Text.AhoCorasick.Searcher.build / containsAny / containsAll
 *   Data.Text.AhoCorasick.Replacer.build / run / runWithLimit
 *
 * The reference is a pure Haskell library with no plugin API; its only FFI precedent is the
 * Rust comparison benchmark
 *   foreign import ccall unsafe "perform_ac" :: CBool -> CSize -> Ptr U8Slice -> Ptr U8Slice -> IO CSize
 * (benchmark/rust-ffi/app/Main.hs:28-29, U8Slice :32-46; libacbench/src/lib.rs:5-11, :24-30).
 * This header follows that convention exactly: plain pointers and sizes, (ptr, off, len) text
 * slices over pinned arrays passed BY POINTER (`Ptr U8Slice`: GHC's FFI cannot pass a struct by
 * value), the caller owns its memory, every call returns an int status.
 * No torch / CUDA / NCCL types appear in any signature (streams and device pointers are void*).
 *
 * ABI version 2 (round 2): every struct argument is a pointer; one automaton handle serves BOTH
 * case modes, as the reference's `AcMachine` does -- `build` (Automaton.hs:176) knows nothing of
 * case, `runText` / `runLower` (:539-553) pick it per call, `Searcher.setCaseSensitivity`
 * (Searcher.hs:142-145) flips a flag without rebuilding -- so the case mode is an argument of the
 * scan calls; the sharded (multi-GPU) calls and their communicator are part of the ABI.
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

#define AM_ABI_VERSION 2

/* ---- status codes ---------------------------------------------------------------------- */
enum {
  AM_OK = 0,
  AM_E_BADARG = 1,      /* null pointer, negative length, unknown enum value */
  AM_E_OOM = 2,         /* host or device allocation failed */
  AM_E_CUDA = 3,        /* CUDA / NCCL runtime error; am_last_error() has the text */
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
  int32_t device;         /* CUDA device ordinal; -1 = current device; -2 = host image only (introspection, no device) */
  int32_t force_kernel;   /* 0 = auto, 1 = per-segment goto/failure walk, 2 = q-gram filter + goto verify */
  uint64_t reserved[6];
} am_options;

typedef struct am_automaton am_automaton; /* AcMachine (Automaton.hs:108-123): case-agnostic, like the reference's */
typedef struct am_replacer am_replacer;   /* Replacer (Replacer.hs:78-80) */
typedef struct am_comm am_comm;           /* one rank's end of the multi-GPU exchange */

const char *am_last_error(void); /* thread-local text of the last failure */
/* The same text copied into the caller's buffer (always NUL-terminated).  A `ccall safe` from an UNBOUND Haskell
 * thread may come back on another OS thread: bind the call and this read together (`runInBoundThread`). */
size_t am_last_error_copy(char *buf, size_t cap);
int am_abi_version(void);
/* Number of usable sm_100 devices (0 => every compute entry point returns AM_E_NODEVICE). */
int am_device_count(void);

/* ---- Automaton.build (Automaton.hs:176-200) / Searcher.build (Searcher.hs:110-118) -------
 * Copies the needles.  The handle is case-agnostic: the scan calls below take the case mode, and the device image
 * of a mode is built the first time that mode is used (or by am_automaton_prepare).  For AM_IGNORE_CASE scans the
 * caller has already lower-cased the needles, as in the reference (Automaton.hs:543-546, Searcher.hs:107-118), and
 * `lower` (the host's Char.toLower table; it may have n == 0) must have been given here; `lower` == NULL makes a
 * handle that only serves AM_CASE_SENSITIVE.  The handle is immutable for its users and may be shared between threads. */
int am_automaton_build(const am_u8slice *needles, size_t n, const am_lower_table *lower, const am_options *opts,
                       am_automaton **out);
void am_automaton_free(am_automaton *a);
/* Build (and upload) the image of one case mode now instead of at its first use; returns what that build returns. */
int am_automaton_prepare(const am_automaton *a, int case_sensitivity);
/* Introspection (used by tests and by the shard planner).  kernel_kind: 1 = walk, 2 = q-gram filter. */
int am_automaton_info(const am_automaton *a, int case_sensitivity, uint64_t *num_states, uint64_t *max_needle_bytes,
                      uint64_t *halo_bytes, int *kernel_kind);
/* Introspection (host image only, no device): the q-gram filter of the fast path evaluated on the host for every start
 * position of `text`, exactly as filter_kernel evaluates it for a text whose device address is `align` (0..15) modulo
 * 16.  out_flags[i] bit 0: the shared-memory bitmap passes position i; bit 1: the second level passes it too.  The
 * filter may pass positions that start no needle, never the reverse -- the property the CPU tests check.  For
 * AM_IGNORE_CASE `text` is the ORIGINAL (not lowered) text.  Returns AM_E_UNSUPPORTED when the automaton has no
 * filter (empty needle). */
int am_debug_host_filter(const am_automaton *a, int case_sensitivity, const am_u8slice *text, uint32_t align, uint8_t *out_flags);

/* ---- host-buffer entry points (the drop-in calls; H2D/D2H inside) -------------------------
 * `case_sensitivity` selects runText (AM_CASE_SENSITIVE) or runLower (AM_IGNORE_CASE), Automaton.hs:539-553.
 * The host buffer may be pageable (a GHC pinned ByteArray# is not page-locked): texts of 256 MiB and more are
 * page-locked in place for the duration of the call (cudaHostRegister) unless they already are.
 * Searcher.containsAny, Searcher.hs:156-164. */
int am_contains_any(const am_automaton *a, int case_sensitivity, const am_u8slice *hay, int *out_bool);
/* `runText 0 (\n _ -> Step (n + 1))`, benchmark/haskell/app/Main.hs:67-76; tests/Data/Text/AhoCorasickSpec.hs:252-261. */
int am_count_matches(const am_automaton *a, int case_sensitivity, const am_u8slice *hay, uint64_t *out_count);
/* All matches, in the order runWithCase (Automaton.hs:442-534) hands them to its fold:
 * end_pos ascending; at one end_pos longest needle first, later-inserted duplicate first
 * (:263, :373-376).  On AM_E_OVERFLOW *n_found holds the required capacity. */
int am_find_all(const am_automaton *a, int case_sensitivity, const am_u8slice *hay, am_match *out, size_t cap, uint64_t *n_found);
/* Searcher.containsAll, Searcher.hs:173-187 (needle ids = list indices, buildNeedleIdSearcher :167-169): a bit set of the
 * needle ids seen so far lives on the device; the upload stops at the first chunk after which none is outstanding. */
int am_contains_all(const am_automaton *a, int case_sensitivity, const am_u8slice *hay, int *out_bool);

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

int am_count_matches_dev(const am_automaton *a, int case_sensitivity, const am_dev_text *t, void *stream, uint64_t *out_count);
int am_contains_any_dev(const am_automaton *a, int case_sensitivity, const am_dev_text *t, void *stream, int *out_bool);
int am_find_all_dev(const am_automaton *a, int case_sensitivity, const am_dev_text *t, void *stream, am_match *dev_out, size_t cap,
                    uint64_t *n_found);

/* ---- multi-GPU: one process (or thread) per GPU, contiguous shards -------------------------
 * Shard planner (pure function, no device needed): shard r of n covers report range (begin, end] and must be
 * resident from warm_begin. */
int am_shard_plan(uint64_t text_len, uint64_t halo_bytes, uint32_t n_shards, uint32_t r,
                  uint64_t *warm_begin, uint64_t *begin, uint64_t *end);
/* The path's only exchange is one 64-bit match count per rank (-> each rank's offset into the global match list and
 * the total).  It runs over NCCL, which the library loads at run time (dlopen "libnccl.so.2": no link dependency, and
 * inside a process that already holds an NCCL -- torch -- that copy is the one used).  Rank 0 makes the id
 * (ncclGetUniqueId), the host distributes its AM_COMM_ID_BYTES bytes by whatever channel it has (a file, MPI, a
 * socket, torch.distributed's store), every rank calls am_comm_init. */
#define AM_COMM_ID_BYTES 128
int am_comm_unique_id(uint8_t *id /* AM_COMM_ID_BYTES */);
int am_comm_init(int rank, int nranks, const uint8_t *id, int device, am_comm **out);
void am_comm_free(am_comm *c);
typedef struct am_shard_result {
  uint64_t n_local;        /* matches (or needle hits) of this rank's shard */
  uint64_t global_offset;  /* sum of n_local over the ranks before this one */
  uint64_t total;          /* sum over all ranks */
} am_shard_result;
/* Scan this rank's shard (halo in [0, report_begin), positions rebased by pos_base) and all-gather the counts: the scan,
 * the ordering of the matches and the collective are queued on `stream` back to back and the host waits once. */
int am_count_sharded(const am_automaton *a, int case_sensitivity, am_comm *c, const am_dev_text *shard, void *stream, am_shard_result *out);
int am_find_all_sharded(const am_automaton *a, int case_sensitivity, am_comm *c, const am_dev_text *shard, void *stream,
                        am_match *dev_out, size_t cap, am_shard_result *out);
/* containsAny over all shards: any(count > 0) of the same all-gather. */
int am_contains_any_sharded(const am_automaton *a, int case_sensitivity, am_comm *c, const am_dev_text *shard, void *stream, int *out_bool);
/* Shards that are already device-resident WITHOUT their halo: rank r sends the last `halo_bytes` of its shard to rank
 * r + 1, which receives them into dev_buf[0, halo_bytes) in front of its own shard at dev_buf + halo_bytes
 * (ncclSend / ncclRecv over NVLink; rank 0's front is left untouched and must not be reported: report_begin = halo). */
int am_shard_halo_exchange(am_comm *c, void *dev_buf, uint64_t halo_bytes, uint64_t shard_len, void *stream);
/* Small host-side collectives over the same communicator, for hosts without one of their own (the sharded Replacer
 * passes, SURVEY.md section 8e: MAX of the pass priority, SUM of the output lengths). op: 0 = sum, 1 = max, 2 = min. */
int am_comm_allreduce_u64(am_comm *c, uint64_t *value, int op, void *stream);

/* ---- Replacer (Replacer.hs) -----------------------------------------------------------------
 * Replacer.build (:97-116): pair i has priority -i.  For AM_IGNORE_CASE the needles are passed
 * ORIGINAL-cased; the library lowers them with `lower` exactly as `Utf8.lowerUtf8` would
 * (:105-107) and keeps the original byte / code point lengths for the payload (:111-113).
 * `lower` may be given for AM_CASE_SENSITIVE too: it is what a later IgnoreCase run of this replacer uses. */
int am_replacer_build(const am_u8slice *needles, const am_u8slice *replacements, size_t n,
                      int case_sensitivity, const am_lower_table *lower, const am_options *opts,
                      am_replacer **out);
/* The stored form of a Replacer -- what `compose` (:120-133), `mapReplacement` (:136-141) and the derived FromJSON
 * instance (:78-86) start from: the needles exactly as the searcher holds them (nothing is lowered here) and the
 * Payload lengths of the original needles (:111-113).  `case_sensitivity` only names the image to build eagerly. */
int am_replacer_build_stored(const am_u8slice *stored_needles, const uint32_t *len_bytes, const uint32_t *len_code_points,
                             const am_u8slice *replacements, size_t n, int case_sensitivity, const am_lower_table *lower,
                             const am_options *opts, am_replacer **out);
void am_replacer_free(am_replacer *r);
/* Replacer.runWithLimit (:203-242).  max_len = UINT64_MAX is `run` (:200-201).  *exceeded = 1 is `Nothing`.
 * `case_sensitivity` is the mode of THIS run: `setCaseSensitivity` (:151-153) only flips the flag of the searcher
 * and keeps the stored needles (lowered or not, as they were built) and payload lengths, so the host shim keeps
 * that flag and passes it here.  The result buffer is allocated by the library; release it with am_free. */
int am_replacer_run(const am_replacer *r, int case_sensitivity, const am_u8slice *hay, uint64_t max_len, uint8_t **out,
                    uint64_t *out_len, int *exceeded);
/* The same on a device-resident text; the result is a device buffer of the library's (release with am_dev_free). */
int am_replacer_run_dev(const am_replacer *r, int case_sensitivity, const void *dev_text, uint64_t text_len, uint64_t max_len,
                        void *stream, void **dev_out, uint64_t *out_len, int *exceeded);
/* Number of scan passes the last am_replacer_run on this thread performed. */
uint64_t am_replacer_last_passes(void);
/* Number of those passes that scanned the whole text: 1 when the match list could be carried from pass to pass
 * (no empty needle: only the neighbourhood of each replacement is rescanned), else every pass. */
uint64_t am_replacer_last_rescans(void);
/* Device time of the passes of the last am_replacer_run on this thread (ms, CUDA events; needs am_profile_enable(1))
 * and the bytes they read + wrote (scans, carried match lists, rewritten text tiles): the Replacer's roofline terms. */
int am_replacer_last_profile(float *ms, uint64_t *bytes_moved);
void am_free(void *p);
void am_dev_free(void *dev_p);

/* ---- L1 text substrate used by the wrappers (Utf8.hs) ------------------------------------------
 * Host-side helpers so a binding does not need its own copies. */
/* Utf8.lowerUtf8 (:138-140); returns AM_E_OVERFLOW with *out_len = needed size. */
int am_lower_utf8(const am_lower_table *lower, const am_u8slice *text, uint8_t *out, size_t cap, uint64_t *out_len);
/* Utf8.skipCodePointsBackwards (:256-276); AM_E_BADARG where the reference calls `error`. */
int am_skip_code_points_backwards(const am_u8slice *text, int64_t index, int64_t n, int64_t *out_index);

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
