/*
 * am_oracle.h -- CPU oracle for the alfred-margaret Aho-Corasick hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's
 * algorithm (channable/alfred-margaret @ dc202ba, v2.1.1.1) used as the parity
 * checker and as the timed CPU baseline.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (libam_b200.so) never links, loads or calls anything in this directory.
 *
 * Parity pinning: GHC is not available in this image, so the reference itself
 * cannot be run.  The oracle is pinned against every known-answer vector the
 * reference's own tests/README hold for this path; they
 * are transcribed in tests/golden/reference_vectors.json and checked by
 * tests/test_oracle_golden.py.
 *
 * Citations are relative to /root/reference/.
 */
#ifndef AM_ORACLE_H
#define AM_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Mirrors `U8Slice` of benchmark/rust-ffi/app/Main.hs:32-46 and
 * benchmark/rust-ffi/libacbench/src/lib.rs:5-11: (array pointer, off, len),
 * i.e. an unpacked `Text u8data off len`. */
typedef struct {
  const uint8_t *ptr;
  int64_t off;
  int64_t len;
} amo_u8slice;

/* `Match { matchPos, matchValue }`, src/Data/Text/AhoCorasick/Automaton.hs:98-105.
 * `value` is the index of the needle in the list given to amo_build (the host maps
 * it to the caller's payload v). */
typedef struct {
  int64_t pos;
  int64_t value;
} amo_match;

typedef struct amo_machine amo_machine;

/* `Next a = Done | Step`, Automaton.hs:398.  Fold callback: return 0 to Step, 1 to be Done. */
typedef int (*amo_fold_fn)(void *acc, int64_t pos, int64_t value);

enum { AMO_CASE_SENSITIVE = 0, AMO_IGNORE_CASE = 1 }; /* CaseSensitivity.hs:14-16 */

/* `build`, Automaton.hs:176-200. */
int amo_build(const amo_u8slice *needles, size_t n, amo_machine **out);
void amo_free(amo_machine *m);
int64_t amo_num_states(const amo_machine *m);
int64_t amo_num_transitions(const amo_machine *m);
/* Raw packed arrays, for layout tests (Automaton.hs:108-123). */
const uint64_t *amo_transitions(const amo_machine *m);
const uint32_t *amo_offsets(const amo_machine *m);
const uint64_t *amo_root_ascii(const amo_machine *m);

/* Lowering table = GHC `Char.toLower` as data: `lower` is NULL (identity above
 * ASCII) or a dense array of 0x110000 code points.  Utf8.hs:145-151. */
uint32_t amo_lower_code_point(const uint32_t *lower, uint32_t cp);
/* `lowerUtf8`, Utf8.hs:138-140.  Returns bytes written (cap must be >= 4*len/1... see .c). */
int64_t amo_lower_utf8(const uint32_t *lower, const uint8_t *in, int64_t len, uint8_t *out, int64_t cap);
/* `Text.length` (code points). */
int64_t amo_length_code_points(const uint8_t *in, int64_t len);
/* `skipCodePointsBackwards`, Utf8.hs:256-276.  Returns -1 where the reference calls `error`. */
int64_t amo_skip_code_points_backwards(const uint8_t *data, int64_t off, int64_t len, int64_t index, int64_t n);

/* `runWithCase`, Automaton.hs:442-534 (runText :539, runLower :551). */
void amo_run_with_case(const amo_machine *m, int case_sensitivity, const uint32_t *lower,
                       amo_u8slice text, amo_fold_fn f, void *acc);

/* Convenience folds used by tests and the CPU baseline. */
uint64_t amo_count(const amo_machine *m, int cs, const uint32_t *lower, amo_u8slice text);
int amo_contains_any(const amo_machine *m, int cs, const uint32_t *lower, amo_u8slice text);
/* Writes up to cap matches in callback order; returns the total number of matches. */
int64_t amo_find_all(const amo_machine *m, int cs, const uint32_t *lower, amo_u8slice text,
                     amo_match *out, int64_t cap);
/* Searcher.containsAll, Searcher.hs:173-187 (the "next" row, 8f rank 1). */
int amo_contains_all(const amo_machine *m, int64_t num_needles, int cs, const uint32_t *lower, amo_u8slice text);

/* Multi-core count for the `--impl reference` arm: the same sequential fold run on
 * `threads` overlapping shards (halo = max needle bytes, IgnoreCase: 4x code points).
 * NOT something the reference does; reported with its core count. */
uint64_t amo_count_parallel(const amo_machine *m, int cs, const uint32_t *lower, amo_u8slice text,
                            int threads);

/* Same sharding for the all-matches fold (matches of shard r precede those of shard r + 1, so the
 * concatenation is in callback order). */
int64_t amo_find_all_parallel(const amo_machine *m, int cs, const uint32_t *lower, amo_u8slice text,
                              amo_match *out, int64_t cap, int threads);

/* Replacer, src/Data/Text/AhoCorasick/Replacer.hs:97-116 (build), :200-274 (run). */
typedef struct amo_replacer amo_replacer;
/* needles[i] must already be lowered for IgnoreCase (Replacer.hs:105-107, use amo_lower_utf8);
 * len_bytes / len_cps are those of the ORIGINAL needle (:111-113). */
int amo_replacer_build(const amo_u8slice *needles, const int64_t *len_bytes, const int64_t *len_cps,
                       const amo_u8slice *repls, size_t n, int cs, amo_replacer **out);
void amo_replacer_free(amo_replacer *r);
/* `runWithLimit`: returns 0 and sets *exceeded=1 for `Nothing`.  *out is malloc'd, free with amo_buf_free.
 * max_len < 0 means maxBound (`run`).  *passes = number of scans performed. */
int amo_replacer_run(const amo_replacer *r, const uint32_t *lower, amo_u8slice hay, int64_t max_len,
                     uint8_t **out, int64_t *out_len, int *exceeded, int64_t *passes);
void amo_buf_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
