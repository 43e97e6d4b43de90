/* am_shim.c -- by-pointer wrappers for the entry points that take `am_u8slice` by value.
 *
 * GHC's FFI cannot pass a struct by value; the reference's own FFI precedent passes `*const U8Slice`
 * (benchmark/rust-ffi/app/Main.hs:28-52).  The Haskell shim (../Data/Text/AhoCorasick/FFI.hs) imports these.
 * cabal: `c-sources: cbits/am_shim.c`, `include-dirs: include`, `extra-libraries: am_b200`. */
#include "am_b200.h"

int am_shim_contains_any(const am_automaton *a, const am_u8slice *h, int *o) { return am_contains_any(a, *h, o); }
int am_shim_count_matches(const am_automaton *a, const am_u8slice *h, uint64_t *o) { return am_count_matches(a, *h, o); }
int am_shim_find_all(const am_automaton *a, const am_u8slice *h, am_match *o, size_t c, uint64_t *n) { return am_find_all(a, *h, o, c, n); }
int am_shim_contains_all(const am_automaton *a, const am_u8slice *h, int *o) { return am_contains_all(a, *h, o); }
int am_shim_replacer_run(const am_replacer *r, const am_u8slice *h, uint64_t m, uint8_t **o, uint64_t *n, int *x) { return am_replacer_run(r, *h, m, o, n, x); }
