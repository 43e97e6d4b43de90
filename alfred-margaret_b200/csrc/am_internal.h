// am_internal.h -- internal structures of libam_b200 (host build + device image).
//
// The reference keys its automaton on code points and scans edge lists linearly
// (src/Data/Text/AhoCorasick/Automaton.hs:75-123).  This implementation is NOT a port of
// that layout.  It keys on BYTES (valid UTF-8 needles can only occur at code point boundaries
// of valid UTF-8 text) and keeps four device structures:
//
//   filter   : q-gram membership bitmap (q = min(4, shortest needle) bytes), one private copy
//              per shared-memory bank (32 x 4 KiB) so a warp's 32 probes never conflict;
//   jump     : open-addressed table  q-gram -> trie state at depth q;
//   edges    : open-addressed table  (state, byte) -> child  (the flattened goto function);
//   dense    : failure-resolved 256-wide rows for the shallowest states + fail links, used by
//              the per-segment walk kernel (general path: empty needles, huge match density).
//
// Output order is reproduced by ranking needles (longer first, later-inserted duplicate
// first: Automaton.hs:263, :373-376) and sorting (end_pos, rank) keys on the device.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "am_b200.h"

namespace am {

#if defined(__CUDACC__)
#define AM_HD_DECL __host__ __device__ __forceinline__
#else
#define AM_HD_DECL inline
#endif

constexpr uint32_t NONE = 0xFFFFFFFFu;
constexpr uint32_t OUT_FLAG = 0x80000000u;   // stored state id: its output chain (own ++ inherited) is non-empty
constexpr uint32_t OWN_FLAG = 0x40000000u;   // stored state id: some needle ends exactly at this state
constexpr uint32_t ID_MASK = 0x3FFFFFFFu;

constexpr int FILTER_ROWS = 1024;            // rows of 32 words (one word per bank) => 128 KiB image
constexpr int FILTER_WORDS = FILTER_ROWS * 32;
constexpr uint32_t HASH_MUL = 0x9E3779B1u;   // filter hash multiplier
constexpr uint32_t HASH_MUL2 = 0x85EBCA6Bu;  // second-level filter / table hash multiplier
constexpr int FILTER2_LOG2_BITS = 18;        // second-level table: 32 KiB of shared memory, either a 2^18-bit bitmap ...
constexpr int T2_WORDS = (1 << FILTER2_LOG2_BITS) / 32;   // ... or 2048 buckets x 2 x (exact q-gram key, aux)
constexpr int T2_LOG2_BUCKETS = 11;
constexpr uint32_t T2_MAX_EXACT_KEYS = 2048; // load factor <= 0.5
constexpr uint32_t T2_AUX_ANY = 0x100u;      // aux: no constraint on the byte after the q-gram; else that byte's value
constexpr uint32_t T2_AUX_OVERFLOW = 0x200u; // on slot 1's aux: the bucket overflowed at build time => the lookup continues in the next bucket

constexpr uint32_t LOWER_BLOCK_SHIFT = 7;    // two-stage lower-case table: 128 code points per block
constexpr uint32_t LOWER_STAGE1 = 0x110000 >> LOWER_BLOCK_SHIFT;

struct EdgeSlot { uint32_t key_lo, key_hi, child, pad; };   // key = state << 8 | byte ; child carries OUT_FLAG
// Jump table slot: q-gram -> trie state at depth q (state == NONE => empty; state carries the flags).  When the
// sub-trie below that state is a single path that ends in a leaf and holds no other needle end ("simple": true for
// nearly every q-gram of a random needle set), the slot also carries the path: `meta` = JUMP_SIMPLE | tail length,
// `tail_off` = offset of the tail bytes in HostAutomaton::tails (4-byte aligned), and `state` = the LEAF's state --
// or, with JUMP_SINGLE (one needle ends there), directly that needle's rank.  The survivor check is then one
// byte-wise comparison whose loads are all independent, instead of one dependent hashed edge lookup per byte.
// For q > 4 the key is the q-gram's first four bytes and the table is hashed by all q: the bytes 4 .. q - 1 are checked
// against the head of the tail (the tail of a slot starts at needle byte min(q, 4)) or, for a slot that is not simple,
// against `tail_off`, which then holds those bytes; a slot that fails this check is another q-gram's: probing goes on.
struct JumpSlot { uint32_t key, state, tail_off, meta; };
constexpr uint32_t JUMP_SIMPLE = 0x80000000u, JUMP_SINGLE = 0x40000000u, JUMP_TAIL_MASK = 0xFFFFu;

inline uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x;
}

// Two-stage table for Char.toLower above ASCII (Utf8.hs:148-151): delta[stage1[cp >> 7] * 128 + (cp & 127)].
struct LowerTable {
  std::vector<uint16_t> stage1;   // LOWER_STAGE1 entries; block 0 is the all-zero (identity) block
  std::vector<int32_t> stage2;    // blocks * 128 deltas
  bool any_length_change = false; // some pair changes the UTF-8 byte length
  uint32_t lower(uint32_t cp) const {
    if (cp < 128) return (cp >= 'A' && cp <= 'Z') ? cp + 0x20 : cp;
    if (cp >= 0x110000 || stage1.empty()) return cp;
    return (uint32_t)((int32_t)cp + stage2[(size_t)stage1[cp >> LOWER_BLOCK_SHIFT] * 128 + (cp & 127)]);
  }
};
int build_lower_table(const am_lower_table* in, LowerTable* out);

// Host image of the byte-level automaton; states are numbered in BFS order (root = 0).
struct HostAutomaton {
  int case_sensitivity = AM_CASE_SENSITIVE;
  uint32_t num_needles = 0;
  uint32_t num_states = 1;
  uint32_t min_len = 0, max_len = 0;   // bytes, over non-empty needles
  uint32_t max_len_cps = 0;
  uint32_t num_empty = 0;              // empty needles (reported after every successful transition, A.4)
  bool ic_copy_exact = true;           // IgnoreCase: needle variants cover every length-changing pre-image (no fallback needed)
  uint32_t q = 0;                      // filter q-gram length (1..4, 6 or 8), 0 => filter kernel not applicable
  uint32_t rank_bits = 1;
  uint64_t halo_bytes = 0;             // bytes a shard needs before its report range

  std::vector<uint32_t> fail, depth, parent;
  std::vector<uint8_t> in_byte, boundary;       // boundary: the state's byte prefix ends on a code point boundary
  std::vector<uint32_t> child_off, child_state; // CSR, children sorted by byte
  std::vector<uint8_t> child_byte;
  std::vector<uint32_t> own_off, own_rank;      // CSR: ranks of the needles ending exactly at the state, ascending
  std::vector<uint32_t> first_out, next_out;    // output chain (NONE = end); see build
  std::vector<uint32_t> chain_count;            // total matches reported when arriving at the state
  std::vector<uint32_t> rank_of_id, id_of_rank, len_of_rank;

  uint32_t dense_states = 0;                    // rows in `dense`
  std::vector<uint32_t> dense;                  // dense_states * 256, failure-resolved, OUT_FLAG tagged
  // Class-compressed failure-resolved automaton for the walk kernel: byte -> class (0 = "no needle
  // contains this byte"), then next = cdfa[state << cdfa_shift | class] for the first cdfa_states states.
  uint8_t cls[256] = {0};
  uint32_t num_classes = 1, cdfa_shift = 1, cdfa_states = 0;
  std::vector<uint32_t> cdfa;
  std::vector<EdgeSlot> edges; uint32_t edge_mask = 0;
  std::vector<JumpSlot> jump; uint32_t jump_mask = 0;
  std::vector<uint8_t> tails;                   // tail bytes of the simple jump slots
  std::vector<uint32_t> filter;                 // FILTER_WORDS, bank-replicated
  std::vector<uint32_t> filter2;                // T2_WORDS: bitmap, or exact key buckets when t2_exact
  std::vector<uint32_t> gbits; uint32_t gbits_log2 = 0;   // q > 4: second level in global memory (2^gbits_log2 bits); q = 4 with bitmap T2: third level (prefix keys, gp_hash)
  bool ic_fold_ok = true;                       // IgnoreCase: the case variants of every needle's first code points fit the cells (one-pass scan)
  uint32_t t2_exact = 0, t2_empty_key = 0xFFFFFFFFu;
  uint32_t filter_keys = 0;                     // distinct q-grams
  LowerTable lower;
};

int build_host_automaton(const am_u8slice* needles, size_t n, int cs, const am_lower_table* lower,
                         HostAutomaton* out, std::string* err);

inline uint32_t qgram_mask(uint32_t q) { return q >= 4 ? 0xFFFFFFFFu : ((1u << (8 * q)) - 1u); }
// Build-time variants of the filter kernel (A/B-tested on the GPU):
//   FK_S2         stride-2 probe for q = 4: ONE bitmap word answers "does a needle start at p" and "does a needle
//                 start at p + 1" for an even p.  The row is hashed from the three bytes the two q-grams share
//                 (text[p+1..p+4)); the bit is picked by the low 5 bits of the byte that is private to each
//                 (text[p] resp. text[p+4]).  Half the hashes / address computations / shared-memory loads per
//                 text byte; every needle inserts two cells.
//   FK_COPIES_S2, copies of the q-gram bitmap side by side in the 32 shared-memory banks, for the stride-2 and the
//   FK_COPIES_S1  stride-1 (q < 4) probe.  Lane l probes copy l mod COPIES, and a copy spans 32 / COPIES banks
//                 (consecutive rows in consecutive banks), so the lanes that share a copy spread over its banks by
//                 the low row bits: 32 copies (32 Ki bits each) are conflict-free, 16 (64 Ki bits) cost ~1.5
//                 wavefronts per probe, 8 -> ~2.1, 4 -> ~2.6, 2 (512 Ki bits) -> ~3.1, 1 -> ~3.5.  Fewer copies =
//                 fewer false positives.  The stride-2 probe issues half the loads, so it affords 2 copies
//                 (measured best: 16/8/4/2/1 copies -> 1.93/1.72/1.63/1.57/1.50* ms per 4 GiB; *with later changes
//                 1 copy was 1 % slower than 2); the stride-1 probe stays at 16.
//   FK_WB         stride-1 probe: take the bit index from the low 5 bits of the multiplicative hash and form the
//                 address with IMAD (FMA pipe) -- one ALU-pipe instruction less per probe.
//   FK_TAIL       survivors whose q-gram leads into a single needle path are checked by one tail comparison
//                 (JumpSlot) instead of a walk through the hashed goto table.
#ifndef FK_S2
#define FK_S2 1
#endif
#ifndef FK_TAIL
#define FK_TAIL 1
#endif
#ifndef FK_COPIES_S2
#define FK_COPIES_S2 2
#endif
#ifndef FK_COPIES_S2_BIG     // needle sets beyond the exact second level (> T2_MAX_EXACT_KEYS q-grams): ONE copy of 1 Mi bits
#define FK_COPIES_S2_BIG 1   // (a few more bank conflicts per probe, half the false candidates -- they dominate there)
#endif
#ifndef FK_COPIES_S1
#define FK_COPIES_S1 16
#endif
#ifndef FK_WB
#define FK_WB 1
#endif
constexpr int filter_rowbits(int copies) { return copies == 32 ? 10 : copies == 16 ? 11 : copies == 8 ? 12 : copies == 4 ? 13 : copies == 2 ? 14 : 15; }
constexpr int FILTER_ROWBITS_S1 = filter_rowbits(FK_COPIES_S1);
static_assert((1 << filter_rowbits(FK_COPIES_S2)) * FK_COPIES_S2 == FILTER_WORDS && (1 << filter_rowbits(FK_COPIES_S2_BIG)) * FK_COPIES_S2_BIG == FILTER_WORDS &&
              (1 << FILTER_ROWBITS_S1) * FK_COPIES_S1 == FILTER_WORDS, "filter geometry");
inline bool filter_is_s2(uint32_t q) { return FK_S2 && q == 4; }
// Copies of the bitmap for an automaton: `exact` = its q-grams fit the exact second-level table (t2_exact).
constexpr int filter_copies_s2(bool exact) { return exact ? FK_COPIES_S2 : FK_COPIES_S2_BIG; }
inline int filter_copies(uint32_t q, bool exact) { return q > 4 ? 1 : filter_is_s2(q) ? filter_copies_s2(exact) : FK_COPIES_S1; }
// Filter cell of a (masked) q-gram: row and bit 0..31.  Must match the device code.
inline void filter_cell(uint32_t g, uint32_t* row, uint32_t* bit) {
  const uint32_t x = g * HASH_MUL;
  *row = x >> (32 - FILTER_ROWBITS_S1);
  const uint32_t s = FK_WB ? (x & 31u) : ((x >> 15) & 31u);
  *bit = 31u - s;   // the kernel rotates left by s and tests bit 31
}
// Stride-2 cells of a 4-gram g = n0 | n1 << 8 | n2 << 16 | n3 << 24 (FK_S2).  The kernel hashes the 4-gram X that
// starts at p + 1 with HASH_MUL << 8, which drops X's top byte: the row depends on text[p+1..p+4) only.
//   cell A (needle starts at the even position p):      row of (n1, n2, n3), bit chosen by n0
//   cell B (needle starts at the odd position p + 1):   row of (n0, n1, n2), bit chosen by n3
constexpr uint32_t HASH_MUL_S2 = HASH_MUL << 8;
// Stride-2 cells of a needle whose first 4 bytes are g (byte i at bits 8 i):
//   cell A (needle starts at the even position p):      row of bytes 1 .. 3, bit chosen by byte 0
//   cell B (needle starts at the odd position p + 1):   row of bytes 0 .. 2, bit chosen by byte 3
inline void filter_cells_s2(uint32_t g, int rowbits, uint32_t* row_a, uint32_t* bit_a, uint32_t* row_b, uint32_t* bit_b) {
  *row_a = ((g >> 8) * HASH_MUL_S2) >> (32 - rowbits);
  *bit_a = 31u - (g & 31u);            // the kernel rotates left by text[p] and tests bit 31
  *row_b = (g * HASH_MUL_S2) >> (32 - rowbits);
  *bit_b = 31u - ((g >> 24) & 31u);    // ... by text[p + 4]
}
// Longer q-grams (q = 6, 8: needle sets too large for the exact second level whose shortest needle has >= 6 / >= 8 bytes).
// A set of 10^5 needles would fill a sixth of the 1 Mi-bit bitmap with one-bit stride-2 cells (two per needle): a fifth of
// all text positions would pass.  These sets take ONE cell per needle, probed at every position (stride 1), that sets TWO
// bits of its word (a blocked Bloom filter, k = 2: ~3 % pass at 10^5 needles):
//   t = X(p) * HASH_MUL                      X(i): the 4-gram at i
//   y = t + X(p + q - 4) * HASH_MUL_B        all q bytes
//   word = row (y >> 17);  bits 31 - (y & 31) and 31 - (t & 31): the kernel rotates the word left by y and by t
// Must match fk_probe16_long.
constexpr uint32_t HASH_MUL_B = 0xCC9E2D51u;
constexpr int FILTER_ROWBITS_LONG = 15;       // one copy of the bitmap
AM_HD_DECL void long_cell(uint64_t g, uint32_t q, uint32_t* row, uint32_t* bit_y, uint32_t* bit_t) {
  const uint32_t t = (uint32_t)g * HASH_MUL;
  const uint32_t y = t + (uint32_t)(g >> (8 * (q - 4))) * HASH_MUL_B;
  *row = y >> (32 - FILTER_ROWBITS_LONG);
  *bit_y = 31u - (y & 31u);
  *bit_t = 31u - (t & 31u);
}
// IgnoreCase automata: the filter works on FOLDED bytes, every byte | 0x20 -- an ASCII letter and its upper case fold to
// the same byte, every other byte only loses one bit.  `Char.toLower` above ASCII is not byte-local (Я D0 AF -> я D1 8F), so
// the NEEDLE side enumerates instead: the bitmap cells and the second level hold the folded q-grams of every case variant
// of a needle's first code points (the same-length pre-images of each code point under the caller's toLower table;
// am_build.cpp step 8).  The kernel folds the text with one OR per word before it probes and never sees `Char.toLower`;
// folding can add candidates, never lose one.  The survivors are verified on exactly lowered code points.
constexpr uint32_t FOLD_MASK = 0x20202020u;
AM_HD_DECL uint32_t fold20(uint32_t w) { return w | FOLD_MASK; }
inline uint32_t filter2_bit(uint32_t g) { return (g * HASH_MUL2) >> (32 - FILTER2_LOG2_BITS); }
// Second level for q = 4 needle sets too large for the exact table (T2_MAX_EXACT_KEYS): three bitmaps in the 32 KiB.
//   T2A (64 Ki bits):  4-grams that END a needle (a needle of exactly four bytes);
//   T2B (128 Ki bits) and T2C (64 Ki bits, independent hash): the 5-grams of the needles that go on -- a candidate whose
//   4-gram is a needle prefix survives only if its fifth byte continues some needle as well.
// A candidate survives if T2A[g4] or (T2B[g5] and T2C[g5]).  Must match fk_phase_a.
constexpr int T2A_LOG2 = 16, T2B_LOG2 = 17, T2C_LOG2 = 16;
constexpr uint32_t T2A_WORD0 = 0, T2B_WORD0 = (1u << T2A_LOG2) / 32, T2C_WORD0 = T2B_WORD0 + (1u << T2B_LOG2) / 32;
static_assert(T2C_WORD0 + (1u << T2C_LOG2) / 32 == (uint32_t)T2_WORDS, "T2 bitmap geometry");
constexpr uint32_t HASH_MUL3 = 0xC2B2AE35u;
// Second level for q > 4 (large needle sets): a bitmap over the whole q-gram (lo = bytes 0..3, hi = bytes 4..q-1) in
// GLOBAL memory -- 2^20 .. 2^27 bits, ~256 bits per key, resident in the 126 MB L2 -- because 32 KiB of shared memory
// cannot tell 10^5 keys apart.  The kernel looks the level-1 candidates up warp-cooperatively, 32 independent loads per
// round (fk_global_rounds).  Must match the kernel.
AM_HD_DECL uint32_t gq_hash(uint32_t lo, uint32_t hi) {
  uint32_t h = (lo * HASH_MUL2) ^ (hi * HASH_MUL3);
  h ^= h >> 15;
  return h * HASH_MUL;
}
// q = 4 images whose second level is the three shared-memory bitmaps also carry a THIRD level in global memory: a bitmap over the
// needles' (folded) prefixes of min(length, 8) bytes, keyed with that length -- closed needles of 4 .. 7 bytes and the 8-byte
// prefixes of the longer ones.  verify_kernel tests a survivor's eight carried text bytes against it (five independent loads) before
// it decodes or lowers anything: at 10^4 needles five survivors in six are false positives of the 32 KiB bitmaps.
AM_HD_DECL uint32_t gp_hash(uint32_t lo, uint32_t hi, uint32_t len) { return gq_hash(lo, hi + len * 0x85EBCA6Bu); }
AM_HD_DECL uint32_t t2a_bit(uint32_t g4) { return (g4 * HASH_MUL2) >> (32 - T2A_LOG2); }
AM_HD_DECL uint32_t t2b_bit(uint32_t g4, uint32_t b4) { return ((g4 * HASH_MUL2) ^ (b4 * HASH_MUL)) >> (32 - T2B_LOG2); }
AM_HD_DECL uint32_t t2c_bit(uint32_t g4, uint32_t b4) { return ((g4 ^ (b4 * 0x01000193u)) * HASH_MUL3) >> (32 - T2C_LOG2); }
inline uint32_t t2_bucket(uint32_t g) { return (g * HASH_MUL2) >> (32 - T2_LOG2_BUCKETS); }
// Table hashes (host + device); callers mask with the table's power-of-two mask.
#if defined(__CUDACC__)
#define AM_HD __host__ __device__ __forceinline__
#else
#define AM_HD inline
#endif
AM_HD uint32_t jump_hash(uint32_t g, uint32_t g_hi = 0) { uint32_t h = g * HASH_MUL2 + g_hi * HASH_MUL3; return h ^ (h >> 15); }
AM_HD uint32_t edge_hash(uint32_t state, uint32_t byte) {
  uint32_t h = state * HASH_MUL + byte * 0x01000193u; return h ^ (h >> 15);
}

}  // namespace am
