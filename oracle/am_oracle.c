/*
 * am_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY, see am_oracle.h).
 *
 * A behavioural restatement, in plain C, of
 *   src/Data/Text/AhoCorasick/Automaton.hs  :176-200, :249-380 (build), :442-534 (run)
 *   src/Data/Text/AhoCorasick/Searcher.hs   :156-164 (containsAny), :173-187 (containsAll)
 *   src/Data/Text/AhoCorasick/Replacer.hs   :97-116 (build), :159-274 (run)
 *   src/Data/Text/Utf8.hs                   :131-151 (lowering), :256-276, :337-350 (decode)
 * of channable/alfred-margaret @ dc202ba.  It keeps the reference's packed memory
 * layout (Word64 transitions / Word32 offsets / 128-entry root table / linear edge
 * scan) because it doubles as the timed CPU baseline.  Nothing here is shared with
 * the CUDA product, which uses a different (byte-level, filter + goto) formulation.
 */
#include "am_oracle.h"

#include <stdlib.h>
#include <string.h>
#include <pthread.h>

/* ---- Transition packing, Automaton.hs:75-94, :130-160 ------------------------------- */
#define WILDCARD 0x200000ull
static inline uint64_t new_transition(uint32_t cp, uint32_t st) { return ((uint64_t)st << 32) | cp; }
static inline uint64_t new_wildcard(uint32_t st) { return ((uint64_t)st << 32) | WILDCARD; }
static inline int t_is_wildcard(uint64_t t) { return (t & WILDCARD) == WILDCARD; }
static inline uint32_t t_code_unit(uint64_t t) { return (uint32_t)(t & 0x1fffff); }
static inline uint32_t t_state(uint64_t t) { return (uint32_t)(t >> 32); }

struct amo_machine {
  int64_t num_states;
  int64_t num_transitions;
  uint64_t *transitions;   /* machineTransitions, Automaton.hs:112 */
  uint32_t *offsets;       /* machineOffsets, :116 (numStates + 1 entries) */
  uint64_t root_ascii[128];/* machineRootAsciiTransitions, :119 */
  /* machineValues (:109) is a boxed Vector of cons lists that share tails:
   * values[s] = own(s) ++ values[fail s] (:373-376).  We keep the same sharing:
   * own lists in CSR form plus, per state, the first list cell to visit. */
  int64_t *own_off;        /* num_states + 1 */
  int64_t *own_val;        /* needle indices; later-inserted duplicate first (:263) */
  int32_t *first_out;      /* first state on [s, fail s, fail fail s, ..] with a non-empty own list, or -1 */
  int32_t *next_out;       /* first_out[fail t] for t != root, -1 for root */
  int64_t max_needle_bytes;
  int64_t max_needle_cps;
};

/* ---- UTF-8, Utf8.hs:183-218, :337-350 ------------------------------------------------ */
/* decodeN: no validation; a stray continuation byte decodes as itself (cu0 < 0xc0). */
static inline int decode_at(const uint8_t *d, int64_t idx, int64_t array_end, uint32_t *cp) {
  uint32_t cu0 = d[idx];
  if (cu0 < 0xc0) { *cp = cu0; return 1; }
  /* The reference reads past the slice on truncated input ("returns garbage");
   * we read 0 instead of faulting.  Valid UTF-8 never gets here out of range. */
  uint32_t cu1 = idx + 1 < array_end ? d[idx + 1] : 0;
  if (cu0 < 0xe0) { *cp = ((cu0 & 0x1f) << 6) | (cu1 & 0x3f); return 2; }
  uint32_t cu2 = idx + 2 < array_end ? d[idx + 2] : 0;
  if (cu0 < 0xf0) { *cp = ((cu0 & 0xf) << 12) | ((cu1 & 0x3f) << 6) | (cu2 & 0x3f); return 3; }
  uint32_t cu3 = idx + 3 < array_end ? d[idx + 3] : 0;
  *cp = ((cu0 & 0x7) << 18) | ((cu1 & 0x3f) << 12) | ((cu2 & 0x3f) << 6) | (cu3 & 0x3f);
  return 4;
}

static inline int encode_cp(uint32_t c, uint8_t *o) { /* unicode2utf8, Utf8.hs:154-160 */
  if (c < 0x80) { o[0] = (uint8_t)c; return 1; }
  if (c < 0x800) { o[0] = 0xc0 | (c >> 6); o[1] = 0x80 | (c & 0x3f); return 2; }
  if (c < 0x10000) { o[0] = 0xe0 | (c >> 12); o[1] = 0x80 | ((c >> 6) & 0x3f); o[2] = 0x80 | (c & 0x3f); return 3; }
  o[0] = 0xf0 | (c >> 18); o[1] = 0x80 | ((c >> 12) & 0x3f); o[2] = 0x80 | ((c >> 6) & 0x3f); o[3] = 0x80 | (c & 0x3f);
  return 4;
}

uint32_t amo_lower_code_point(const uint32_t *lower, uint32_t cp) { /* Utf8.hs:131-135, :148-151 */
  if (cp < 128) return (cp >= 'A' && cp <= 'Z') ? cp + 0x20 : cp;
  if (lower && cp < 0x110000) return lower[cp];
  return cp;
}

int64_t amo_lower_utf8(const uint32_t *lower, const uint8_t *in, int64_t len, uint8_t *out, int64_t cap) {
  int64_t i = 0, o = 0;
  while (i < len) {
    uint32_t cp; int n = decode_at(in, i, len, &cp);
    uint8_t buf[4]; int m = encode_cp(amo_lower_code_point(lower, cp), buf);
    if (o + m > cap) return -1;
    memcpy(out + o, buf, (size_t)m);
    o += m; i += n;
  }
  return o;
}

int64_t amo_length_code_points(const uint8_t *in, int64_t len) {
  int64_t n = 0;
  for (int64_t i = 0; i < len; i++) n += (in[i] & 0xc0) != 0x80;
  return n;
}

int64_t amo_skip_code_points_backwards(const uint8_t *data, int64_t off, int64_t len, int64_t index0, int64_t n0) {
  /* Utf8.hs:256-276 */
  if (index0 >= len) return -1;
  int64_t index = index0 + off, n = n0;
  for (;;) {
    if (index >= 0 && (data[index] & 0xc0) == 0x80) { index--; continue; } /* atTrailingByte */
    if (n == 0) return index < 0 ? -1 : index - off;
    if (index < 0) return -1; /* reference would read before the array; we stop */
    index--; n--;
  }
}

/* ---- Construction, Automaton.hs:176-380 ------------------------------------------------ */
typedef struct { uint64_t key; uint32_t val; uint32_t used; } edge_slot;
typedef struct { edge_slot *slots; uint64_t mask; uint64_t count; } edge_map;

static uint64_t hash64(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x; }
static int edge_map_init(edge_map *m, uint64_t cap) { uint64_t c = 16; while (c < cap) c <<= 1; m->slots = calloc(c, sizeof(edge_slot)); m->mask = c - 1; m->count = 0; return m->slots ? 0 : -1; }
static int edge_map_find(const edge_map *m, uint64_t key, uint32_t *val) {
  for (uint64_t i = hash64(key) & m->mask;; i = (i + 1) & m->mask) {
    if (!m->slots[i].used) return 0;
    if (m->slots[i].key == key) { *val = m->slots[i].val; return 1; }
  }
}
static int edge_map_grow(edge_map *m);
static int edge_map_put(edge_map *m, uint64_t key, uint32_t val) {
  if ((m->count + 1) * 2 > m->mask + 1 && edge_map_grow(m)) return -1;
  for (uint64_t i = hash64(key) & m->mask;; i = (i + 1) & m->mask) {
    if (!m->slots[i].used) { m->slots[i].key = key; m->slots[i].val = val; m->slots[i].used = 1; m->count++; return 0; }
  }
}
static int edge_map_grow(edge_map *m) {
  edge_map n; if (edge_map_init(&n, (m->mask + 1) * 2)) return -1;
  for (uint64_t i = 0; i <= m->mask; i++) if (m->slots[i].used) edge_map_put(&n, m->slots[i].key, m->slots[i].val);
  free(m->slots); *m = n; return 0;
}
#define EDGE_KEY(state, cp) (((uint64_t)(state) << 21) | (uint64_t)(cp))

typedef struct { uint32_t cp, child; } child_ent;
static int cmp_child_desc(const void *a, const void *b) {
  uint32_t x = ((const child_ent *)a)->cp, y = ((const child_ent *)b)->cp;
  return x < y ? 1 : x > y ? -1 : 0;
}

int amo_build(const amo_u8slice *needles, size_t n, amo_machine **out) {
  amo_machine *m = calloc(1, sizeof *m);
  if (!m) return -1;
  edge_map em; if (edge_map_init(&em, 1024)) { free(m); return -1; }

  /* buildTransitionMap (:249-292): fold the needles in list order; a new state's id is
   * the running counter.  term_state[i] = state where needle i ends. */
  int64_t num_states = 1;
  uint32_t *term_state = malloc(sizeof(uint32_t) * (n ? n : 1));
  int64_t cap_states = 1024;
  uint32_t *parent = malloc(sizeof(uint32_t) * cap_states), *in_cp = malloc(sizeof(uint32_t) * cap_states);
  parent[0] = 0; in_cp[0] = 0;
  for (size_t i = 0; i < n; i++) {
    const uint8_t *d = needles[i].ptr + needles[i].off; int64_t len = needles[i].len;
    uint32_t state = 0; int64_t idx = 0, ncp = 0;
    while (idx < len) {
      uint32_t cp; int k = decode_at(d, idx, len, &cp); cp &= 0x1fffff;
      uint32_t next;
      if (!edge_map_find(&em, EDGE_KEY(state, cp), &next)) {
        next = (uint32_t)num_states++;
        if (num_states > cap_states) { cap_states *= 2; parent = realloc(parent, sizeof(uint32_t) * cap_states); in_cp = realloc(in_cp, sizeof(uint32_t) * cap_states); }
        parent[next] = state; in_cp[next] = cp;
        edge_map_put(&em, EDGE_KEY(state, cp), next);
      }
      state = next; idx += k; ncp++;
    }
    term_state[i] = state;
    if (len > m->max_needle_bytes) m->max_needle_bytes = len;
    if (ncp > m->max_needle_cps) m->max_needle_cps = ncp;
  }
  m->num_states = num_states;

  /* Children lists per state (IntMap State), kept in DESCENDING code point order, which
   * is the order `makeTransitions` produces by prepending over ascending keys (:190-192). */
  int64_t *child_off = calloc((size_t)num_states + 1, sizeof(int64_t));
  for (int64_t s = 1; s < num_states; s++) child_off[parent[s] + 1]++;
  for (int64_t s = 0; s < num_states; s++) child_off[s + 1] += child_off[s];
  child_ent *children = malloc(sizeof(child_ent) * (size_t)(num_states ? num_states : 1));
  int64_t *fill = malloc(sizeof(int64_t) * (size_t)num_states);
  memcpy(fill, child_off, sizeof(int64_t) * (size_t)num_states);
  for (int64_t s = 1; s < num_states; s++) { child_ent e = { in_cp[s], (uint32_t)s }; children[fill[parent[s]]++] = e; }
  for (int64_t s = 0; s < num_states; s++) qsort(children + child_off[s], (size_t)(child_off[s + 1] - child_off[s]), sizeof(child_ent), cmp_child_desc);

  /* buildFallbackMap (:336-362) over foldBreadthFirst (:309-332).  The reference's queue
   * visits a level in a peculiar (group-FIFO, descending-within-group) order; the result
   * only depends on every shallower state having been processed, which any BFS gives. */
  uint32_t *fallback = calloc((size_t)num_states, sizeof(uint32_t));
  uint32_t *queue = malloc(sizeof(uint32_t) * (size_t)num_states);
  int64_t qh = 0, qt = 0; queue[qt++] = 0;
  while (qh < qt) {
    uint32_t state = queue[qh++];
    for (int64_t c = child_off[state]; c < child_off[state + 1]; c++) {
      uint32_t input = children[c].cp, next = children[c].child;
      /* getFallback fallbacks state input (:342-352) */
      uint32_t fb = 0, st = state;
      while (st != 0) {
        uint32_t f = fallback[st], hit;
        if (edge_map_find(&em, EDGE_KEY(f, input), &hit)) { fb = hit; break; }
        st = f;
      }
      fallback[next] = fb;
      queue[qt++] = next;
    }
  }

  /* Own values: insertWith (++) state [value] (:263) => the later-inserted duplicate first. */
  m->own_off = calloc((size_t)num_states + 1, sizeof(int64_t));
  for (size_t i = 0; i < n; i++) m->own_off[term_state[i] + 1]++;
  for (int64_t s = 0; s < num_states; s++) m->own_off[s + 1] += m->own_off[s];
  m->own_val = malloc(sizeof(int64_t) * (n ? n : 1));
  memcpy(fill, m->own_off, sizeof(int64_t) * (size_t)num_states);
  for (size_t i = n; i-- > 0;) m->own_val[fill[term_state[i]]++] = (int64_t)i;

  /* buildValueMap (:367-380): values[s] = own(s) ++ values[fail s]; BFS order guarantees
   * fail s is done first.  Stored as shared tails (see struct). */
  m->first_out = malloc(sizeof(int32_t) * (size_t)num_states);
  m->next_out = malloc(sizeof(int32_t) * (size_t)num_states);
  for (int64_t q = 0; q < qt; q++) {
    uint32_t s = queue[q];
    int has_own = m->own_off[s + 1] > m->own_off[s];
    int32_t inherited = s == 0 ? -1 : m->first_out[fallback[s]];
    m->next_out[s] = inherited;
    m->first_out[s] = has_own ? (int32_t)s : inherited;
  }

  /* packTransitions (:166-172): per state, children (descending cp) then the wildcard. */
  m->num_transitions = (num_states - 1) + num_states;
  m->transitions = malloc(sizeof(uint64_t) * (size_t)m->num_transitions);
  m->offsets = malloc(sizeof(uint32_t) * ((size_t)num_states + 1));
  int64_t t = 0;
  for (int64_t s = 0; s < num_states; s++) {
    m->offsets[s] = (uint32_t)t;
    for (int64_t c = child_off[s]; c < child_off[s + 1]; c++) m->transitions[t++] = new_transition(children[c].cp, children[c].child);
    m->transitions[t++] = new_wildcard(fallback[s]);
  }
  m->offsets[num_states] = (uint32_t)t;

  /* buildAsciiTransitionLookupTable (:301-306) */
  for (uint32_t i = 0; i < 128; i++) {
    uint32_t st;
    m->root_ascii[i] = edge_map_find(&em, EDGE_KEY(0, i), &st) ? new_transition(i, st) : new_wildcard(0);
  }

  free(em.slots); free(term_state); free(parent); free(in_cp); free(child_off); free(children); free(fill); free(fallback); free(queue);
  *out = m;
  return 0;
}

void amo_free(amo_machine *m) {
  if (!m) return;
  free(m->transitions); free(m->offsets); free(m->own_off); free(m->own_val); free(m->first_out); free(m->next_out); free(m);
}
int64_t amo_num_states(const amo_machine *m) { return m->num_states; }
int64_t amo_num_transitions(const amo_machine *m) { return m->num_transitions; }
const uint64_t *amo_transitions(const amo_machine *m) { return m->transitions; }
const uint32_t *amo_offsets(const amo_machine *m) { return m->offsets; }
const uint64_t *amo_root_ascii(const amo_machine *m) { return m->root_ascii; }

/* ---- runWithCase, Automaton.hs:442-534 -------------------------------------------------- */
/* Specialised on (case, callback) the way GHC specialises the INLINE fold at its call site. */
#define RUN_BODY(LOWER_EXPR, ON_MATCH) {                                                            \
  __label__ collect, next_input;                                                                    \
  const uint8_t *u8data = text.ptr;                                                                 \
  const int64_t initial_offset = text.off, limit = text.off + text.len;                             \
  const uint64_t *transitions = m->transitions; const uint32_t *offsets = m->offsets;               \
  const uint64_t *root_ascii = m->root_ascii;                                                       \
  int64_t offset = initial_offset; uint32_t state = 0;                                              \
  while (offset < limit) { /* consumeInput :468-480 */                                              \
    uint32_t cp; offset += decode_at(u8data, offset, limit, &cp);                                   \
    cp = (LOWER_EXPR);                                                                              \
    uint64_t t;                                                                                     \
    for (;;) { /* followCodePoint :482-486 */                                                       \
      if (state == 0 && cp < 128) { /* lookupRootAsciiTransition :514-520 */                        \
        t = root_ascii[cp];                                                                         \
        if (t_is_wildcard(t)) goto next_input;                                                      \
        goto collect;                                                                               \
      }                                                                                             \
      for (uint32_t i = offsets[state];; i++) { /* lookupTransition :489-510 */                     \
        t = transitions[i];                                                                         \
        if (t_is_wildcard(t)) {                                                                     \
          if (state == 0) goto next_input;                                                          \
          state = t_state(t); break; /* follow the fallback edge, retry the same code point */      \
        }                                                                                           \
        if (t_code_unit(t) == cp) goto collect;                                                     \
      }                                                                                             \
    }                                                                                               \
  collect: /* collectMatches :522-534 */                                                            \
    state = t_state(t);                                                                             \
    for (int32_t o = m->first_out[state]; o >= 0; o = m->next_out[o])                              \
      for (int64_t j = m->own_off[o]; j < m->own_off[o + 1]; j++) { ON_MATCH }                      \
  next_input:;                                                                                      \
  } }

void amo_run_with_case(const amo_machine *m, int cs, const uint32_t *lower, amo_u8slice text, amo_fold_fn f, void *acc) {
  if (cs == AMO_IGNORE_CASE) {
    RUN_BODY(amo_lower_code_point(lower, cp), if (f(acc, offset - initial_offset, m->own_val[j])) return;)
  } else {
    RUN_BODY(cp, if (f(acc, offset - initial_offset, m->own_val[j])) return;)
  }
}

uint64_t amo_count(const amo_machine *m, int cs, const uint32_t *lower, amo_u8slice text) {
  /* countMatches, benchmark/haskell/app/Main.hs:67-76: runText 0 (\n _ -> Step (n + 1)) */
  uint64_t n = 0;
  if (cs == AMO_IGNORE_CASE) { RUN_BODY(amo_lower_code_point(lower, cp), n++;) }
  else { RUN_BODY(cp, n++;) }
  return n;
}

int amo_contains_any(const amo_machine *m, int cs, const uint32_t *lower, amo_u8slice text) {
  /* Searcher.hs:156-164: f _ _ = Done True */
  if (cs == AMO_IGNORE_CASE) { RUN_BODY(amo_lower_code_point(lower, cp), (void)j; return 1;) }
  else { RUN_BODY(cp, (void)j; return 1;) }
  return 0;
}

int64_t amo_find_all(const amo_machine *m, int cs, const uint32_t *lower, amo_u8slice text, amo_match *out, int64_t cap) {
  int64_t n = 0;
#define EMIT if (n < cap) { out[n].pos = offset - initial_offset; out[n].value = m->own_val[j]; } n++;
  if (cs == AMO_IGNORE_CASE) { RUN_BODY(amo_lower_code_point(lower, cp), EMIT) }
  else { RUN_BODY(cp, EMIT) }
#undef EMIT
  return n;
}

int amo_contains_all(const amo_machine *m, int64_t num_needles, int cs, const uint32_t *lower, amo_u8slice text) {
  /* Searcher.hs:173-187: IntSet of outstanding needle ids, Done when it becomes empty;
   * result = IS.null of the final set (so zero needles => True without scanning). */
  if (num_needles == 0) return 1;
  uint8_t *seen = calloc((size_t)num_needles, 1); int64_t remaining = num_needles; int done = 0;
#define SEE { int64_t v = m->own_val[j]; if (v < num_needles && !seen[v]) { seen[v] = 1; if (--remaining == 0) { done = 1; goto finished; } } }
  if (cs == AMO_IGNORE_CASE) { RUN_BODY(amo_lower_code_point(lower, cp), SEE) }
  else { RUN_BODY(cp, SEE) }
#undef SEE
finished:
  free(seen);
  return done;
}

typedef struct { const amo_machine *m; int cs; const uint32_t *lower; amo_u8slice text; int r, threads; int64_t halo; uint64_t count; } shard_job;

static void *shard_main(void *arg) {
  shard_job *j = arg;
  amo_u8slice text = j->text;
  const uint8_t *base = text.ptr + text.off;
  int64_t b = text.len * j->r / j->threads, e = text.len * (j->r + 1) / j->threads;
  /* shard boundaries must sit on code point boundaries */
  while (b > 0 && b < text.len && (base[b] & 0xc0) == 0x80) b++;
  while (e < text.len && (base[e] & 0xc0) == 0x80) e++;
  j->count = 0;
  if (b >= e) return NULL;
  int64_t w = b - j->halo; if (w < 0) w = 0;
  while (w > 0 && (base[w] & 0xc0) == 0x80) w--;
  amo_u8slice shard = { text.ptr, text.off + w, e - w };
  amo_u8slice warm = { text.ptr, text.off + w, b - w };
  j->count = amo_count(j->m, j->cs, j->lower, shard) - amo_count(j->m, j->cs, j->lower, warm);
  return NULL;
}

uint64_t amo_count_parallel(const amo_machine *m, int cs, const uint32_t *lower, amo_u8slice text, int threads) {
  if (threads <= 1 || text.len < (int64_t)threads * 4096) return amo_count(m, cs, lower, text);
  if (threads > 256) threads = 256;
  /* Depth of any state <= max needle length, so a warm-up of that many symbols before the
   * shard start reproduces the sequential state.  Count only matches ending inside the shard
   * (= count over warm-up + shard minus count over the warm-up alone). */
  int64_t halo = cs == AMO_IGNORE_CASE ? 4 * m->max_needle_cps : m->max_needle_bytes;
  shard_job jobs[256]; pthread_t tids[256];
  for (int r = 0; r < threads; r++) {
    shard_job j = { m, cs, lower, text, r, threads, halo, 0 }; jobs[r] = j;
    pthread_create(&tids[r], NULL, shard_main, &jobs[r]);
  }
  uint64_t total = 0;
  for (int r = 0; r < threads; r++) { pthread_join(tids[r], NULL); total += jobs[r].count; }
  return total;
}

typedef struct { const amo_machine *m; int cs; const uint32_t *lower; amo_u8slice text; int r, threads; int64_t halo; amo_match *buf; int64_t n, cap; } fa_job;

static void *fa_main(void *arg) {
  fa_job *j = arg;
  amo_u8slice text = j->text;
  const uint8_t *base = text.ptr + text.off;
  int64_t b = text.len * j->r / j->threads, e = text.len * (j->r + 1) / j->threads;
  while (b > 0 && b < text.len && (base[b] & 0xc0) == 0x80) b++;
  while (e < text.len && (base[e] & 0xc0) == 0x80) e++;
  j->n = 0;
  if (b >= e) return NULL;
  int64_t w = b - j->halo; if (w < 0) w = 0;
  while (w > 0 && (base[w] & 0xc0) == 0x80) w--;
  amo_u8slice shard = { text.ptr, text.off + w, e - w };
  for (;;) {
    int64_t n = amo_find_all(j->m, j->cs, j->lower, shard, j->buf, j->cap);
    if (n <= j->cap) { j->n = n; break; }
    free(j->buf); j->cap = n + n / 8 + 16; j->buf = malloc(sizeof(amo_match) * (size_t)j->cap);
  }
  /* keep the matches ending inside (b, e], rebased to the whole text */
  int64_t k = 0;
  for (int64_t i = 0; i < j->n; i++) {
    int64_t pos = j->buf[i].pos + w;
    if (pos > b) { j->buf[k].pos = pos; j->buf[k].value = j->buf[i].value; k++; }
  }
  j->n = k;
  return NULL;
}

int64_t amo_find_all_parallel(const amo_machine *m, int cs, const uint32_t *lower, amo_u8slice text,
                              amo_match *out, int64_t cap, int threads) {
  if (threads <= 1 || text.len < (int64_t)threads * 4096) return amo_find_all(m, cs, lower, text, out, cap);
  if (threads > 256) threads = 256;
  int64_t halo = cs == AMO_IGNORE_CASE ? 4 * m->max_needle_cps : m->max_needle_bytes;
  fa_job jobs[256]; pthread_t tids[256];
  for (int r = 0; r < threads; r++) {
    int64_t c0 = text.len / threads / 256 + 1024;
    fa_job j = { m, cs, lower, text, r, threads, halo, malloc(sizeof(amo_match) * (size_t)c0), 0, c0 }; jobs[r] = j;
    pthread_create(&tids[r], NULL, fa_main, &jobs[r]);
  }
  int64_t total = 0;
  for (int r = 0; r < threads; r++) {
    pthread_join(tids[r], NULL);
    for (int64_t i = 0; i < jobs[r].n; i++) { if (total < cap) out[total] = jobs[r].buf[i]; total++; }
    free(jobs[r].buf);
  }
  return total;
}

/* ---- Replacer, Replacer.hs ---------------------------------------------------------------- */
struct amo_replacer {
  amo_machine *machine;
  int cs;
  int64_t n;
  int64_t *len_bytes, *len_cps;  /* Payload.needleLengthBytes / needleLengthCodePoints (:59-70) */
  uint8_t **repl; int64_t *repl_len;
};

int amo_replacer_build(const amo_u8slice *needles, const int64_t *len_bytes, const int64_t *len_cps,
                       const amo_u8slice *repls, size_t n, int cs, amo_replacer **out) {
  amo_replacer *r = calloc(1, sizeof *r);
  if (!r) return -1;
  if (amo_build(needles, n, &r->machine)) { free(r); return -1; }
  r->cs = cs; r->n = (int64_t)n;
  r->len_bytes = malloc(sizeof(int64_t) * (n ? n : 1)); r->len_cps = malloc(sizeof(int64_t) * (n ? n : 1));
  r->repl = malloc(sizeof(uint8_t *) * (n ? n : 1)); r->repl_len = malloc(sizeof(int64_t) * (n ? n : 1));
  for (size_t i = 0; i < n; i++) {
    r->len_bytes[i] = len_bytes[i]; r->len_cps[i] = len_cps[i];
    r->repl_len[i] = repls[i].len; r->repl[i] = malloc((size_t)(repls[i].len ? repls[i].len : 1));
    memcpy(r->repl[i], repls[i].ptr + repls[i].off, (size_t)repls[i].len);
  }
  *out = r; return 0;
}

void amo_replacer_free(amo_replacer *r) {
  if (!r) return;
  for (int64_t i = 0; i < r->n; i++) free(r->repl[i]);
  free(r->repl); free(r->repl_len); free(r->len_bytes); free(r->len_cps); amo_free(r->machine); free(r);
}
void amo_buf_free(void *p) { free(p); }

typedef struct { int64_t start, len; } rmatch; /* Replacer.Match (:159); the replacement is the pass's needle's */
typedef struct {
  const amo_replacer *r; const uint8_t *hay; int64_t hay_len;
  int64_t threshold, p_best;
  rmatch *ms; int64_t n, cap; int64_t best_needle; int err;
} rpass;

static int cmp_rmatch(const void *a, const void *b) {
  const rmatch *x = a, *y = b;
  if (x->start != y->start) return x->start < y->start ? -1 : 1;
  if (x->len != y->len) return x->len < y->len ? -1 : 1;
  return 0;
}

/* prependMatch (:252-260) + makeMatch (:264-274) */
static int replacer_fold(void *acc, int64_t pos, int64_t value) {
  rpass *p = acc;
  int64_t p_match = -value; /* needlePriority = -i (:111) */
  if (!(p_match < p->threshold)) return 0;
  if (p_match < p->p_best) return 0;
  if (p_match > p->p_best) { p->p_best = p_match; p->n = 0; p->best_needle = value; }
  rmatch mm;
  if (p->r->cs == AMO_CASE_SENSITIVE) { mm.start = pos - p->r->len_bytes[value]; mm.len = p->r->len_bytes[value]; }
  else {
    int64_t start = amo_skip_code_points_backwards(p->hay, 0, p->hay_len, pos - 1, p->r->len_cps[value] - 1);
    if (start < 0) { p->err = 1; return 1; }
    mm.start = start; mm.len = pos - start;
  }
  if (p->n == p->cap) { p->cap = p->cap ? p->cap * 2 : 64; p->ms = realloc(p->ms, sizeof(rmatch) * (size_t)p->cap); }
  p->ms[p->n++] = mm;
  return 0;
}

int amo_replacer_run(const amo_replacer *r, const uint32_t *lower, amo_u8slice hay_in, int64_t max_len,
                     uint8_t **out, int64_t *out_len, int *exceeded, int64_t *passes) {
  int64_t len = hay_in.len;
  uint8_t *hay = malloc((size_t)(len ? len : 1));
  memcpy(hay, hay_in.ptr + hay_in.off, (size_t)len);
  *exceeded = 0; if (passes) *passes = 0;
  const int64_t min_priority = 1 - r->n; /* :217 */
  int64_t threshold = 1;                 /* :211 */
  rpass p = { r, 0, 0, 0, 0, 0, 0, 0, 0, 0 };
  for (;;) { /* go threshold haystack (:219-242) */
    p.hay = hay; p.hay_len = len; p.threshold = threshold; p.p_best = INT64_MIN; p.n = 0; p.best_needle = -1;
    amo_u8slice cur = { hay, 0, len };
    amo_run_with_case(r->machine, r->cs, lower, cur, replacer_fold, &p);
    if (passes) (*passes)++;
    if (p.err) { free(hay); free(p.ms); return -2; }
    if (p.n == 0) break; /* (_, []) -> Just haystack */
    const uint8_t *repl = r->repl[p.best_needle]; int64_t rl = r->repl_len[p.best_needle];
    /* replacementLength matches haystack (:183-187), on the matches BEFORE removeOverlap (:240) */
    int64_t new_len = len;
    for (int64_t i = 0; i < p.n; i++) new_len += rl - p.ms[i].len;
    if (max_len >= 0 && new_len > max_len) { *exceeded = 1; free(hay); free(p.ms); *out = NULL; *out_len = 0; return 0; }
    qsort(p.ms, (size_t)p.n, sizeof(rmatch), cmp_rmatch); /* sort (:241); the fold prepends, order is irrelevant after sort */
    /* removeOverlap (:191-198) */
    int64_t kept = 0;
    for (int64_t i = 0; i < p.n; i++) {
      if (kept == 0 || p.ms[i].start >= p.ms[kept - 1].start + p.ms[kept - 1].len) p.ms[kept++] = p.ms[i];
    }
    /* replace (:163-180) */
    int64_t final_len = len;
    for (int64_t i = 0; i < kept; i++) final_len += rl - p.ms[i].len;
    uint8_t *nh = malloc((size_t)(final_len > 0 ? final_len : 1));
    int64_t src = 0, dst = 0;
    for (int64_t i = 0; i < kept; i++) {
      memcpy(nh + dst, hay + src, (size_t)(p.ms[i].start - src)); dst += p.ms[i].start - src;
      memcpy(nh + dst, repl, (size_t)rl); dst += rl;
      src = p.ms[i].start + p.ms[i].len;
    }
    memcpy(nh + dst, hay + src, (size_t)(len - src)); dst += len - src;
    free(hay); hay = nh; len = dst;
    if (p.p_best == min_priority) break; /* :241 */
    threshold = p.p_best;               /* go p ... (:242) */
  }
  free(p.ms);
  *out = hay; *out_len = len;
  return 0;
}
