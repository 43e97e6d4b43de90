// am_build.cpp -- host-side construction of the byte-level automaton image.
//
// Semantics follow `build` in src/Data/Text/AhoCorasick/Automaton.hs:176-200 (trie :249-292,
// failure links :336-362, merged outputs :367-380); the LAYOUT is this library's own (see
// am_internal.h).  What must be preserved for bit-exact results:
//   * a needle given twice is reported twice, the later-inserted one first (:263);
//   * values[s] = own(s) ++ values[fail s] (:373-376): at one end position, longer needles first;
//   * the empty needle sits at the root, is inherited by every state and is reported after
//     every successful transition, never while sitting at the root (:499-503, :517-519).
#include <algorithm>
#include <cstring>
#include <numeric>
#include <unordered_map>
#include <unordered_set>

#include "am_internal.h"

namespace am {

static int utf8_len(uint32_t c) { return c < 0x80 ? 1 : c < 0x800 ? 2 : c < 0x10000 ? 3 : 4; }

int build_lower_table(const am_lower_table* in, LowerTable* out) {
  out->stage1.assign(LOWER_STAGE1, 0);
  out->stage2.assign(128, 0);
  out->any_length_change = false;
  if (!in || in->n == 0) return AM_OK;
  if (!in->pairs) return AM_E_BADARG;
  std::unordered_map<uint32_t, uint16_t> block_of;  // cp >> 7 -> block id
  for (size_t i = 0; i < in->n; i++) {
    uint32_t from = in->pairs[i].from_cp, to = in->pairs[i].to_cp;
    if (from >= 0x110000 || to >= 0x110000) return AM_E_BADARG;
    if (from < 128) continue;  // ASCII is lowered by toLowerAscii (Utf8.hs:131-135, :150)
    uint32_t b = from >> LOWER_BLOCK_SHIFT;
    auto it = block_of.find(b);
    uint16_t id;
    if (it == block_of.end()) {
      if (out->stage2.size() / 128 >= 0xFFFF) return AM_E_BADARG;
      id = (uint16_t)(out->stage2.size() / 128);
      out->stage2.resize(out->stage2.size() + 128, 0);
      block_of.emplace(b, id);
      out->stage1[b] = id;
    } else {
      id = it->second;
    }
    out->stage2[(size_t)id * 128 + (from & 127)] = (int32_t)to - (int32_t)from;
    if (utf8_len(from) != utf8_len(to)) out->any_length_change = true;
  }
  return AM_OK;
}

namespace {

struct Builder {
  // temporary trie in insertion order
  std::vector<uint32_t> parent{0};
  std::vector<uint8_t> in_byte{0};
  std::unordered_map<uint64_t, uint32_t> edge;  // state << 8 | byte -> child
  uint32_t add(uint32_t s, uint8_t b) {
    uint64_t k = ((uint64_t)s << 8) | b;
    auto it = edge.find(k);
    if (it != edge.end()) return it->second;
    uint32_t c = (uint32_t)parent.size();
    parent.push_back(s); in_byte.push_back(b);
    edge.emplace(k, c);
    return c;
  }
};

uint32_t next_pow2(uint64_t x) { uint32_t p = 16; while (p < x) p <<= 1; return p; }

}  // namespace

int build_host_automaton(const am_u8slice* needles, size_t n, int cs, const am_lower_table* lower,
                         HostAutomaton* A, std::string* err) {
  if (n > 0 && !needles) { *err = "needles is null"; return AM_E_BADARG; }
  if (cs != AM_CASE_SENSITIVE && cs != AM_IGNORE_CASE) { *err = "unknown case sensitivity"; return AM_E_BADARG; }
  if (n >= (1ull << 31)) { *err = "too many needles"; return AM_E_BADARG; }
  A->case_sensitivity = cs;
  A->num_needles = (uint32_t)n;
  int rc = build_lower_table(cs == AM_IGNORE_CASE ? lower : nullptr, &A->lower);
  if (rc != AM_OK) { *err = "bad lower table"; return rc; }

  // ---- 1. trie over bytes, insertion order ----------------------------------------------------
  // IgnoreCase (runLower) is served by the CaseSensitive kernels over a lower-cased COPY of the text, which keeps
  // byte offsets only if every lowering keeps its UTF-8 length.  The few code points whose lowering changes
  // length (K U+212A -> k, Å U+212B -> å, ẞ -> ß, İ -> i, Ⱥ -> ⱥ, ...) are therefore left UNCHANGED in the copy
  // and each needle is also inserted in the variants where such a pre-image stands for its lowered code point
  // (same needle id).  A needle with a code point that is not its own lower case can never match (the reference
  // lowers every text code point, Automaton.hs:478-480) and is not inserted at all.
  std::unordered_map<uint32_t, std::vector<uint32_t>> preimages;   // lowered cp -> length-changing pre-images
  std::unordered_map<uint32_t, std::vector<uint32_t>> same_pre;    // lowered cp -> pre-images of the same UTF-8 length (filter cells, step 8)
  std::unordered_set<uint32_t> image;
  std::vector<std::vector<uint8_t>> inserted;                      // every byte string the trie holds (needles and their variants)
  A->ic_copy_exact = true;
  if (cs == AM_IGNORE_CASE && lower)
    for (size_t i = 0; i < lower->n; i++) {
      const uint32_t from = lower->pairs[i].from_cp, to = lower->pairs[i].to_cp;
      if (from >= 128 && utf8_len(from) != utf8_len(to)) preimages[to].push_back(from);
      else if (from >= 128 && from != to) same_pre[to].push_back(from);
      image.insert(to);
    }
  // A code point can occur in the lowered stream iff it is its own lower case or the image of another one.
  auto reachable = [&](uint32_t c) { return A->lower.lower(c) == c || image.count(c) != 0; };
  // GHC's toLower is idempotent, so a length-changing code point is never itself an image; a caller-supplied
  // table that breaks this cannot use the copy scheme (the kept original would also stand for itself).
  for (auto& kv : preimages)
    for (uint32_t from : kv.second)
      if (reachable(from)) A->ic_copy_exact = false;
  if (!A->ic_copy_exact) preimages.clear();
  auto decode = [](const uint8_t* d, uint32_t len, uint32_t i, uint32_t* cp) -> uint32_t {   // decodeN, Utf8.hs:344-350
    uint32_t c0 = d[i];
    if (c0 < 0xC0) { *cp = c0; return 1; }
    uint32_t c1 = i + 1 < len ? d[i + 1] : 0;
    if (c0 < 0xE0) { *cp = ((c0 & 0x1F) << 6) | (c1 & 0x3F); return 2; }
    uint32_t c2 = i + 2 < len ? d[i + 2] : 0;
    if (c0 < 0xF0) { *cp = ((c0 & 0xF) << 12) | ((c1 & 0x3F) << 6) | (c2 & 0x3F); return 3; }
    uint32_t c3 = i + 3 < len ? d[i + 3] : 0;
    *cp = ((c0 & 7) << 18) | ((c1 & 0x3F) << 12) | ((c2 & 0x3F) << 6) | (c3 & 0x3F); return 4;
  };
  auto encode = [](uint32_t c, std::vector<uint8_t>* o) {
    if (c < 0x80) o->push_back((uint8_t)c);
    else if (c < 0x800) { o->push_back(0xC0 | (c >> 6)); o->push_back(0x80 | (c & 0x3F)); }
    else if (c < 0x10000) { o->push_back(0xE0 | (c >> 12)); o->push_back(0x80 | ((c >> 6) & 0x3F)); o->push_back(0x80 | (c & 0x3F)); }
    else { o->push_back(0xF0 | (c >> 18)); o->push_back(0x80 | ((c >> 12) & 0x3F)); o->push_back(0x80 | ((c >> 6) & 0x3F)); o->push_back(0x80 | (c & 0x3F)); }
  };
  Builder B;
  B.edge.reserve(n * 8 + 16);
  std::vector<std::pair<uint32_t, uint32_t>> terms;   // (trie state, needle index); a needle may end in several states (variants)
  terms.reserve(n);
  std::vector<uint32_t> len_bytes(n), len_cps(n);
  A->min_len = 0xFFFFFFFFu; A->max_len = 0; A->max_len_cps = 0; A->num_empty = 0;
  // variants per needle and states overall before the automaton gives up on the lowered-copy scheme (a 16-letter needle
  // with 12 letters that have a length-changing pre-image has 4 096 variants; realistic sets stay far below both)
  constexpr uint64_t MAX_VARIANTS = 4096;
  constexpr size_t MAX_VARIANT_STATES = 8u << 20;
  auto insert = [&](const uint8_t* d, uint32_t len, uint32_t id) {
    uint32_t s = 0;
    for (uint32_t k = 0; k < len; k++) s = B.add(s, d[k]);
    terms.emplace_back(s, id);
    if (len > 0) { A->min_len = std::min(A->min_len, len); A->max_len = std::max(A->max_len, len); inserted.emplace_back(d, d + len); }
  };
  for (size_t i = 0; i < n; i++) {
    if (needles[i].len < 0 || needles[i].off < 0 || (needles[i].len > 0 && !needles[i].ptr)) { *err = "bad needle slice"; return AM_E_BADARG; }
    if ((uint64_t)needles[i].len >= (1ull << 24)) { *err = "needle longer than 16 MiB"; return AM_E_BADARG; }
    const uint8_t* d = needles[i].ptr + needles[i].off;
    const uint32_t len = (uint32_t)needles[i].len;
    uint32_t cps = 0;
    for (uint32_t k = 0; k < len; k++) cps += (d[k] & 0xC0) != 0x80;
    len_bytes[i] = len; len_cps[i] = cps;
    if (len == 0) A->num_empty++;
    A->max_len_cps = std::max(A->max_len_cps, cps);
    if (cs != AM_IGNORE_CASE) { insert(d, len, (uint32_t)i); continue; }
    // IgnoreCase: dead needles, variants
    std::vector<uint32_t> ncp;
    bool dead = false;
    for (uint32_t k = 0; k < len;) { uint32_t c; k += decode(d, len, k, &c); ncp.push_back(c); if (!reachable(c)) dead = true; }
    if (dead) continue;                                   // can never match: every text code point is lowered first
    insert(d, len, (uint32_t)i);
    uint64_t combos = 1;
    for (uint32_t c : ncp) { auto it = preimages.find(c); if (it != preimages.end()) combos *= 1 + it->second.size(); if (combos > MAX_VARIANTS) break; }
    if (combos == 1) continue;
    if (combos > MAX_VARIANTS || B.parent.size() > MAX_VARIANT_STATES) { A->ic_copy_exact = false; continue; }   // too many variants: this automaton uses the sentinel + fallback scheme
    std::vector<uint32_t> choice(ncp.size(), 0);
    for (uint64_t v = 1; v < combos; v++) {                // odometer over the pre-image choices (0 = the lowered code point itself)
      for (size_t p = 0; p < ncp.size(); p++) {
        auto it = preimages.find(ncp[p]);
        const uint32_t radix = it == preimages.end() ? 1 : 1 + (uint32_t)it->second.size();
        if (++choice[p] < radix) break;
        choice[p] = 0;
      }
      std::vector<uint8_t> bytes;
      for (size_t p = 0; p < ncp.size(); p++) encode(choice[p] == 0 ? ncp[p] : preimages[ncp[p]][choice[p] - 1], &bytes);
      insert(bytes.data(), (uint32_t)bytes.size(), (uint32_t)i);
    }
  }
  if (A->min_len == 0xFFFFFFFFu) A->min_len = 0;
  const uint32_t S = (uint32_t)B.parent.size();
  if (S >= ID_MASK) { *err = "too many states"; return AM_E_BADARG; }
  A->num_states = S;

  // ---- 2. BFS renumbering (shallow states get small ids; fail(s) < s) -----------------------------
  std::vector<uint32_t> tmp_child_off(S + 1, 0);
  for (uint32_t s = 1; s < S; s++) tmp_child_off[B.parent[s] + 1]++;
  for (uint32_t s = 0; s < S; s++) tmp_child_off[s + 1] += tmp_child_off[s];
  std::vector<uint32_t> tmp_children(S ? S - 1 : 0);
  {
    std::vector<uint32_t> fill(tmp_child_off.begin(), tmp_child_off.end() - 1);
    for (uint32_t s = 1; s < S; s++) tmp_children[fill[B.parent[s]]++] = s;
    for (uint32_t s = 0; s < S; s++)
      std::sort(tmp_children.begin() + tmp_child_off[s], tmp_children.begin() + tmp_child_off[s + 1],
                [&](uint32_t a, uint32_t b) { return B.in_byte[a] < B.in_byte[b]; });
  }
  std::vector<uint32_t> new_id(S), order(S);
  {
    uint32_t head = 0, tail = 0;
    order[tail++] = 0;
    while (head < tail) {
      uint32_t s = order[head];
      new_id[s] = head++;
      for (uint32_t c = tmp_child_off[s]; c < tmp_child_off[s + 1]; c++) order[tail++] = tmp_children[c];
    }
  }
  A->parent.resize(S); A->in_byte.resize(S); A->depth.resize(S); A->boundary.resize(S);
  A->child_off.assign(S + 1, 0); A->child_state.resize(S ? S - 1 : 0); A->child_byte.resize(S ? S - 1 : 0);
  std::vector<uint8_t> rem(S, 0);  // continuation bytes still expected after this state's prefix
  for (uint32_t k = 0; k < S; k++) {
    uint32_t old = order[k];
    A->parent[k] = new_id[B.parent[old]];
    A->in_byte[k] = B.in_byte[old];
    if (k == 0) { A->depth[0] = 0; rem[0] = 0; A->boundary[0] = 1; }
    else {
      uint32_t p = A->parent[k]; uint8_t b = A->in_byte[k];
      A->depth[k] = A->depth[p] + 1;
      if (rem[p] > 0) rem[k] = rem[p] - 1;
      else rem[k] = b < 0xC0 ? 0 : b < 0xE0 ? 1 : b < 0xF0 ? 2 : 3;  // decodeN's length rule, Utf8.hs:344-350
      A->boundary[k] = rem[k] == 0;
    }
    A->child_off[k + 1] = A->child_off[k] + (tmp_child_off[old + 1] - tmp_child_off[old]);
  }
  for (uint32_t k = 0; k < S; k++) {
    uint32_t old = order[k], o = A->child_off[k];
    for (uint32_t c = tmp_child_off[old]; c < tmp_child_off[old + 1]; c++, o++) {
      A->child_state[o] = new_id[tmp_children[c]];
      A->child_byte[o] = B.in_byte[tmp_children[c]];
    }
  }
  auto goto_child = [&](uint32_t s, uint8_t b) -> uint32_t {
    const uint8_t* lo = A->child_byte.data() + A->child_off[s];
    const uint8_t* hi = A->child_byte.data() + A->child_off[s + 1];
    const uint8_t* it = std::lower_bound(lo, hi, b);
    if (it != hi && *it == b) return A->child_state[it - A->child_byte.data()];
    return NONE;
  };

  // ---- 3. failure links (standard AC; equals buildFallbackMap :336-362) ----------------------------
  A->fail.assign(S, 0);
  for (uint32_t s = 1; s < S; s++) {  // BFS order: parent and all shallower states are done
    uint32_t p = A->parent[s]; uint8_t b = A->in_byte[s];
    uint32_t f = 0;
    if (p != 0) {
      uint32_t st = p;
      while (st != 0) {
        uint32_t g = A->fail[st];
        uint32_t hit = goto_child(g, b);
        if (hit != NONE) { f = hit; break; }
        st = g;
      }
    }
    A->fail[s] = f;
  }

  // ---- 4. needle ranks and output chains ---------------------------------------------------------------
  // Rank needles by (byte length descending, index descending).  All matches ending at one
  // position are suffix-nested and distinct, so ascending rank is the reference's order there.
  A->id_of_rank.resize(n);
  std::iota(A->id_of_rank.begin(), A->id_of_rank.end(), 0u);
  std::sort(A->id_of_rank.begin(), A->id_of_rank.end(), [&](uint32_t a, uint32_t b) {
    if (len_bytes[a] != len_bytes[b]) return len_bytes[a] > len_bytes[b];
    return a > b;
  });
  A->rank_of_id.resize(n); A->len_of_rank.resize(n);
  for (uint32_t r = 0; r < n; r++) { A->rank_of_id[A->id_of_rank[r]] = r; A->len_of_rank[r] = len_bytes[A->id_of_rank[r]]; }
  A->rank_bits = 1; while ((1ull << A->rank_bits) < n) A->rank_bits++;

  A->own_off.assign(S + 1, 0);
  for (auto& tm : terms) A->own_off[new_id[tm.first] + 1]++;
  for (uint32_t s = 0; s < S; s++) A->own_off[s + 1] += A->own_off[s];
  {
    std::vector<uint32_t> fill(A->own_off.begin(), A->own_off.end() - 1);
    // ascending rank within a state => index descending (later duplicate first)
    std::vector<std::pair<uint32_t, uint32_t>> by_rank;   // (rank, state)
    by_rank.reserve(terms.size());
    for (auto& tm : terms) by_rank.emplace_back(A->rank_of_id[tm.second], new_id[tm.first]);
    std::sort(by_rank.begin(), by_rank.end());
    A->own_rank.resize(terms.size());
    for (auto& rs : by_rank) A->own_rank[fill[rs.second]++] = rs.first;
  }
  // chain(s) = own(s) ++ chain(fail s), but only code-point-boundary states report, and the root
  // itself never reports (its own list -- the empty needles -- is only inherited).
  A->first_out.assign(S, NONE); A->next_out.assign(S, NONE); A->chain_count.assign(S, 0);
  const bool root_has_own = A->own_off[1] > A->own_off[0];
  for (uint32_t s = 1; s < S; s++) {
    if (!A->boundary[s]) continue;  // mid code point: nothing is reported here
    uint32_t f = A->fail[s];
    uint32_t inherited, inherited_count;
    if (f == 0) { inherited = root_has_own ? 0u : NONE; inherited_count = A->own_off[1] - A->own_off[0]; }
    else { inherited = A->first_out[f]; inherited_count = A->chain_count[f]; }
    bool has_own = A->own_off[s + 1] > A->own_off[s];
    A->next_out[s] = inherited;
    A->first_out[s] = has_own ? s : inherited;
    A->chain_count[s] = (A->own_off[s + 1] - A->own_off[s]) + inherited_count;
  }
  auto tagged = [&](uint32_t s) -> uint32_t {
    return s | (A->first_out[s] != NONE ? OUT_FLAG : 0u) | (A->own_off[s + 1] > A->own_off[s] ? OWN_FLAG : 0u);
  };

  // ---- 5. halo ---------------------------------------------------------------------------------------------
  // CaseSensitive: a match ending after `begin` starts at most max_len - 1 bytes before it.
  // IgnoreCase: the walk restarts on a code point boundary at least max_len_cps code points
  // (<= 4 bytes each) before the first code point it reports, +3 for snapping forward.
  if (cs == AM_CASE_SENSITIVE) A->halo_bytes = A->max_len > 0 ? A->max_len - 1 : 0;
  else A->halo_bytes = A->max_len_cps > 0 ? 4ull * A->max_len_cps + 4 : 0;

  // ---- 6. dense rows for the shallowest states (generic fallback of ac_step) ---------------------------------
  {
    uint32_t cap = 256;
    A->dense_states = std::min(S, cap);
    A->dense.assign((size_t)A->dense_states * 256, 0);
    for (uint32_t s = 0; s < A->dense_states; s++) {
      uint32_t* row = A->dense.data() + (size_t)s * 256;
      if (s == 0) { for (int b = 0; b < 256; b++) row[b] = 0; }
      else std::memcpy(row, A->dense.data() + (size_t)A->fail[s] * 256, 256 * sizeof(uint32_t));  // fail(s) < s
      for (uint32_t c = A->child_off[s]; c < A->child_off[s + 1]; c++) row[A->child_byte[c]] = tagged(A->child_state[c]);
    }
  }
  // ---- 6b. class-compressed failure-resolved automaton (walk kernel) ------------------------------------------
  {
    bool used[256] = {false};
    for (uint8_t cb : A->child_byte) used[cb] = true;
    uint32_t nc = 1;
    for (int b = 0; b < 256; b++) A->cls[b] = used[b] ? (uint8_t)(nc++ & 0xFF) : 0;
    if (nc > 256) {  // all 256 byte values occur: class 0 must still mean "no edge anywhere"; use the identity map
      for (int b = 0; b < 256; b++) A->cls[b] = (uint8_t)b;
      nc = 256;
    }
    A->num_classes = nc;
    A->cdfa_shift = 1; while ((1u << A->cdfa_shift) < nc) A->cdfa_shift++;
    const uint64_t stride = 1ull << A->cdfa_shift;
    const uint64_t budget_words = (256ull << 20) / 4;                       // 256 MiB of rows at most
    A->cdfa_states = (uint32_t)std::min<uint64_t>(S, budget_words / stride);
    A->cdfa.assign((size_t)A->cdfa_states * stride, 0);
    for (uint32_t s = 0; s < A->cdfa_states; s++) {
      uint32_t* row = A->cdfa.data() + (size_t)s * stride;
      if (s != 0) std::memcpy(row, A->cdfa.data() + (size_t)A->fail[s] * stride, stride * sizeof(uint32_t));  // fail(s) < s
      for (uint32_t c = A->child_off[s]; c < A->child_off[s + 1]; c++) row[A->cls[A->child_byte[c]]] = tagged(A->child_state[c]);
    }
  }

  // ---- 7. hashed goto edges ----------------------------------------------------------------------------------------
  {
    uint32_t cap = next_pow2((uint64_t)(S - 1) * 2 + 16);
    A->edges.assign(cap, EdgeSlot{NONE, NONE, NONE, 0});
    A->edge_mask = cap - 1;
    for (uint32_t s = 0; s < S; s++)
      for (uint32_t c = A->child_off[s]; c < A->child_off[s + 1]; c++) {
        uint32_t b = A->child_byte[c];
        uint64_t key = ((uint64_t)s << 8) | b;
        uint32_t i = edge_hash(s, b) & A->edge_mask;
        while (A->edges[i].child != NONE) i = (i + 1) & A->edge_mask;
        A->edges[i] = EdgeSlot{(uint32_t)key, (uint32_t)(key >> 32), tagged(A->child_state[c]), 0};
      }
  }

  // ---- 8. q-gram filter + jump table (filter kernel; not applicable with empty needles) ------------------------
  // q = the shortest needle's length up to 4; needle sets too large for the exact second level whose shortest needle
  // has >= 6 (>= 8) bytes take 6- (8-)grams: the bitmaps then answer for more of every needle, and a set of 10^5
  // needles over a small alphabet -- every 4-gram of which is some needle's prefix -- still filters.
  //
  // The two filter levels are built from "gram instances": (first q bytes, byte q or CLOSED) of every string the trie
  // holds.  IgnoreCase: of every CASE VARIANT of those strings' first code points -- each code point replaced by any of its
  // same-length pre-images under the caller's toLower table (é <- É, я <- Я, ǳ <- ǲ Ǳ) -- folded (| 0x20: ASCII letters).
  // So the kernel can probe the ORIGINAL text folded with one OR per word and never lowers a code point in its hot loop.
  A->q = 0; A->filter_keys = 0; A->ic_fold_ok = true;
  if (A->num_empty == 0 && A->min_len > 0) {
    const bool ic = cs == AM_IGNORE_CASE;
    constexpr uint32_t CLOSED = 0x100u;
    constexpr size_t MAX_CASE_VARIANTS = 256;
    struct Instance { uint64_t gram; uint32_t next; };
    auto gram_mask = [](uint32_t q) -> uint64_t { return q >= 8 ? ~0ull : ((1ull << (8 * q)) - 1ull); };
    auto instances_of = [&](uint32_t q, std::vector<Instance>* out) {
      std::unordered_map<uint64_t, std::vector<uint16_t>> seen;   // gram -> `next` values already emitted
      out->clear();
      auto emit = [&](const uint8_t* b, size_t len) {
        uint64_t g = 0;
        for (uint32_t k = 0; k < q; k++) g |= (uint64_t)b[k] << (8 * k);
        uint32_t nx = len > q ? b[q] : CLOSED;
        if (ic) { g = (g | 0x2020202020202020ull) & gram_mask(q); if (nx != CLOSED) nx |= 0x20u; }
        auto& v = seen[g];
        if (std::find(v.begin(), v.end(), (uint16_t)nx) == v.end()) { v.push_back((uint16_t)nx); out->push_back(Instance{g, nx}); }
      };
      std::vector<uint8_t> buf;
      for (const auto& str : inserted) {
        if (str.size() < q) continue;                          // (cannot happen: q <= min_len)
        if (!ic) { emit(str.data(), str.size()); continue; }
        // code points that reach into bytes [0, q]: their case variants
        std::vector<std::vector<uint32_t>> opts;
        std::vector<uint32_t> cps;
        uint32_t covered = 0;
        size_t combos = 1;
        while (covered < std::min<size_t>(str.size(), q + 1)) {
          uint32_t c; covered += decode(str.data(), (uint32_t)str.size(), covered, &c);
          cps.push_back(c);
          std::vector<uint32_t> o{c};
          auto it = same_pre.find(c);
          if (c >= 128 && it != same_pre.end()) for (uint32_t f : it->second) if (utf8_len(f) == utf8_len(c)) o.push_back(f);
          combos *= o.size();
          opts.push_back(std::move(o));
        }
        if (combos > MAX_CASE_VARIANTS) { A->ic_fold_ok = false; combos = 1; for (auto& o : opts) o.resize(1); }   // this automaton scans a lowered copy instead
        std::vector<size_t> choice(opts.size(), 0);
        for (size_t v = 0; v < combos; v++) {
          buf.clear();
          for (size_t p = 0; p < opts.size(); p++) encode(opts[p][choice[p]], &buf);
          buf.insert(buf.end(), str.begin() + std::min<size_t>(covered, str.size()), str.end());   // (only the length matters beyond byte q)
          emit(buf.data(), buf.size());
          for (size_t p = 0; p < opts.size(); p++) { if (++choice[p] < opts[p].size()) break; choice[p] = 0; }
        }
      }
    };
    auto distinct_grams = [](const std::vector<Instance>& v) { std::unordered_set<uint64_t> d; for (auto& x : v) d.insert(x.gram); return d.size(); };
    std::vector<Instance> inst;
    A->q = std::min<uint32_t>(4, A->min_len);
    instances_of(A->q, &inst);
    A->t2_exact = distinct_grams(inst) <= T2_MAX_EXACT_KEYS;  // the exact second level holds the distinct (folded) q-grams: does it fit?
    if (!A->t2_exact && A->min_len >= 6) {
      A->q = A->min_len >= 8 ? 8 : 6;
      instances_of(A->q, &inst);
    }
    const uint32_t q = A->q;
    // ---- level 1: bitmap cells of every distinct (folded) gram --------------------------------------------------------
    A->filter.assign(FILTER_WORDS, 0);
    const int copies = filter_copies(q, A->t2_exact != 0);
    {
      std::unordered_set<uint64_t> done;
      for (auto& x : inst) {
        if (!done.insert(x.gram).second) continue;
        if (q > 4) {             // long q-gram: one cell, two bits of its word (stride-1 probe)
          uint32_t row, by, bt;
          long_cell(x.gram, q, &row, &by, &bt);
          A->filter[row] |= (1u << by) | (1u << bt);
        } else if (filter_is_s2(q)) {   // stride-2 probe: one cell per parity of the start position
          uint32_t ra, ba, rb, bb;
          filter_cells_s2((uint32_t)x.gram, filter_rowbits(copies), &ra, &ba, &rb, &bb);
          for (int c = 0; c < copies; c++) {
            A->filter[(size_t)ra * copies + c] |= 1u << ba;
            A->filter[(size_t)rb * copies + c] |= 1u << bb;
          }
        } else {
          uint32_t row, bit;
          filter_cell((uint32_t)x.gram, &row, &bit);
          for (int c = 0; c < copies; c++) A->filter[(size_t)row * copies + c] |= 1u << bit;
        }
      }
    }
    // ---- jump table: exact q-gram -> trie state (or the tail of its single needle path) -----------------------------------
    std::vector<std::pair<uint64_t, uint32_t>> keys;          // (q-gram, state at depth q)
    for (uint32_t s = 0; s < S; s++) {
      if (A->depth[s] != q) { if (A->depth[s] > q) break; continue; }
      uint64_t g = 0; uint32_t t = s;
      for (uint32_t k = q; k-- > 0;) { g |= (uint64_t)A->in_byte[t] << (8 * k); t = A->parent[t]; }
      keys.emplace_back(g, s);
    }
    A->filter_keys = (uint32_t)keys.size();
    uint32_t cap = next_pow2((uint64_t)keys.size() * 2 + 16);
    A->jump.assign(cap, JumpSlot{0, NONE, 0, 0});
    A->tails.clear();
    A->jump_mask = cap - 1;
    const uint32_t tail_from = std::min<uint32_t>(q, 4);       // the tail of a slot starts at this needle byte
    for (auto& kv : keys) {
      const uint32_t key_lo = (uint32_t)kv.first, key_hi = (uint32_t)(kv.first >> 32);
      uint32_t i = jump_hash(key_lo, key_hi) & A->jump_mask;
      while (A->jump[i].state != NONE) i = (i + 1) & A->jump_mask;
      JumpSlot slot{key_lo, tagged(kv.second), key_hi, 0};    // not simple: tail_off holds the q-gram's bytes 4 .. q-1
      {  // simple sub-trie?  follow the only child while the state holds no needle end
        uint32_t t = kv.second;
        std::vector<uint8_t> tail;
        for (uint32_t k = tail_from; k < q; k++) tail.push_back((uint8_t)(kv.first >> (8 * k)));   // the q-gram's own bytes beyond the key
        bool simple = true;
        for (;;) {
          const uint32_t nch = A->child_off[t + 1] - A->child_off[t];
          const uint32_t nown = A->own_off[t + 1] - A->own_off[t];
          if (nch == 0) { simple = nown > 0; break; }               // leaf
          if (nch > 1 || nown > 0) { simple = false; break; }
          tail.push_back(A->child_byte[A->child_off[t]]);
          t = A->child_state[A->child_off[t]];
        }
        if (FK_TAIL && simple && tail.size() <= JUMP_TAIL_MASK) {
          const uint32_t nown = A->own_off[t + 1] - A->own_off[t];
          slot.tail_off = (uint32_t)A->tails.size();
          slot.meta = JUMP_SIMPLE | (uint32_t)tail.size();
          if (nown == 1) { slot.meta |= JUMP_SINGLE; slot.state = A->own_rank[A->own_off[t]]; }
          else slot.state = t;
          A->tails.insert(A->tails.end(), tail.begin(), tail.end());
          while (A->tails.size() % 4) A->tails.push_back(0);
        }
      }
      A->jump[i] = slot;
    }
    // ---- level 2 -----------------------------------------------------------------------------------------------------------
    A->gbits.clear(); A->gbits_log2 = 0;
    if (A->t2_exact) {
      // shared memory, 32 KiB: the exact (folded) keys; aux = the (folded) byte that must follow the q-gram when every
      // string through it continues with that byte
      std::unordered_map<uint32_t, uint32_t> aux_of;
      for (auto& x : inst) {
        const uint32_t g = (uint32_t)x.gram, aux = x.next == CLOSED ? T2_AUX_ANY : x.next;
        auto it = aux_of.find(g);
        if (it == aux_of.end()) aux_of.emplace(g, aux);
        else if (it->second != aux) it->second = T2_AUX_ANY;
      }
      uint32_t empty = 0xFFFFFFFFu;  // any value that is not a key (a text q-gram equal to it is rejected later)
      while (aux_of.count(empty)) empty--;
      A->t2_empty_key = empty;
      // buckets of {key0, aux0, key1, aux1}; almost always one probe per lookup: a bucket that would take a
      // third key is flagged "overflow" and only then does a lookup continue in the next bucket
      A->filter2.assign(T2_WORDS, empty);
      for (uint32_t bkt = 0; bkt < (1u << T2_LOG2_BUCKETS); bkt++) { A->filter2[4 * (size_t)bkt + 1] = T2_AUX_ANY; A->filter2[4 * (size_t)bkt + 3] = T2_AUX_ANY; }
      std::vector<std::pair<uint32_t, uint32_t>> sorted_aux(aux_of.begin(), aux_of.end());
      std::sort(sorted_aux.begin(), sorted_aux.end());        // deterministic table layout
      for (auto& ga : sorted_aux) {
        const uint32_t g = ga.first, aux = ga.second;
        uint32_t hb = t2_bucket(g);
        for (;;) {
          uint32_t* slot = A->filter2.data() + 4 * (size_t)hb;
          if (slot[0] == empty) { slot[0] = g; slot[1] = aux; break; }
          if (slot[2] == empty) { slot[2] = g; slot[3] = (slot[3] & T2_AUX_OVERFLOW) | aux; break; }
          slot[3] |= T2_AUX_OVERFLOW;   // full: lookups that miss here go on to the next bucket
          hb = (hb + 1) & ((1u << T2_LOG2_BUCKETS) - 1);
        }
      }
    } else if (q > 4) {
      // global memory: a bitmap over the whole q-gram, ~256 bits per key, 2^20 .. 2^27 bits (128 KiB .. 16 MiB: L2-resident)
      A->filter2.assign(T2_WORDS, 0);
      const size_t nkeys = distinct_grams(inst);
      A->gbits_log2 = 20;
      while (A->gbits_log2 < 27 && (1ull << A->gbits_log2) < (uint64_t)nkeys * 256) A->gbits_log2++;
      A->gbits.assign((size_t)1 << (A->gbits_log2 - 5), 0);
      for (auto& x : inst) {
        const uint32_t b = gq_hash((uint32_t)x.gram, (uint32_t)(x.gram >> 32)) >> (32 - A->gbits_log2);
        A->gbits[b >> 5] |= 1u << (b & 31);
      }
    } else {
      A->filter2.assign(T2_WORDS, 0);
      auto set_bit = [&](uint32_t word0, uint32_t bit) { A->filter2[word0 + (bit >> 5)] |= 1u << (bit & 31); };
      for (auto& x : inst) {
        const uint32_t g = (uint32_t)x.gram;
        if (q < 4) { set_bit(0, filter2_bit(g)); continue; }
        if (x.next == CLOSED) { set_bit(T2A_WORD0, t2a_bit(g)); continue; }   // q = 4: closed 4-grams + the 5-grams of the strings that go on
        set_bit(T2B_WORD0, t2b_bit(g, x.next));
        set_bit(T2C_WORD0, t2c_bit(g, x.next));
      }
      if (q == 4 && A->ic_fold_ok) {
        // third level (global memory): the (folded) prefix of min(length, 8) bytes of every string the trie holds -- IgnoreCase: of
        // every case variant of the code points that reach into it -- keyed with its length (gp_hash)
        std::unordered_set<uint64_t> keys[9];
        bool ok = true;
        std::vector<uint8_t> buf;
        for (const auto& str : inserted) {
          const uint32_t L = (uint32_t)std::min<size_t>(str.size(), 8);
          auto emit = [&](const uint8_t* b) {
            uint64_t g = 0;
            for (uint32_t k = 0; k < L; k++) g |= (uint64_t)b[k] << (8 * k);
            if (ic) g = (g | 0x2020202020202020ull) & gram_mask(L);
            keys[L].insert(g);
          };
          if (!ic) { emit(str.data()); continue; }
          std::vector<std::vector<uint32_t>> opts;
          uint32_t covered = 0;
          size_t combos = 1;
          while (covered < L) {
            uint32_t c; covered += decode(str.data(), (uint32_t)str.size(), covered, &c);
            std::vector<uint32_t> o{c};
            auto it = same_pre.find(c);
            if (c >= 128 && it != same_pre.end()) for (uint32_t f : it->second) if (utf8_len(f) == utf8_len(c)) o.push_back(f);
            combos *= o.size();
            opts.push_back(std::move(o));
            if (combos > MAX_CASE_VARIANTS) break;
          }
          if (combos > MAX_CASE_VARIANTS) { ok = false; break; }   // (this image does without the third level)
          std::vector<size_t> choice(opts.size(), 0);
          for (size_t v = 0; v < combos; v++) {
            buf.clear();
            for (size_t p = 0; p < opts.size(); p++) encode(opts[p][choice[p]], &buf);
            buf.resize(std::max<size_t>(buf.size(), 8), 0);
            emit(buf.data());
            for (size_t p = 0; p < opts.size(); p++) { if (++choice[p] < opts[p].size()) break; choice[p] = 0; }
          }
        }
        if (ok) {
          size_t nkeys = 0;
          for (auto& k : keys) nkeys += k.size();
          A->gbits_log2 = 20;
          while (A->gbits_log2 < 27 && (1ull << A->gbits_log2) < (uint64_t)nkeys * 256) A->gbits_log2++;
          A->gbits.assign((size_t)1 << (A->gbits_log2 - 5), 0);
          for (uint32_t L = 4; L <= 8; L++)
            for (uint64_t g : keys[L]) {
              const uint32_t b = gp_hash((uint32_t)g, (uint32_t)(g >> 32), L) >> (32 - A->gbits_log2);
              A->gbits[b >> 5] |= 1u << (b & 31);
            }
        }
      }
    }
  }
  return AM_OK;
}

}  // namespace am
