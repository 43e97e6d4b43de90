/* A plain C host driving EVERY hot-path symbol of include/am_b200.h by pointer -- what GHC's `foreign import ccall`
 * does (benchmark/rust-ffi/app/Main.hs:28-29 in the reference).  Built with gcc by tests/test_gpu_abi_c.py on the GPU
 * box; prints one "name=value" line per check, the test compares them with the reference's known answers
 * (README.md:44-100, tests/Data/Text/AhoCorasickSpec.hs:99-118). */
#include <cuda_runtime_api.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "am_b200.h"

#define CHECK(call)                                                                 \
  do {                                                                              \
    int rc_ = (call);                                                               \
    if (rc_ != AM_OK) { printf("FAILED %s -> %d: %s\n", #call, rc_, am_last_error()); return 1; } \
  } while (0)

static am_u8slice S(const char *s) { am_u8slice x; x.ptr = (const uint8_t *)s; x.off = 0; x.len = (int64_t)strlen(s); return x; }

int main(void) {
  am_u8slice needles[3] = {S("tshirt"), S("shirts"), S("shorts")};
  am_lower_pair pairs[2] = {{0xC9, 0xE9}, {0x212A, 0x6B}};          /* É -> é, KELVIN SIGN -> k (changes its UTF-8 length) */
  am_lower_table lower = {pairs, 2};
  am_options opts; memset(&opts, 0, sizeof opts); opts.device = 0;
  am_automaton *a = 0;
  CHECK(am_automaton_build(needles, 3, &lower, &opts, &a));
  CHECK(am_automaton_prepare(a, AM_IGNORE_CASE));
  uint64_t states = 0, maxlen = 0, halo = 0; int kind = 0;
  CHECK(am_automaton_info(a, AM_CASE_SENSITIVE, &states, &maxlen, &halo, &kind));
  printf("info=%llu,%llu,%llu,%d\n", (unsigned long long)states, (unsigned long long)maxlen, (unsigned long long)halo, kind);

  /* runText: all matches of "sweatshirts and shirtshirts" (README.md:94-100, here in ascending order) */
  am_u8slice hay = S("sweatshirts and shirtshirts");
  am_match out[16]; uint64_t n = 0;
  CHECK(am_find_all(a, AM_CASE_SENSITIVE, &hay, out, 16, &n));
  printf("find_all=");
  for (uint64_t i = 0; i < n; i++) printf("%llu:%u%s", (unsigned long long)out[i].end_pos, out[i].needle_id, i + 1 < n ? "," : "");
  printf("\n");
  if (am_find_all(a, AM_CASE_SENSITIVE, &hay, out, 2, &n) != AM_E_OVERFLOW) { printf("FAILED overflow protocol\n"); return 1; }
  printf("overflow_needs=%llu\n", (unsigned long long)n);

  /* the same handle in both case modes (Searcher.hs:142-145, AhoCorasickSpec.hs:169-179) */
  am_u8slice shout = S("Short TSHIRTS");
  int b0 = -1, b1 = -1; uint64_t c0 = 0, c1 = 0;
  CHECK(am_contains_any(a, AM_CASE_SENSITIVE, &shout, &b0));
  CHECK(am_contains_any(a, AM_IGNORE_CASE, &shout, &b1));
  CHECK(am_count_matches(a, AM_CASE_SENSITIVE, &shout, &c0));
  CHECK(am_count_matches(a, AM_IGNORE_CASE, &shout, &c1));
  printf("contains_any=%d,%d count=%llu,%llu\n", b0, b1, (unsigned long long)c0, (unsigned long long)c1);
  int all0 = -1, all1 = -1;
  am_u8slice three = S("shorts, tshirts");
  CHECK(am_contains_all(a, AM_CASE_SENSITIVE, &three, &all0));
  CHECK(am_contains_all(a, AM_CASE_SENSITIVE, &shout, &all1));
  printf("contains_all=%d,%d\n", all0, all1);

  /* device-resident and sharded calls on a text that sits at offset 3 of a device buffer, with a slice offset on the host side */
  const char *raw = "xxxsweatshirts and shirtshirts";
  void *d = 0; am_match *d_out = 0;
  if (cudaMalloc(&d, 64) != cudaSuccess || cudaMalloc((void **)&d_out, 16 * sizeof(am_match)) != cudaSuccess) return 2;
  cudaMemcpy(d, raw, strlen(raw), cudaMemcpyHostToDevice);
  am_dev_text t; t.dev_text = (const char *)d + 3; t.text_len = 27; t.report_begin = 0; t.pos_base = 1000;
  CHECK(am_count_matches_dev(a, AM_CASE_SENSITIVE, &t, 0, &c0));
  CHECK(am_contains_any_dev(a, AM_CASE_SENSITIVE, &t, 0, &b0));
  CHECK(am_find_all_dev(a, AM_CASE_SENSITIVE, &t, 0, d_out, 16, &n));
  cudaMemcpy(out, d_out, n * sizeof(am_match), cudaMemcpyDeviceToHost);
  printf("dev=%llu,%d,%llu first=%llu:%u\n", (unsigned long long)c0, b0, (unsigned long long)n, (unsigned long long)out[0].end_pos, out[0].needle_id);
  am_u8slice off_slice; off_slice.ptr = (const uint8_t *)raw; off_slice.off = 3; off_slice.len = 27;
  CHECK(am_count_matches(a, AM_CASE_SENSITIVE, &off_slice, &c1));
  uint64_t w = 0, b = 0, e = 0;
  CHECK(am_shard_plan(27, halo, 2, 1, &w, &b, &e));
  t.pos_base = 0;
  am_dev_text sh = t; sh.dev_text = (const char *)t.dev_text + w; sh.text_len = e - w; sh.report_begin = b - w; sh.pos_base = w;
  am_comm *comm = 0;
  CHECK(am_comm_init(0, 1, 0, 0, &comm));
  am_shard_result r0, r1;
  CHECK(am_find_all_sharded(a, AM_CASE_SENSITIVE, comm, &sh, 0, d_out, 16, &r1));
  am_dev_text sh0 = t; sh0.text_len = b;
  CHECK(am_count_sharded(a, AM_CASE_SENSITIVE, comm, &sh0, 0, &r0));
  CHECK(am_contains_any_sharded(a, AM_CASE_SENSITIVE, comm, &sh, 0, &b0));
  uint64_t v = 41; CHECK(am_comm_allreduce_u64(comm, &v, 0, 0));
  CHECK(am_shard_halo_exchange(comm, d, 5, 22, 0));
  printf("slice_off=%llu shards=%llu+%llu any=%d total=%llu allreduce=%llu\n", (unsigned long long)c1, (unsigned long long)r0.n_local,
         (unsigned long long)r1.n_local, b0, (unsigned long long)r1.total, (unsigned long long)v);
  am_comm_free(comm);

  /* Replacer: build lowers the needles of an IgnoreCase replacer (Replacer.hs:105-107); the run takes the case flag */
  am_u8slice rn[2] = {S("foo"), S("bar")}, rr[2] = {S("BAR"), S("BAZ")};
  am_replacer *rep = 0;
  CHECK(am_replacer_build(rn, rr, 2, AM_IGNORE_CASE, &lower, &opts, &rep));
  am_u8slice foo = S("Foo foo");
  uint8_t *res = 0; uint64_t res_len = 0; int exceeded = -1;
  CHECK(am_replacer_run(rep, AM_IGNORE_CASE, &foo, UINT64_MAX, &res, &res_len, &exceeded));          /* AhoCorasickSpec.hs:117-118 */
  printf("replace_ic=%.*s\n", (int)res_len, (const char *)res);
  printf("passes=%llu\n", (unsigned long long)am_replacer_last_passes());
  am_free(res);
  CHECK(am_replacer_run(rep, AM_CASE_SENSITIVE, &foo, UINT64_MAX, &res, &res_len, &exceeded));       /* setCaseSensitivity: same stored needles */
  printf("replace_cs=%.*s\n", (int)res_len, (const char *)res);
  am_free(res);
  CHECK(am_replacer_run(rep, AM_IGNORE_CASE, &foo, 3, &res, &res_len, &exceeded));                   /* runWithLimit -> Nothing */
  printf("exceeded=%d\n", exceeded);
  void *d_res = 0;
  cudaMemcpy(d, "Foo foo", 7, cudaMemcpyHostToDevice);
  CHECK(am_replacer_run_dev(rep, AM_IGNORE_CASE, d, 7, UINT64_MAX, 0, &d_res, &res_len, &exceeded));
  char back[32]; memset(back, 0, sizeof back);
  cudaMemcpy(back, d_res, res_len, cudaMemcpyDeviceToHost);
  printf("replace_dev=%s\n", back);
  am_dev_free(d_res);
  am_replacer_free(rep);
  /* the stored form (compose / FromJSON): needle "éclair" with the payload lengths of "ÉCLAIR" */
  am_u8slice sn[1] = {S("\xc3\xa9" "clair")}, sr[1] = {S("bolt")};
  uint32_t lb[1] = {7}, lc[1] = {6};
  CHECK(am_replacer_build_stored(sn, lb, lc, sr, 1, AM_IGNORE_CASE, &lower, &opts, &rep));
  am_u8slice un = S("un \xc3\x89" "clair");
  CHECK(am_replacer_run(rep, AM_IGNORE_CASE, &un, UINT64_MAX, &res, &res_len, &exceeded));
  printf("replace_stored=%.*s\n", (int)res_len, (const char *)res);
  am_free(res);
  am_replacer_free(rep);

  /* L1 text substrate */
  uint8_t low[16]; uint64_t low_len = 0;
  am_u8slice up = S("\xc3\x89" "A\xe2\x84\xaa");                                                    /* É A K(U+212A) */
  CHECK(am_lower_utf8(&lower, &up, low, sizeof low, &low_len));
  int64_t idx = -1;
  am_u8slice poo = S("\xf0\x9f\x92\xa9\xf0\x9f\x92\xa9");
  CHECK(am_skip_code_points_backwards(&poo, 7, 1, &idx));                                            /* Utf8Spec.hs:115-154 */
  printf("lower_len=%llu skip=%lld\n", (unsigned long long)low_len, (long long)idx);
  char msg[64];
  am_count_matches(a, 9, &hay, &c0);
  am_last_error_copy(msg, sizeof msg);
  printf("error=%s\n", msg);
  cudaFree(d); cudaFree(d_out);
  am_automaton_free(a);
  printf("done\n");
  return 0;
}
