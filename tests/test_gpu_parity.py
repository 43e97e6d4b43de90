"""Parity of the CUDA path with the CPU oracle (bit-exact: same (CodeUnitIndex, needle) sequence).

All calls go through the C ABI (ctypes).  `-m gpu`: needs a B200.
"""
import ctypes as C

import numpy as np
import pytest

from helpers import as_pairs, naive_find_all, needles_haystack

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def am():
    import alfred_margaret_b200 as pkg
    from alfred_margaret_b200 import _ffi
    assert _ffi.lib().am_device_count() >= 1, "no sm_100 device: the CUDA path cannot run"
    return pkg


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available()
    return torch


def gpu_pairs(m, hay):
    return as_pairs(m.find_all(hay))


def machine(am, needles, cs=0, **kw):
    return am.automaton.AcMachine([(n, i) for i, n in enumerate(needles)], case_sensitivity=cs, **kw)


# ---- the reference's own vectors, through the CUDA path ----------------------------------------------
def test_golden_count(am, golden):
    for v in golden["count"]:
        m = machine(am, v["needles"], v["cs"])
        assert m.count_matches(v["haystack"]) == v["expected"], v["src"]


def test_golden_find_all(am, golden):
    for v in golden["find_all"]:
        for kind in (0, 1):
            m = machine(am, v["needles"], force_kernel=kind)
            got = [(p, v["needles"][i]) for p, i in gpu_pairs(m, v["haystack"])]
            assert got == [tuple(x) for x in v["expected"]], v["src"]
    v = golden["example_file"]
    m = machine(am, v["needles"])
    assert m.count_matches(v["haystack"]) == v["expected_count"]
    assert [p for p, _ in gpu_pairs(m, v["haystack"])] == v["expected_end_positions"]


def test_golden_contains_any(am, golden):
    for v in golden["contains_any"]:
        s = am.searcher.build(v["cs"], v["needles"])
        assert am.searcher.contains_any(s, v["haystack"]) is v["expected"], v["src"]


def test_golden_quirks(am, golden):
    for v in golden["derived_quirks"]:
        m = machine(am, v["needles"])
        assert m.info()["kernel_kind"] == 1           # empty needle => general walk kernel
        assert gpu_pairs(m, v["haystack"]) == [tuple(x) for x in v["expected"]], v["src"]


def test_run_text_fold_and_early_exit(am):
    """runText with a fold (Automaton.hs:539-541); Done stops the fold (:530-532)."""
    A = am.automaton
    m = A.build([("tshirt", "T"), ("shirts", "S"), ("shorts", "O")])
    acc = A.run_text([], lambda a, mt: A.Step([(mt.pos, mt.value)] + a), m, "sweatshirts and shirtshirts")
    assert acc == [(27, "S"), (26, "T"), (22, "S"), (11, "S"), (10, "T")]          # README.md:94-100 (prepend order)
    first = A.run_text(None, lambda a, mt: A.Done((mt.pos, mt.value)), m, "sweatshirts and shirtshirts")
    assert first == (10, "T")
    n = A.run_text(0, lambda a, _: A.Step(a + 1), m, "short tshirts")
    assert n == 2
    # the SAME machine runs in the other case mode too, with no rebuild (Automaton.hs:539-553)
    assert A.run_lower(0, lambda a, _: A.Step(a + 1), m, "Short TSHIRTS") == 2
    assert A.run_text(0, lambda a, _: A.Step(a + 1), m, "Short TSHIRTS") == 0


def test_one_machine_both_case_modes(am, golden, oracle, lower_dense):
    """The reference's own test helper (tests/Data/Text/AhoCorasickSpec.hs:252-261):
        countMatches caseSensitivity needles haystack =
          let act = Aho.build $ zip needles (repeat ()); onMatch !n _ = Aho.Step (n + 1)
          in  Aho.runWithCase caseSensitivity 0 onMatch act haystack
    `Aho.build` knows nothing of the case mode; ONE machine must serve both."""
    A = am.automaton

    def count_matches(case_sensitivity, needles, haystack, act=None):
        act = act or A.build(list(zip(needles, [()] * len(needles))))
        return A.run_with_case(case_sensitivity, 0, lambda n, _: A.Step(n + 1), act, haystack)

    for v in golden["count"]:
        assert count_matches(v["cs"], v["needles"], v["haystack"]) == v["expected"], v["src"]
    rng = np.random.default_rng(77)
    for it in range(60):
        needles, hay = needles_haystack(rng, big=60)
        ln = [am.utf8.lower_utf8(n) for n in needles]
        hb = hay.encode("utf-8")
        act = A.build([(n, ()) for n in ln], force_kernel=int(rng.integers(0, 3)) if all(ln) else 0)
        om = oracle.Machine(ln)
        for cs in (0, 1, 0):                                   # alternate: the images of both modes live side by side
            want = as_pairs(om.find_all(hb, cs=cs, lower=lower_dense))
            assert count_matches(cs, ln, hb, act) == len(want)
            assert as_pairs(act.find_all(hb, case=cs)) == want
    # Searcher.setCaseSensitivity flips the flag and keeps needles and automaton (Searcher.hs:142-145)
    S = am.searcher
    s0 = S.build(0, ["tshirt", "shirts", "shorts"])
    s1 = S.set_case_sensitivity(1, s0)
    assert S.automaton(s1).handle is S.automaton(s0).handle or S.automaton(s1).handle.value == S.automaton(s0).handle.value
    assert S.contains_any(s0, "Short TSHIRTS") is False and S.contains_any(s1, "Short TSHIRTS") is True
    assert S.case_sensitivity(s1) == 1 and S.case_sensitivity(S.set_case_sensitivity(0, s1)) == 0


def test_two_devices_in_one_process(am, oracle, torch_cuda):
    """am_options.device: the dynamic shared-memory opt-in, the SM count and the profiling events are per device."""
    if torch_cuda.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from alfred_margaret_b200 import synth
    needles = synth.random_needles(200, 42)
    hay = synth.fill_host(0, 1 << 20, 43)
    synth.plant_host(hay, 0, 44, needles)
    want = as_pairs(oracle.Machine(needles).find_all(hay))
    before = torch_cuda.cuda.current_device()
    for dev in (0, 1, 0):
        for kind in (2, 1):
            m = machine(am, needles, device=dev, force_kernel=kind)
            assert gpu_pairs(m, hay) == want and m.count_matches(hay) == len(want)
    assert torch_cuda.cuda.current_device() == before           # the library restores the caller's current device


# ---- differential tests on the reference's generator ---------------------------------------------------
@pytest.mark.parametrize("kind", [0, 1])
def test_random_case_sensitive(am, oracle, kind):
    rng = np.random.default_rng(100 + kind)
    for it in range(150):
        needles, hay = needles_haystack(rng, big=60)
        hb = hay.encode("utf-8")
        m = machine(am, needles, force_kernel=kind)
        want = as_pairs(oracle.Machine(needles).find_all(hb))
        assert gpu_pairs(m, hb) == want, (needles, hay)
        assert m.count_matches(hb) == len(want)
        assert m.contains_any(hb) == (len(want) > 0)
        # Text slices with off != 0 (TestInstances.hs:26-33)
        pad = int(rng.integers(1, 20))
        buf = np.frombuffer(b"\xff" * pad + hb + b"\xff" * 3, dtype=np.uint8)
        assert gpu_pairs(m, am.utf8.Text(buf, pad, len(hb))) == want


@pytest.mark.parametrize("kind", [0, 1])
def test_random_ignore_case(am, oracle, lower_dense, kind):
    """kind 0: filter kernel on a lowered copy of the text (falls back to the walk when a code point changes
    length under lowering); kind 1: the per-code-point walk kernel."""
    rng = np.random.default_rng(200 + kind)
    for it in range(150):
        needles, hay = needles_haystack(rng, big=60)
        hb = hay.encode("utf-8")
        ln = [am.utf8.lower_utf8(n) for n in needles]
        m = machine(am, ln, cs=1, force_kernel=kind)
        want = as_pairs(oracle.Machine(ln).find_all(hb, cs=1, lower=lower_dense))
        assert gpu_pairs(m, hb) == want, (needles, hay)
        assert m.count_matches(hb) == len(want)
        assert m.contains_any(hb) == (len(want) > 0)


def test_contains_all(am, oracle, lower_dense):
    rng = np.random.default_rng(300)
    S = am.searcher
    assert S.contains_all(S.build_needle_id_searcher(0, [""]), "abc") is False     # AhoCorasickSpec.hs:196-200
    for it in range(60):
        needles, hay = needles_haystack(rng)
        needles = [n for n in needles if n]
        hb = hay.encode("utf-8")
        assert S.contains_all(S.build_needle_id_searcher(0, needles), hb) == all(n.encode() in hb for n in needles)
        ln = [am.utf8.lower_utf8(n) for n in needles]
        lh = am.utf8.lower_utf8(hb)
        assert S.contains_all(S.build_needle_id_searcher(1, ln), hb) == all(n in lh for n in ln)


def test_edge_cases(am, oracle):
    m = machine(am, ["abc", "bcd"])
    assert gpu_pairs(m, "") == [] and m.count_matches("") == 0 and m.contains_any("") is False
    assert gpu_pairs(m, "ab") == []                       # shorter than every needle
    assert gpu_pairs(m, "abcd") == [(3, 0), (4, 1)]
    assert machine(am, []).count_matches("abc") == 0      # no needles
    # duplicates, nested needles, a needle equal to the whole text, single-byte needles
    needles = ["a", "aa", "aaa", "a", "aaaa"]
    hay = "aaaa"
    assert gpu_pairs(machine(am, needles), hay) == as_pairs(oracle.Machine(needles).find_all(hay))
    # dense matches: every position matches (exercises queue overflow / staging overflow paths)
    hay = "a" * 70000
    for kind in (0, 1):
        mm = machine(am, ["a", "aa"], force_kernel=kind)
        got = mm.find_all(hay)
        assert len(got) == 70000 + 69999 == mm.count_matches(hay)
        assert as_pairs(got) == as_pairs(oracle.Machine(["a", "aa"]).find_all(hay))
    # a long needle (halo larger than a segment)
    long_needle = "xy" * 700
    hay = "ab" * 3000 + long_needle + "ab" * 100 + long_needle[:-1]
    for kind in (0, 1):
        assert gpu_pairs(machine(am, [long_needle, "ba"], force_kernel=kind), hay) == as_pairs(oracle.Machine([long_needle, "ba"]).find_all(hay))


# ---- device-resident calls, shards, alignment --------------------------------------------------------------
def test_device_resident_and_unaligned(am, oracle, torch_cuda):
    torch = torch_cuda
    from alfred_margaret_b200 import synth
    needles = synth.random_needles(300, 42, 2, 9, b"abcd")
    host = synth.fill_host(0, 1 << 20, 43, b"abcd")
    synth.plant_host(host, 0, 44, needles)
    om = oracle.Machine(needles)
    dev = torch.empty(host.size + 64, dtype=torch.uint8, device="cuda")
    for kind in (0, 1):
        m = machine(am, needles, force_kernel=kind)
        for off in (0, 1, 5, 16, 33):
            view = dev[off:off + host.size]
            view.copy_(torch.from_numpy(host))
            want = as_pairs(om.find_all(host))
            assert m.count_matches_dev(view.data_ptr(), host.size) == len(want)
            out = torch.empty((len(want) + 8) * 2, dtype=torch.int64, device="cuda")
            n = m.find_all_dev(view.data_ptr(), host.size, out.data_ptr(), len(want) + 8)
            assert n == len(want)
            rec = out[: 2 * n].cpu().numpy().view(am.automaton.MATCH_DTYPE)
            assert as_pairs(rec) == want
            with pytest.raises(OverflowError):
                m.find_all_dev(view.data_ptr(), host.size, out.data_ptr(), max(0, len(want) - 1))
        # shards: matches are partitioned by end position, halo = max needle length - 1
        view = dev[:host.size]
        view.copy_(torch.from_numpy(host))
        halo = m.info()["halo_bytes"]
        L = am._ffi.lib()
        want = as_pairs(om.find_all(host))
        # a one-rank communicator: the sharded entry points of the C ABI without a second GPU
        comm = am.sharded.Comm(0, 1, None)
        out = torch.empty((len(want) + 8) * 2, dtype=torch.int64, device="cuda")
        assert comm.count(m, view.data_ptr(), host.size) == (len(want), 0, len(want))
        assert comm.find_all(m, view.data_ptr(), host.size, out.data_ptr(), len(want) + 8) == (len(want), 0, len(want))
        assert as_pairs(out[: 2 * len(want)].cpu().numpy().view(am.automaton.MATCH_DTYPE)) == want
        assert comm.contains_any(m, view.data_ptr(), host.size) is True and comm.allreduce(5, "max") == 5
        comm.close()
        for n_shards in (2, 3, 8):
            got, total = [], 0
            for r in range(n_shards):
                w, b, e = C.c_uint64(), C.c_uint64(), C.c_uint64()
                assert L.am_shard_plan(host.size, halo, n_shards, r, C.byref(w), C.byref(b), C.byref(e)) == 0
                ptr = view.data_ptr() + w.value
                ln = e.value - w.value
                rb = b.value - w.value
                cnt = m.count_matches_dev(ptr, ln, report_begin=rb, pos_base=w.value)
                out = torch.empty((cnt + 1) * 2, dtype=torch.int64, device="cuda")
                n = m.find_all_dev(ptr, ln, out.data_ptr(), cnt + 1, report_begin=rb, pos_base=w.value)
                assert n == cnt
                got += as_pairs(out[: 2 * n].cpu().numpy().view(am.automaton.MATCH_DTYPE))
                total += cnt
            assert total == len(want) and got == want


def test_synth_device_matches_host(am, torch_cuda):
    torch = torch_cuda
    from alfred_margaret_b200 import synth
    needles = synth.random_needles(64, 42)
    n = (1 << 20) + 123
    for first in (0, 4096 * 3 + 17):
        dev = torch.empty(n, dtype=torch.uint8, device="cuda")
        synth.fill_dev(dev.data_ptr(), n, first, 43)
        synth.plant_dev(dev.data_ptr(), n, first, 44, needles)
        host = synth.fill_host(first, n, 43)
        synth.plant_host(host, first, 44, needles)
        assert np.array_equal(dev.cpu().numpy(), host)


def test_config2_downscaled(am, oracle, torch_cuda):
    """BASELINE.json config 2 (1 000 needles 4-16 B, a-z) on 64 MiB: full list equality + checksums."""
    torch = torch_cuda
    from alfred_margaret_b200 import synth
    needles = synth.random_needles(1000, 42)
    n = 64 << 20
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    synth.fill_dev(dev.data_ptr(), n, 0, 43)
    synth.plant_dev(dev.data_ptr(), n, 0, 44, needles)
    host = dev.cpu().numpy()
    om = oracle.Machine(needles)
    want = om.find_all(host)
    assert len(want) > n // 4096                      # every block got its planted needle
    for kind in (0, 1):
        m = machine(am, needles, force_kernel=kind)
        assert m.info()["kernel_kind"] == (2 if kind == 0 else 1)
        assert m.count_matches_dev(dev.data_ptr(), n) == len(want)
        assert m.contains_any_dev(dev.data_ptr(), n) is True
        out = torch.empty(len(want) * 2, dtype=torch.int64, device="cuda")
        k = m.find_all_dev(dev.data_ptr(), n, out.data_ptr(), len(want))
        rec = out[: 2 * k].cpu().numpy().view(am.automaton.MATCH_DTYPE)
        assert k == len(want)
        assert np.array_equal(rec["end_pos"].astype(np.int64), want["pos"]) and np.array_equal(rec["needle_id"].astype(np.int64), want["value"])
    digits = torch.full((1 << 20,), ord("7"), dtype=torch.uint8, device="cuda")      # all-miss haystack
    assert m.contains_any_dev(digits.data_ptr(), digits.numel()) is False


def test_golden_splitter(am, golden):
    """Splitter: AhoCorasickSpec.hs:220-244."""
    S = am.splitter
    for v in golden["splitter"]:
        sp = S.build(v["separator"])
        got = S.split_ignore_case(sp, v["haystack"]) if v["ignore_case"] else S.split(sp, v["haystack"])
        assert [x.decode("utf-8") for x in got] == v["expected"], v["src"]
    # overlapping separators are ignored left to right (stepAccum, Splitter.hs:158-170)
    assert S.split(S.build("aa"), "aaaaa") == [b"", b"", b"a"]
    assert S.split(S.build("x"), "") == [b""]


def test_ignore_case_length_preserving_text(am, oracle, lower_dense):
    """Mixed-case text whose lowerings all keep their UTF-8 length (ASCII, Latin-1, Greek, Cyrillic, 4-byte
    code points, multi-byte code points straddling 16-byte granules): the filter kernel's fast IgnoreCase path."""
    rng = np.random.default_rng(77)
    cps = list("abcdefghijABCDEFGHIJ .,") * 3 + list("éÉöÖßåÅяЯжЖωΩλΛ") + list("𝄞💩€") + list("ǳǲǱ")
    needles = sorted({"".join(str(c).lower() for c in rng.choice(cps, size=int(rng.integers(2, 7)))) for _ in range(400)})
    needles = [n for n in needles if n.strip()]
    hay = "".join(rng.choice(cps, size=300000)).encode("utf-8")
    ln = [am.utf8.lower_utf8(n) for n in needles]
    want = oracle.Machine(ln).find_all(hay, cs=1, lower=lower_dense, cap=1 << 20)
    assert len(want) > 10000
    for kind in (0, 1):
        m = machine(am, ln, cs=1, force_kernel=kind)
        for off in (0, 3):                                   # unaligned slices too
            buf = np.frombuffer(b"\xc3" * off + hay, dtype=np.uint8)
            got = m.find_all(am.utf8.Text(buf, off, len(hay)))
            assert len(got) == len(want) and np.array_equal(got["end_pos"].astype(np.int64), want["pos"]) and np.array_equal(got["needle_id"].astype(np.int64), want["value"]), (kind, off)
        assert m.count_matches(hay) == len(want)


def test_ignore_case_length_changing_variants(am, oracle, lower_dense):
    """Code points whose lower case has another UTF-8 length (K U+212A -> k, Å U+212B -> å, ẞ -> ß, İ -> i,
    Ⱥ -> ⱥ, Ⱦ -> ⱦ): the filter path keeps them in the lowered copy and matches them through needle variants
    (am_build.cpp step 1); needles that are not lower case can never match (Automaton.hs:478-480); a needle
    with too many variants switches the automaton to the mark-and-fall-back scheme.  All must equal runLower."""
    rng = np.random.default_rng(91)
    cps = list("kKaAiIsSbt ") * 2 + list("\u212a\u212båÅßẞ\u0130ⱥȺⱦȾ") * 2 + list("é𝄞")
    hay = "".join(rng.choice(cps, size=200000)).encode("utf-8")
    pool = list("kaist") + list("åßⱥⱦé")
    base = sorted({"".join(rng.choice(pool, size=int(rng.integers(1, 6)))) for _ in range(300)})
    sets = {
        "variants": base,
        "dead needles": base[:50] + ["K", "\u212a", "ak\u212b", "Åk", "ẞ", "İi", "kȺ"],    # never match
        "many variants": base[:50] + ["kkkkkkk", "iiißßßkk"],                             # 128 / 256 variants each
        "fallback": base[:50] + ["k" * 13, "iiißßßkkiißßk"],                              # > 4 096 variants each
        "empty needle": base[:20] + [""],
    }
    for name, needles in sets.items():
        ln = [n.encode("utf-8") for n in needles]
        want = oracle.Machine(ln).find_all(hay, cs=1, lower=lower_dense, cap=1 << 22)
        assert len(want) > 1000
        for kind in (0, 1, 2):
            if kind == 2 and "" in needles:
                continue
            m = machine(am, ln, cs=1, force_kernel=kind)
            for off in (0, 5):
                buf = np.frombuffer(b"\xe2" * off + hay, dtype=np.uint8)
                got = m.find_all(am.utf8.Text(buf, off, len(hay)))
                assert len(got) == len(want) and np.array_equal(got["end_pos"].astype(np.int64), want["pos"]) and np.array_equal(got["needle_id"].astype(np.int64), want["value"]), (name, kind, off)
            assert m.count_matches(hay) == len(want), (name, kind)
    # a dead needle alone: no match at all, on either kernel
    for kind in (0, 1):
        m = machine(am, ["\u212a".encode(), b"K"], cs=1, force_kernel=kind)
        assert m.count_matches(hay) == 0 and not m.contains_any(hay)


def test_config3_downscaled_ignore_case(am, oracle, lower_dense):
    """BASELINE.json config 3 down-scaled: 2 000 lower-case needles (20 % with non-ASCII code points), IgnoreCase,
    an 8 MiB mixed-case UTF-8 haystack with 1/2/3/4-byte code points incl. length-changing ones (K, Å, ẞ, İ, Ⱥ)."""
    rng = np.random.default_rng(52)
    ascii_l = "abcdefghijklmnopqrstuvwxyz"
    extra = "éößåяωǳⱥ"
    needles = set()
    while len(needles) < 2000:
        n = int(rng.integers(4, 17))
        pool = ascii_l + (extra * 3 if rng.random() < 0.2 else "")
        needles.add("".join(pool[int(i)] for i in rng.integers(0, len(pool), size=n)))
    needles = sorted(needles)
    cps = list(ascii_l + ascii_l.upper()) * 6 + list(" .,;-") * 4 + list("éÉöÖßåÅяЯωΩǳǲǱ") * 2 + list("ẞKÅȺⱥİ") + list("𝄞💩")
    parts = []
    for _ in range(6000):
        if rng.random() < 0.3:   # plant a needle with random per-code-point upper-casing
            n = needles[int(rng.integers(0, len(needles)))]
            parts.append("".join((c.upper() if (rng.random() < 0.5 and len(c.upper()) == 1) else c) for c in n))
        parts.append("".join(cps[int(i)] for i in rng.integers(0, len(cps), size=int(rng.integers(50, 400)))))
    hay = "".join(parts).encode("utf-8")
    m = machine(am, needles, cs=1)
    want = oracle.Machine(needles).find_all(hay, cs=1, lower=lower_dense)
    assert len(want) > 1500
    got = m.find_all(hay)
    assert len(got) == len(want) and np.array_equal(got["end_pos"].astype(np.int64), want["pos"]) and np.array_equal(got["needle_id"].astype(np.int64), want["value"])
    assert m.count_matches(hay) == len(want)


def test_large_set_three_filter_levels(am, oracle, lower_dense):
    """Needle sets beyond the exact second level on 4-grams (> 2 048 distinct q-grams): warp-cooperative bitmap level, the third
    level in global memory (prefixes of min(length, 8) bytes, every case variant) and verify_kernel's queue of the survivors that
    pass it.  IgnoreCase with non-ASCII needles on mixed-case UTF-8 text (C3's shape, 6 000 needles) and the same machine
    CaseSensitive; count, full list and containsAny (match present / absent: the queue's block-wide early exit)."""
    from alfred_margaret_b200 import workloads
    needles = [n.decode("utf-8") for n in workloads.c3_needles(6000)]
    hay = workloads.c3_unit([n.encode("utf-8") for n in needles])[: 3 << 20].copy()
    while (hay[-1] & 0xC0) == 0x80 or hay[-1] >= 0xC0: hay = hay[:-1]          # end on a code point boundary
    hb = hay.tobytes()
    om = oracle.Machine(needles)
    m = machine(am, needles, cs=1)
    assert m.info()["kernel_kind"] == 2
    for cs in (1, 0):
        want = om.find_all(hb, cs=cs, lower=lower_dense if cs else None, cap=1 << 22)
        assert len(want) > (300 if cs else 5)
        got = m.find_all(hb, case=cs)
        assert len(got) == len(want) and np.array_equal(got["end_pos"].astype(np.int64), want["pos"]) and np.array_equal(got["needle_id"].astype(np.int64), want["value"]), cs
        assert m.count_matches(hb, case=cs) == len(want)
        assert m.contains_any(hb, case=cs) is True
    quiet = ("zq " * 300000).encode()                          # no needle holds "zq ": every survivor dies in the levels or the check
    assert len(om.find_all(quiet, cs=1, lower=lower_dense)) == 0
    assert m.contains_any(quiet, case=1) is False and m.count_matches(quiet, case=1) == 0


def test_config5_downscaled_many_needles(am, oracle, torch_cuda):
    """BASELINE.json config 5 down-scaled: 100 000 needles (6-16 B), 32 MiB haystack in 4 shards with halos;
    per-shard lists concatenate to the single-shard list; counts all-gathered (here: summed) match."""
    torch = torch_cuda
    from alfred_margaret_b200 import sharded, synth
    needles = synth.random_needles(100000, 72, 6, 16)
    n = 32 << 20
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    synth.fill_dev(dev.data_ptr(), n, 0, 73)
    synth.plant_dev(dev.data_ptr(), n, 0, 74, needles)
    host = dev.cpu().numpy()
    want = oracle.Machine(needles).find_all(host, threads=8, cap=1 << 20)
    m = machine(am, needles)
    halo = m.info()["halo_bytes"]
    parts = []
    for r in range(4):
        w, b, e = sharded.shard_plan(n, halo, 4, r)
        cnt = m.count_matches_dev(dev.data_ptr() + w, e - w, report_begin=b - w, pos_base=w)
        out = torch.empty(2 * (cnt + 1), dtype=torch.int64, device="cuda")
        k = m.find_all_dev(dev.data_ptr() + w, e - w, out.data_ptr(), cnt + 1, report_begin=b - w, pos_base=w)
        assert k == cnt
        parts.append(out[: 2 * k].cpu().numpy().view(am.automaton.MATCH_DTYPE))
    got = np.concatenate(parts)
    assert len(got) == len(want) and np.array_equal(got["end_pos"].astype(np.int64), want["pos"]) and np.array_equal(got["needle_id"].astype(np.int64), want["value"])


def test_large_positions_and_kernel_choice(am, oracle, torch_cuda):
    """pos_base beyond 2^32 (sort keys use bits(len + pos_base) + rank bits); kernel heuristic for big needle sets."""
    torch = torch_cuda
    from alfred_margaret_b200 import synth
    needles = synth.random_needles(500, 5, 3, 8, b"abc")
    n = 1 << 20
    host = synth.fill_host(0, n, 6, b"abc")
    want = oracle.Machine(needles).find_all(host, cap=1 << 20)
    dev = torch.from_numpy(host).cuda()
    base = (1 << 41) + 12345
    for kind in (0, 1):
        m = machine(am, needles, force_kernel=kind)
        out = torch.empty(2 * (len(want) + 1), dtype=torch.int64, device="cuda")
        k = m.find_all_dev(dev.data_ptr(), n, out.data_ptr(), len(want) + 1, pos_base=base)
        rec = out[: 2 * k].cpu().numpy().view(am.automaton.MATCH_DTYPE)
        assert k == len(want) and np.array_equal(rec["end_pos"].astype(np.int64) - base, want["pos"]) and np.array_equal(rec["needle_id"].astype(np.int64), want["value"])
    assert machine(am, synth.random_needles(1000, 42)).info()["kernel_kind"] == 2
    assert machine(am, synth.random_needles(40000, 43, 6, 12)).info()["kernel_kind"] == 2
    big = machine(am, synth.random_needles(100000, 43, 6, 12))
    assert big.info()["kernel_kind"] == 2                      # shortest needle 6 bytes: 6-grams keep the filter selective
    walk = machine(am, synth.random_needles(100000, 43, 6, 12), force_kernel=1)
    hay = synth.fill_host(0, 1 << 20, 44)
    synth.plant_host(hay, 0, 45, synth.random_needles(100000, 43, 6, 12))
    assert len(big.find_all(hay)) > 200 and as_pairs(walk.find_all(hay)) == as_pairs(big.find_all(hay))
    four = machine(am, synth.random_needles(100000, 46, 4, 12))
    assert four.info()["kernel_kind"] == 1                     # 4-byte needles force 4-grams: too many for the shared-memory bitmap


@pytest.mark.parametrize("lens", [(6, 16), (8, 16), (6, 7), (8, 8)])
def test_long_qgram_filter(am, oracle, torch_cuda, lens):
    """Needle sets beyond the exact second level whose shortest needle has >= 6 (>= 8) bytes take 6- (8-)gram stride-2 cells
    and a Bloom second level; the jump table is keyed by the whole q-gram.  Both kernels against the oracle, at every
    alignment of the device text, with prefix-sharing needles (same first four bytes, different q-gram) and duplicates."""
    torch = torch_cuda
    from alfred_margaret_b200 import synth
    needles = synth.random_needles(6000, 91 + lens[0], lens[0], lens[1], b"abcdefgh")
    stem = needles[0][:4]
    needles += [stem + b"abcd" + b"x" * (lens[0] - 4), stem + b"abce" + b"x" * (lens[0] - 4), stem + b"bbbbbbbb", needles[1], needles[2] + b"zz"]
    n = (2 << 20) + 77
    host = synth.fill_host(0, n, 7, b"abcdefgh")
    synth.plant_host(host, 0, 8, needles, block=256)
    want = oracle.Machine(needles).find_all(host, threads=4, cap=1 << 22)
    assert len(want) > 8000
    m = machine(am, needles)
    assert m.info()["kernel_kind"] == 2
    dev = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
    for off in (0, 1, 2, 7, 16):
        view = dev[off:off + n]
        view.copy_(torch.from_numpy(host))
        assert m.count_matches_dev(view.data_ptr(), n) == len(want)
        out = torch.empty(2 * (len(want) + 1), dtype=torch.int64, device="cuda")
        k = m.find_all_dev(view.data_ptr(), n, out.data_ptr(), len(want) + 1)
        rec = out[: 2 * k].cpu().numpy().view(am.automaton.MATCH_DTYPE)
        assert k == len(want) and np.array_equal(rec["end_pos"].astype(np.int64), want["pos"]) and np.array_equal(rec["needle_id"].astype(np.int64), want["value"]), off
    assert m.contains_any(host) is True and m.contains_any(b"zzzzzzzzzzzzzzzzzzzz" * 100) is False


def test_launch_span_boundary(am, oracle, torch_cuda):
    """Texts longer than one filter-kernel launch span (2^31 bytes): matches around the launch boundary equal the
    oracle's, and the total equals the sum over shards cut elsewhere (size-independent consistency property)."""
    torch = torch_cuda
    from alfred_margaret_b200 import sharded, synth
    needles = synth.random_needles(1000, 42)
    n = (1 << 31) + (96 << 20) + 7
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    synth.fill_dev(dev.data_ptr(), n, 0, 43)
    synth.plant_dev(dev.data_ptr(), n, 0, 44, needles)
    m = machine(am, needles)
    total = m.count_matches_dev(dev.data_ptr(), n)
    out = torch.empty(2 * (total + 1), dtype=torch.int64, device="cuda")
    assert m.find_all_dev(dev.data_ptr(), n, out.data_ptr(), total + 1) == total
    rec = out[: 2 * total].cpu().numpy().view(am.automaton.MATCH_DTYPE)
    assert np.all(np.diff(rec["end_pos"].astype(np.int64)) >= 0)                  # sortedness at full size
    # window of 2 MiB around the 2^31 launch boundary vs the oracle
    lo, hi = (1 << 31) - (1 << 20), (1 << 31) + (1 << 20)
    window = dev[lo:hi].cpu().numpy()
    want = oracle.Machine(needles).find_all(window, cap=1 << 16)
    sel = rec[(rec["end_pos"] > lo + 16) & (rec["end_pos"] <= hi)]
    want = want[want["pos"] > 16]
    assert len(sel) == len(want) and np.array_equal(sel["end_pos"].astype(np.int64) - lo, want["pos"]) and np.array_equal(sel["needle_id"].astype(np.int64), want["value"])
    # shard sums (3 shards, cuts not at the launch boundary)
    halo = m.info()["halo_bytes"]
    s = 0
    for r in range(3):
        w, b, e = sharded.shard_plan(n, halo, 3, r)
        s += m.count_matches_dev(dev.data_ptr() + w, e - w, report_begin=b - w, pos_base=w)
    assert s == total


def test_ac_bench_protocol(am, golden, tmp_path, capsys):
    """tools/ac_bench.py speaks the reference harness's protocol (5 timings on stdout, count on stderr)."""
    import importlib.util, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("ac_bench", os.path.join(root, "alfred-margaret_b200", "tools", "ac_bench.py"))
    ac = importlib.util.module_from_spec(spec); spec.loader.exec_module(ac)
    v = golden["example_file"]
    p = tmp_path / "example.txt"
    p.write_bytes(("\n".join(v["needles"]) + "\n\n" + v["haystack"]).encode("utf-8"))
    ac.main([str(p)])
    cap = capsys.readouterr()
    assert cap.err.strip() == str(v["expected_count"])
    assert len([x for x in cap.out.strip().split("\t") if x]) == 5


def test_stride2_cells_and_tail_compare(am, oracle, torch_cuda):
    """q = 4 filter: one bitmap word answers two start positions (even p: cell A, odd p + 1: cell B) and survivors on
    a single needle path are checked by a tail comparison.  Dense small-alphabet text so that needles start at every
    residue of the 16-byte granule, the 1 KiB pair, the 4 KiB warp chunk and the 128 KiB CTA tile; needle sets with
    shared prefixes (not a single path), duplicates (several ranks at one leaf), a needle of exactly q bytes, bytes
    that share their low 5 bits (cells collide, only false positives), multi-byte UTF-8; unaligned device pointers."""
    torch = torch_cuda
    rng = np.random.default_rng(4242)
    sets = [
        ["abca", "bcab", "cabc", "abcab", "abcabc", "bca" + "abc" * 4, "abca"],          # prefixes of each other, duplicate
        ["aaaa", "aaaab", "baaaa", "abab", "ababab", "bbbbbbbbbbbbbbbb"],                  # self-overlapping
        ["aAb!", "!bAa", "Aa!b1", "a!A!a!A", "ABAB", "abab"],                             # 'a' 'A' '!' share their low 5 bits
        ["åbcå", "💩ab", "ab💩", "ßßab", "abßß", "ẞẞ"],                                   # multi-byte code points
    ]
    alphabets = ["abc", "ab", "aA!b1B", "abåß💩ẞ"]
    for needles, alpha in zip(sets, alphabets):
        n_cp = (300 << 10) + int(rng.integers(0, 64))
        idx = rng.integers(0, len(alpha), size=n_cp)
        hb = "".join(alpha[int(i)] for i in idx[:n_cp]).encode("utf-8")
        want = oracle.Machine(needles).find_all(np.frombuffer(hb, dtype=np.uint8), cap=len(hb) * 4)
        m = machine(am, needles, force_kernel=2)
        assert m.info()["kernel_kind"] == 2
        for shift in (0, 1, 7, 15):
            dev = torch.empty(len(hb) + 64, dtype=torch.uint8, device="cuda")
            dev[shift:shift + len(hb)] = torch.frombuffer(bytearray(hb), dtype=torch.uint8).cuda()
            ptr = dev.data_ptr() + shift
            n = m.count_matches_dev(ptr, len(hb))
            assert n == len(want), (needles, shift)
            out = torch.empty(2 * (n + 16), dtype=torch.int64, device="cuda")
            got_n = m.find_all_dev(ptr, len(hb), out.data_ptr(), n + 16)
            assert got_n == n
            got = out.cpu().numpy()[: 2 * n].view(am.automaton.MATCH_DTYPE)
            assert np.array_equal(got["end_pos"].astype(np.int64), want["pos"]) and np.array_equal(got["needle_id"].astype(np.int64), want["value"]), (needles, shift)


def test_host_scans_pipelined_in_chunks(am, oracle):
    """am_count_matches / am_contains_any with HOST buffers upload the text in 64 MiB chunks while the previous chunk is
    scanned (each chunk like a shard: halo before it, matches that end inside it); containsAny stops at the first chunk
    with a match.  Matches that straddle the chunk boundaries must be counted exactly once."""
    from alfred_margaret_b200 import synth
    needles = synth.random_needles(1000, 42)
    CH = 64 << 20
    n = 3 * CH + 12345
    hay = synth.fill_host(0, n, 7, b"0123456789")                 # no needle byte at all
    m = machine(am, needles)
    assert m.count_matches(hay) == 0 and m.contains_any(hay) is False
    long_needle = max(needles, key=len)
    spots = [CH - len(long_needle) // 2, 2 * CH - len(long_needle), 2 * CH, 3 * CH - len(long_needle) + 1, n - len(long_needle)]
    for at in spots:                                              # across / at / next to every chunk boundary, and at the very end
        hay[at:at + len(long_needle)] = np.frombuffer(long_needle, dtype=np.uint8)
    want = oracle.Machine(needles).find_all(hay, threads=8, cap=1 << 16)
    assert len(want) >= len(spots)
    assert m.count_matches(hay) == len(want)
    assert m.contains_any(hay) is True
    got = m.find_all(hay)
    assert len(got) == len(want) and np.array_equal(got["end_pos"].astype(np.int64), want["pos"]) and np.array_equal(got["needle_id"].astype(np.int64), want["value"])
    # IgnoreCase takes the same chunked path (lowered copy per chunk)
    mi = machine(am, needles, cs=1)
    up = hay.copy()
    for at in spots:
        up[at:at + len(long_needle)] = np.frombuffer(long_needle.upper(), dtype=np.uint8)
    assert mi.count_matches(up) == len(want) and mi.contains_any(up) is True and m.count_matches(up) == 0


def test_ignore_case_one_pass_any_text(am, oracle, lower_dense, torch_cuda):
    """runLower on the filter kernel is ONE pass over the original text, whatever it holds: the probe sees folded bytes,
    the survivors are lowered code point by code point.  ASCII text, text with single code points above ASCII at awkward
    places, and a text dense in multi-byte code points whose lower case changes the lead byte (Я), a continuation byte
    (É, Ω) or the UTF-8 length (K, ẞ: matched by needle variants) must all equal the oracle's runLower."""
    from alfred_margaret_b200 import synth
    needles = synth.random_needles(1000, 42)
    n = (4 << 20) + 123
    hay = synth.fill_host(0, n, 43, synth.AZ + synth.AZ.upper() + b" .,0123456789")
    synth.plant_host(hay, 0, 44, [x.upper() if i % 2 else x for i, x in enumerate(needles)], block=2048)
    om = oracle.Machine(needles)
    m = machine(am, needles, cs=1)
    assert m.info()["kernel_kind"] == 2
    for variant in ("ascii", "one code point above ASCII at the end", "one in the middle"):
        h = hay.copy()
        if variant != "ascii":
            at = n - 9 if "end" in variant else n // 2
            h[at:at + 2] = np.frombuffer("É".encode("utf-8"), dtype=np.uint8)
        want = om.find_all(h, cs=1, lower=lower_dense, cap=1 << 20)
        assert len(want) > 1000
        got = m.find_all(h)
        assert len(got) == len(want) and np.array_equal(got["end_pos"].astype(np.int64), want["pos"]) and np.array_equal(got["needle_id"].astype(np.int64), want["value"]), variant
        assert m.count_matches(h) == len(want) and m.contains_any(h) is True
        dev = torch_cuda.from_numpy(h).cuda()                      # device-resident, unaligned
        assert m.count_matches_dev(dev.data_ptr() + 3, n - 3) == len(om.find_all(h[3:], cs=1, lower=lower_dense, cap=1 << 20))
    # dense multi-byte text; needles with non-ASCII code points; exact and Bloom second levels (1 200 / 6 000 needles)
    rng = np.random.default_rng(9)
    for count, lo_len in ((1200, 4), (6000, 4), (6000, 6)):
        pool = list("abcdefghikst") * 2 + list("åßяéωǳ")
        nd = sorted({"".join(rng.choice(pool, size=int(rng.integers(lo_len, 11)))) for _ in range(count)})
        nb = [x.encode("utf-8") for x in nd]
        cps = list("abcdefghikst ") * 2 + list("ABCDEFGHIKST") + list("åßяéωǳ") + list("ÅẞЯÉΩǲǱKİ") + ["𝄞", "€"]
        up = {"å": "Å", "я": "Я", "é": "É", "ω": "Ω", "ǳ": "ǲ", "k": "K", "i": "İ", "ß": "ẞ"}
        parts = []
        for _ in range(1500):                                  # random text with a randomly re-cased needle every ~270 symbols
            parts.append("".join(rng.choice(cps, size=int(rng.integers(200, 340)))))
            w = nd[int(rng.integers(0, len(nd)))]
            parts.append("".join((up.get(c, c.upper()) if rng.random() < 0.5 else c) for c in w))
        text = "".join(parts)
        hb = np.frombuffer(text.encode("utf-8"), dtype=np.uint8).copy()
        want = oracle.Machine(nb).find_all(hb, cs=1, lower=lower_dense, cap=1 << 22)
        assert len(want) > 300, (count, len(want))
        for force in (0, 1):
            mm = machine(am, nb, cs=1, force_kernel=force)
            got = mm.find_all(hb)
            assert len(got) == len(want) and np.array_equal(got["end_pos"].astype(np.int64), want["pos"]) and np.array_equal(got["needle_id"].astype(np.int64), want["value"]), (count, lo_len, force)
            assert mm.count_matches(hb) == len(want)


def test_survivor_flood_hands_over_to_the_walk(am, oracle, torch_cuda):
    """Texts that defeat the q-gram filter -- "aaaa..." against needles of a's, a periodic text made of the needles' own
    q-grams: (nearly) every position survives both filter levels.  The survivor-rate monitor hands such a scan over to the
    per-segment walk (O(n) on any input, like the reference's loop, Automaton.hs:489-510); the results must not change, in
    any mode, also when only PART of a host text (one 64 MiB chunk of several) floods."""
    from alfred_margaret_b200 import synth
    n = (3 << 20) + 5
    cases = [([b"aaaa", b"aaaaaa", b"aaab", b"baaa"], np.full(n, ord("a"), dtype=np.uint8)),
             ([b"abcabc", b"bcabca", b"cabcab", b"abcabcabc"], np.frombuffer((b"abc" * (n // 3 + 1))[:n], dtype=np.uint8).copy())]
    for needles, hay in cases:
        hay[n // 2] = ord("b")
        want = oracle.Machine(needles).find_all(hay, threads=4, cap=4 * n)
        assert len(want) > n
        m = machine(am, needles)
        assert m.info()["kernel_kind"] == 2
        for _ in range(3):                                     # (the third scan starts on the walk kernel: two hand-overs in a row)
            got = m.find_all(hay)
            assert len(got) == len(want) and np.array_equal(got["end_pos"].astype(np.int64), want["pos"]) and np.array_equal(got["needle_id"].astype(np.int64), want["value"])
        assert m.count_matches(hay) == len(want) and m.contains_any(hay) is True
        dev = torch_cuda.from_numpy(hay).cuda()
        assert m.count_matches_dev(dev.data_ptr(), n) == len(want)
    # ordinary text before and after a flooded stretch: the filter comes back (every eighth scan tries it again)
    needles = synth.random_needles(500, 42) + [b"aaaa"]
    hay = synth.fill_host(0, 2 << 20, 43)
    synth.plant_host(hay, 0, 44, needles)
    hay[(1 << 20):(1 << 20) + 300000] = ord("a")
    want = oracle.Machine(needles).find_all(hay, threads=4, cap=1 << 22)
    m = machine(am, needles)
    for _ in range(10):
        got = m.find_all(hay)
        assert len(got) == len(want) and np.array_equal(got["end_pos"].astype(np.int64), want["pos"]) and np.array_equal(got["needle_id"].astype(np.int64), want["value"])
