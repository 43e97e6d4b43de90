"""Replacer.build / run / runWithLimit on the device vs the reference's vectors and the CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def am():
    import alfred_margaret_b200 as pkg
    from alfred_margaret_b200 import _ffi
    assert _ffi.lib().am_device_count() >= 1
    return pkg


def test_golden_replacer(am, golden):
    R = am.replacer
    for v in golden["replacer"]:
        r = R.build(v["cs"], [tuple(p) for p in v["pairs"]])
        assert R.run(r, v["haystack"]).decode("utf-8") == v["expected"], v["src"]


def test_run_with_limit(am):
    R = am.replacer
    r = R.build(0, [("a", "bbbb")])
    assert R.run_with_limit(r, 8, "aa") == b"bbbbbbbb"
    assert R.run_with_limit(r, 7, "aa") is None                      # Replacer.hs:240
    r = R.build(0, [("aaa", "xxxxx")])                               # replacementLength BEFORE removeOverlap
    assert R.run_with_limit(r, 6, "aaaa") is None and R.run_with_limit(r, 8, "aaaa") == b"xxxxxa"
    with pytest.raises(am._ffi.AmError):
        R.build(1, [("", "x")])                                      # empty needle + IgnoreCase: reference diverges


def test_set_case_sensitivity_keeps_the_stored_needles(am, oracle, lower_dense):
    """`setCaseSensitivity` (Replacer.hs:151-153) only flips the searcher's flag: the needles stay as `build` stored them
    (lowered iff built IgnoreCase, :105-107), the payload lengths stay those of the original needles (:111-113)."""
    R = am.replacer
    r = R.set_case_sensitivity(0, R.build(1, [("Foo", "x")]))        # stored needle: "foo"
    assert R.run(r, "foo Foo") == b"x Foo"                            # (a rebuild from "Foo" would give "foo x")
    assert R.replacer_case_sensitivity(r) == 0
    r = R.set_case_sensitivity(1, R.build(0, [("Foo", "x")]))        # stored needle: "Foo" -- never matches lowered text
    assert R.run(r, "foo Foo FOO") == b"foo Foo FOO"
    r = R.set_case_sensitivity(1, R.build(0, [("foo", "x")]))
    assert R.run(r, "foo Foo FOO") == b"x x x"
    # compose / mapReplacement work on the stored form: no second lowering, lengths kept (:120-141)
    a, b = R.build(1, [("ÉCLAIR", "bolt")]), R.build(1, [("BOLT", "Blitz")])
    c = R.compose(a, b)
    assert R.run(c, "un Éclair") == b"un Blitz"
    assert R.run(R.map_replacement(lambda rep: rep.upper(), c), "un éclair") == b"un BLITZ"
    assert R.to_json(c)["replacerSearcher"]["needles"][0] == ["éclair", {"needlePriority": 0, "needleLengthBytes": 7, "needleLengthCodePoints": 6, "needleReplacement": "bolt"}]
    assert R.run(R.from_json(R.to_json(c)), "un Éclair") == b"un Blitz"
    rng = np.random.default_rng(5)
    for it in range(40):                                             # both switches against the oracle on its generator
        alpha = "abAB"
        pairs = [("".join(alpha[int(i)] for i in rng.integers(0, 4, size=int(rng.integers(1, 4)))),
                  "".join(alpha[int(i)] for i in rng.integers(0, 4, size=int(rng.integers(0, 4))))) for _ in range(int(rng.integers(1, 5)))]
        hay = "".join(alpha[int(i)] for i in rng.integers(0, 4, size=int(rng.integers(0, 40))))
        lowered = [(n.lower(), rep) for n, rep in pairs]
        # built IgnoreCase, run CaseSensitive == a CaseSensitive replacer over the lowered needles (ASCII: lengths agree)
        assert R.run(R.set_case_sensitivity(0, R.build(1, pairs)), hay) == oracle.Replacer(lowered, cs=0, lower=lower_dense).run(hay)
        # built CaseSensitive over lower-case needles, run IgnoreCase == an IgnoreCase replacer
        assert R.run(R.set_case_sensitivity(1, R.build(0, lowered)), hay) == oracle.Replacer(lowered, cs=1, lower=lower_dense).run(hay)


def test_device_resident_run(am, oracle):
    import torch
    R = am.replacer
    r = R.build(0, [("tshirt", "banana"), ("shirt", "pear")])
    text = b"sweatshirts and shirttshirts " * 1000
    dev = torch.frombuffer(bytearray(text), dtype=torch.uint8).cuda()
    ptr, n = R.run_dev(r, dev.data_ptr(), len(text))
    out = torch.empty(n, dtype=torch.uint8, device="cuda")
    import ctypes
    cudart = ctypes.CDLL("libcudart.so.12")                       # (already loaded by torch)
    cudart.cudaMemcpy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
    assert cudart.cudaMemcpy(out.data_ptr(), ptr, n, 3) == 0      # cudaMemcpyDeviceToDevice
    R.free_dev(ptr)
    assert bytes(out.cpu().numpy()) == b"sweabananas and pearbananas " * 1000 == R.run(r, text)
    ptr, n = R.run_dev(R.build(0, [("zzz", "y")]), dev.data_ptr(), len(text))    # nothing to replace: a copy of the input comes back
    assert n == len(text) and ptr != dev.data_ptr()
    R.free_dev(ptr)


def test_properties_vs_oracle(am, oracle, lower_dense):
    """AhoCorasickSpec.hs:137-163 generators; every result compared with the oracle's Replacer."""
    R = am.replacer
    rng = np.random.default_rng(17)
    alpha = "abAB"
    def gen_hay(maxlen):
        return "".join(("İ" if rng.random() < 0.03 else ("ß" if rng.random() < 0.02 else alpha[int(rng.integers(0, 4))])) for _ in range(int(rng.integers(0, maxlen))))
    def gen_pairs():
        return [("".join(alpha[int(i)] for i in rng.integers(0, 4, size=int(rng.integers(1, 4)))),
                 "".join(alpha[int(i)] for i in rng.integers(0, 4, size=int(rng.integers(0, 4))))) for _ in range(int(rng.integers(0, 5)))]
    for it in range(120):
        hay, p1, p2 = gen_hay(40), gen_pairs(), gen_pairs()
        for cs in (0, 1):
            o12 = oracle.Replacer(p1 + p2, cs=cs, lower=lower_dense).run(hay)
            r1, r2 = R.build(cs, p1), R.build(cs, p2)
            r12 = R.compose(r1, r2)
            got = R.run(r12, hay)
            assert got == o12, (cs, p1, p2, hay)
            assert R.run(r2, R.run(r1, hay)) == got                    # compose law (:137-148)
        expected = hay
        for n, rep in p1:
            expected = expected.replace(n, rep)
        assert R.run(R.build(0, p1), hay).decode("utf-8") == expected  # == sequential Text.replace (:154-163)
    assert R.compose(R.build(0, []), R.build(1, [])) is None


def test_overlap_chains_and_empty_needle(am, oracle):
    R = am.replacer
    for pairs, hay in [([("aaa", "b")], "a" * 100001), ([("aa", "a")], "a" * 5000), ([("ab", ""), ("ba", "x")], "ab" * 3000 + "a"),
                       ([("", "-"), ("b", "")], "abcabc"), ([("abc", "abcabc")], "abc" * 2000)]:
        assert R.run(R.build(0, pairs), hay) == oracle.Replacer(pairs).run(hay), pairs


def test_config4_downscaled(am, oracle):
    """BASELINE.json config 4 down-scaled: 300 (needle, replacement) pairs, 8 MiB a-z haystack with plants;
    result bytes and pass count identical to the oracle."""
    from alfred_margaret_b200 import synth
    rng = np.random.default_rng(62)
    needles = synth.random_needles(300, 62, 4, 12)
    repls = [bytes(rng.integers(ord("A"), ord("Z") + 1, size=int(rng.integers(0, 16)), dtype=np.uint8)) for _ in needles]
    for i in range(0, 300, 20):     # cascading: some replacements contain a lower-priority needle
        repls[i] = repls[i][:3] + needles[(i + 7) % 300]
    hay = synth.fill_host(0, 8 << 20, 63)
    synth.plant_host(hay, 0, 64, needles[:64])
    pairs = list(zip(needles, repls))
    o = oracle.Replacer(pairs)
    want = o.run(hay)
    r = am.replacer.build(0, pairs)
    got = am.replacer.run(r, hay)
    assert got == want
    assert r.last_passes == o.passes


def test_config4_idempotence_at_scale(am):
    """BASELINE.json config 4 shape at 256 MiB (size-independent properties): with non-empty upper-case
    replacements no replacement can create or join a lower-case needle, so the result contains no needle and
    running the replacer again is the identity; lengths add up."""
    import numpy as np
    from alfred_margaret_b200 import synth
    rng = np.random.default_rng(64)
    needles = synth.random_needles(5000, 62, 5, 16)
    repls = [bytes(rng.integers(ord("A"), ord("Z") + 1, size=int(rng.integers(1, 25)), dtype=np.uint8)) for _ in needles]
    hay = synth.fill_host(0, 256 << 20, 63)
    synth.plant_host(hay, 0, 64, needles[:64])
    r = am.replacer.build(0, list(zip(needles, repls)))
    out = am.replacer.run(r, hay)
    assert 40 <= r.last_passes <= 5001
    s = am.searcher.build(0, needles)
    assert am.searcher.contains_any(s, out) is False
    again = am.replacer.run(r, out)
    assert again == out and r.last_passes == 1
    assert am.replacer.run_with_limit(r, len(out) - 1, hay) is None and am.replacer.run_with_limit(r, len(out) + (1 << 20), hay) == out


def test_incremental_passes_equal_full_rescans(am, oracle, monkeypatch):
    """the match list is carried from pass to pass (old matches that touch no edit are
    shifted, only the neighbourhood of each replacement is rescanned).  Same bytes and pass count as the literal
    form (AM_REPLACER_RESCAN=1: a full scan per pass) and as the oracle, on inputs built to stress the carry:
    replacements that create lower-priority needles, deletions (matches across the junction), adjacent and
    overlapping occurrences, replacements longer and shorter than the needle, needles that are substrings of others."""
    R = am.replacer
    rng = np.random.default_rng(77)
    cases = [
        ([("ab", ""), ("ba", "x"), ("aa", "b"), ("xb", "ab")], "ab" * 3000 + "a" + "ba" * 100),
        ([("abc", "c"), ("cc", "abab"), ("ab", "ba"), ("bab", ""), ("aa", "c")], "abc" * 500 + "cab" * 500 + "aabbcc" * 300),
        ([("aaa", "a"), ("aa", "bb"), ("bbb", "ab"), ("ab", "")], "a" * 4097 + "b" * 100 + "ab" * 777),
        ([("tshirt", "banana"), ("shirt", "pear"), ("banana", "tshirts"), ("pear", "shirt"), ("anas", "")], "sweatshirts and shirttshirts " * 400),
    ]
    for _ in range(6):   # random cascades over a small alphabet: replacements are themselves made of needle fragments
        frags = ["".join("abc"[int(i)] for i in rng.integers(0, 3, size=int(rng.integers(1, 4)))) for _ in range(6)]
        pairs = []
        for _ in range(int(rng.integers(3, 9))):
            nd = "".join(frags[int(i)] for i in rng.integers(0, 6, size=int(rng.integers(1, 3))))
            rp = "".join(frags[int(i)] for i in rng.integers(0, 6, size=int(rng.integers(0, 3))))
            pairs.append((nd, rp))
        hay = "".join(frags[int(i)] for i in rng.integers(0, 6, size=3000))
        cases.append((pairs, hay))
    for pairs, hay in cases:
        o = oracle.Replacer(pairs)
        want = o.run(hay)
        r = R.build(0, pairs)
        monkeypatch.delenv("AM_REPLACER_RESCAN", raising=False)
        got = R.run(r, hay)
        assert got == want, pairs
        assert r.last_passes == o.passes and r.last_rescans == 1
        monkeypatch.setenv("AM_REPLACER_RESCAN", "1")
        assert R.run(r, hay) == want
        assert r.last_passes == o.passes and r.last_rescans == o.passes
    monkeypatch.delenv("AM_REPLACER_RESCAN", raising=False)
    # run_with_limit sees the same `replacementLength` in both forms
    pairs, hay = cases[1]
    full = oracle.Replacer(pairs).run(hay)
    r = R.build(0, pairs)
    assert R.run_with_limit(r, len(full) + 4096, hay) == oracle.Replacer(pairs).run_with_limit(hay, len(full) + 4096)
    assert R.run_with_limit(r, 10, hay) is None and oracle.Replacer(pairs).run_with_limit(hay, 10) is None


def test_periodic_text_is_one_overlap_cluster(am, oracle):
    """removeOverlap (Replacer.hs:191-198) on a text where every match overlaps its predecessor: "aa" in a long run of a's is
    ONE cluster of a million matches; the kept ones are every second.  (The cluster walk is warp-cooperative; a serial walk
    of this list took seconds.)  Also a needle that overlaps itself at distance 2 and deletions that join."""
    R = am.replacer
    n = (1 << 20) + 3
    for pairs, text in (([("aa", "b")], b"a" * n), ([("aba", "X"), ("bX", "")], b"ab" * (n // 2) + b"a"), ([("aa", ""), ("ba", "c")], b"baaa" * (n // 4))):
        want = oracle.Replacer(pairs, cs=0).run(text)
        r = R.build(0, pairs)
        assert R.run(r, text) == want
        import os
        os.environ["AM_REPLACER_RESCAN"] = "1"
        try:
            assert R.run(r, text) == want                              # the literal pass structure agrees
        finally:
            del os.environ["AM_REPLACER_RESCAN"]
