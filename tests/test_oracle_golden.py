"""The CPU oracle against every known-answer vector the reference holds for the hot path
(vectors transcribed in tests/golden/reference_vectors.json)."""
import numpy as np
import pytest

from helpers import as_pairs, naive_find_all, needles_haystack


def test_utf8_encodings(golden):
    for v in golden["utf8_encodings"]:
        assert list(v["text"].encode("utf-8")) == v["bytes"], v["src"]


def test_count(golden, oracle, lower_dense):
    for v in golden["count"]:
        m = oracle.Machine(v["needles"])
        assert m.count(v["haystack"], cs=v["cs"], lower=lower_dense) == v["expected"], v["src"]


def test_find_all(golden, oracle):
    for v in golden["find_all"]:
        m = oracle.Machine(v["needles"])
        got = [(p, v["needles"][i]) for p, i in as_pairs(m.find_all(v["haystack"]))]
        assert got == [tuple(x) for x in v["expected"]], v["src"]


def test_example_file(golden, oracle):
    v = golden["example_file"]
    m = oracle.Machine(v["needles"])
    assert m.count(v["haystack"]) == v["expected_count"]
    assert [p for p, _ in as_pairs(m.find_all(v["haystack"]))] == v["expected_end_positions"]


def test_contains_any(golden, oracle, lower_dense):
    for v in golden["contains_any"]:
        m = oracle.Machine(v["needles"])
        assert m.contains_any(v["haystack"], cs=v["cs"], lower=lower_dense) is v["expected"], v["src"]


def test_replacer(golden, oracle, lower_dense):
    for v in golden["replacer"]:
        r = oracle.Replacer([tuple(p) for p in v["pairs"]], cs=v["cs"], lower=lower_dense)
        assert r.run(v["haystack"]).decode("utf-8") == v["expected"], v["src"]


def test_skip_code_points_backwards(golden, oracle):
    for v in golden["skip_code_points_backwards"]:
        got = oracle.skip_code_points_backwards(v["text"], v["index"], v["n"])
        assert got == (-1 if v["expected"] == "error" else v["expected"]), v["src"]


def test_lower_code_point_pins(golden, oracle, lower_dense):
    for v in golden["lower_code_point"]:
        assert int(oracle.lib().amo_lower_code_point(lower_dense.ctypes.data, v["from"])) == v["to"], v["note"]
    # Utf8Spec.hs:34-36 (idempotent on the BMP) and :38-43 (ASCII)
    bmp = np.arange(0x10000, dtype=np.uint32)
    once = lower_dense[bmp]
    assert np.array_equal(lower_dense[once], once)
    for c in range(128):
        assert int(oracle.lib().amo_lower_code_point(lower_dense.ctypes.data, c)) == ord(chr(c).lower())


def test_derived_quirks(golden, oracle):
    for v in golden["derived_quirks"]:
        m = oracle.Machine(v["needles"])
        assert as_pairs(m.find_all(v["haystack"])) == [tuple(x) for x in v["expected"]], v["src"]


def test_packed_layout(oracle):
    """Automaton.hs:75-94, :166-172, :190-192, :301-306: the packed arrays of the README automaton."""
    m = oracle.Machine(["tshirt", "shirts", "shorts"])
    tr, off, ra = m.packed()
    assert m.num_states == 17 and tr.size == 16 + 17 and off.size == 18   # C1 = 17 states, 33 entries
    root = tr[off[0]:off[1]]
    cps = [int(t & 0x1FFFFF) for t in root[:-1]]
    assert cps == sorted(cps, reverse=True) == [ord("t"), ord("s")]           # descending code points
    assert int(root[-1] & 0x1FFFFF) == 0 and int(root[-1]) & 0x200000          # wildcard last, falls back to 0
    assert int(root[-1] >> 32) == 0
    for c in range(128):
        wild = bool(int(ra[c]) & 0x200000)
        assert wild == (chr(c) not in "ts")
    assert int(ra[ord("t")] >> 32) == 1                                       # first allocated state


def test_contains_all_properties(oracle, lower_dense):
    """AhoCorasickSpec.hs:196-218."""
    rng = np.random.default_rng(7)
    assert oracle.Machine([""]).contains_all("whatever") is False
    assert oracle.Machine([""]).contains_all("") is False
    for _ in range(200):
        needles, hay = needles_haystack(rng)
        needles = [n for n in needles if n]
        hb = hay.encode("utf-8")
        m = oracle.Machine(needles)
        assert m.contains_all(hb) == all(n.encode("utf-8") in hb for n in needles)
        ln = [oracle.lower_utf8(n, lower_dense) for n in needles]
        lh = oracle.lower_utf8(hb, lower_dense)
        mi = oracle.Machine(ln)
        assert mi.contains_all(hb, cs=1, lower=lower_dense) == all(n in lh for n in ln)


def test_random_vs_naive(oracle, lower_dense):
    """Differential: oracle vs brute force on the reference's fragment-pool generator."""
    rng = np.random.default_rng(11)
    for it in range(400):
        needles, hay = needles_haystack(rng)
        hb = hay.encode("utf-8")
        m = oracle.Machine(needles)
        assert as_pairs(m.find_all(hb)) == naive_find_all(needles, hb)
        # IgnoreCase == CaseSensitive search of lowered needles in the lowered haystack when no
        # code point changes its byte length; counts agree in every case
        ln = [oracle.lower_utf8(n, lower_dense) for n in needles]
        lh = oracle.lower_utf8(hb, lower_dense)
        mi = oracle.Machine(ln)
        assert mi.count(hb, cs=1, lower=lower_dense) == len(naive_find_all(ln, lh))
        # slices with a non-zero offset (TestInstances.hs:26-33)
        pad = bytes(rng.integers(0, 256, size=int(rng.integers(1, 9)), dtype=np.uint8))
        buf = np.frombuffer(pad + hb + pad, dtype=np.uint8)
        assert as_pairs(m.find_all((buf, len(pad), len(hb)))) == naive_find_all(needles, hb)


def test_replacer_properties(oracle, lower_dense):
    """AhoCorasickSpec.hs:137-163: compose law, identity, CaseSensitive == sequential Text.replace."""
    rng = np.random.default_rng(5)
    alpha = "abAB"
    def gen_hay():
        return "".join(("İ" if rng.random() < 0.03 else alpha[int(rng.integers(0, 4))]) for _ in range(int(rng.integers(0, 10))))
    def gen_pairs():
        return [("".join(alpha[int(i)] for i in rng.integers(0, 4, size=int(rng.integers(1, 4)))),
                 "".join(alpha[int(i)] for i in rng.integers(0, 4, size=int(rng.integers(0, 4))))) for _ in range(int(rng.integers(0, 4)))]
    for _ in range(500):
        hay, p1, p2 = gen_hay(), gen_pairs(), gen_pairs()
        for cs in (0, 1):
            r1, r2, r12 = (oracle.Replacer(p, cs=cs, lower=lower_dense) for p in (p1, p2, p1 + p2))
            assert r2.run(r1.run(hay)) == r12.run(hay)
            assert oracle.Replacer([], cs=cs, lower=lower_dense).run(hay) == hay.encode("utf-8")
        expected = hay
        for n, r in p1:
            expected = expected.replace(n, r)
        assert oracle.Replacer(p1).run(hay).decode("utf-8") == expected


def test_run_with_limit(oracle):
    r = oracle.Replacer([("a", "bbbb")])
    assert r.run_with_limit("aa", 8) == b"bbbbbbbb"
    assert r.run_with_limit("aa", 7) is None      # Replacer.hs:240
    # replacementLength is taken BEFORE removeOverlap (Replacer.hs:240): "aaa" in "aaaa" matches twice
    r = oracle.Replacer([("aaa", "xxxxx")])
    assert r.run_with_limit("aaaa", 6) is None and r.run_with_limit("aaaa", 8) == b"xxxxxa"


def test_parallel_count(oracle, lower_dense):
    rng = np.random.default_rng(3)
    from alfred_margaret_b200 import synth
    needles = synth.random_needles(200, 42, 2, 6, b"abc")
    hay = synth.fill_host(0, 1 << 18, 43, b"abc")
    m = oracle.Machine(needles)
    assert m.count(hay, threads=7) == m.count(hay)
    hay2 = "".join(rng.choice(list("aAbBäÄßẞK"), size=50000)).encode("utf-8")
    n2 = [oracle.lower_utf8(x, lower_dense) for x in ["ab", "äß", "k", "bbä"]]
    m2 = oracle.Machine(n2)
    assert m2.count(hay2, cs=1, lower=lower_dense, threads=5) == m2.count(hay2, cs=1, lower=lower_dense)
