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
    with pytest.raises(ValueError):
        A.run_lower(0, lambda a, _: A.Step(a + 1), m, "x")


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


def test_random_ignore_case(am, oracle, lower_dense):
    rng = np.random.default_rng(200)
    for it in range(150):
        needles, hay = needles_haystack(rng, big=60)
        hb = hay.encode("utf-8")
        ln = [am.utf8.lower_utf8(n) for n in needles]
        m = machine(am, ln, cs=1)
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
