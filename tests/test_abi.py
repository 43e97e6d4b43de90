"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/am_b200.h declares, its host-side helpers match the reference's vectors, and compute
entry points fail loudly (AM_E_NODEVICE) instead of falling back to the CPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "am_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(am_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from alfred_margaret_b200 import _ffi
    L = _ffi.lib()
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), "libam_b200.so does not export %s" % n
    assert sorted(_ffi.SYMBOLS) == names, "python binding and header disagree"
    assert L.am_abi_version() == 2


def test_struct_layouts_match_the_header():
    from alfred_margaret_b200 import _ffi
    assert C.sizeof(_ffi.U8Slice) == 24 and C.sizeof(_ffi.Match) == 16 and C.sizeof(_ffi.DevText) == 32
    assert C.sizeof(_ffi.Options) == 56 and C.sizeof(_ffi.LowerPair) == 8


def test_no_cpu_fallback_without_device():
    from alfred_margaret_b200 import _ffi, automaton
    if _ffi.lib().am_device_count() > 0:
        pytest.skip("a device is present")
    with pytest.raises(_ffi.NoDeviceError):
        automaton.AcMachine([("abc", 0)])
    m = automaton.AcMachine([("abc", 0)], device=-2)   # host image only
    for call in (lambda: m.count_matches("abc"), lambda: m.contains_any("abc"), lambda: m.find_all("abc")):
        with pytest.raises(_ffi.NoDeviceError):
            call()


def test_host_image_introspection(oracle):
    from alfred_margaret_b200 import automaton
    m = automaton.AcMachine([(n, i) for i, n in enumerate(["tshirt", "shirts", "shorts"])], device=-2)
    info = m.info()
    assert info == {"num_states": 17, "max_needle_bytes": 6, "halo_bytes": 5, "kernel_kind": 2}
    # byte-level states == code-point-level states for ASCII; more for multi-byte needles
    needles = ["groß", "öffnung", "tür", ""]
    mi = automaton.AcMachine([(n, ()) for n in needles], case_sensitivity=1, device=-2)
    assert mi.info()["kernel_kind"] == 1                      # IgnoreCase / empty needle -> general walk kernel
    # + the IgnoreCase variant "groẞ" (ẞ U+1E9E lowers to ß with another UTF-8 length): 3 more bytes after "gro"
    assert mi.info()["num_states"] == 1 + sum(len(n.encode()) for n in needles) + len("ẞ".encode())
    # needles that are not lower case can never match runLower and are not inserted
    md = automaton.AcMachine([("ABC", 0), ("\u212a", 1), ("ab", 2)], case_sensitivity=1, device=-2)
    assert md.info()["num_states"] == 3
    assert oracle.Machine(needles).num_states == 1 + sum(len(n) for n in needles)


def test_bad_arguments():
    from alfred_margaret_b200 import _ffi
    L = _ffi.lib()
    h = C.c_void_p()
    assert L.am_automaton_build(None, 3, None, None, C.byref(h)) == _ffi.AM_E_BADARG
    opts = _ffi.Options(-2, 0, (C.c_uint64 * 6)())
    assert L.am_automaton_build(None, 0, None, C.byref(opts), C.byref(h)) == _ffi.AM_OK   # no toLower table: a CaseSensitive-only handle
    assert L.am_automaton_prepare(h, 7) == _ffi.AM_E_BADARG
    assert L.am_automaton_prepare(h, 1) == _ffi.AM_E_BADARG                               # IgnoreCase without table
    assert b"toLower" in L.am_last_error()
    buf = C.create_string_buffer(8)
    assert L.am_last_error_copy(buf, 8) > 8 and buf.value == L.am_last_error()[:7]        # truncated, NUL-terminated
    assert L.am_automaton_prepare(h, 0) == _ffi.AM_OK
    L.am_automaton_free(h)
    n = C.c_uint64()
    sl = _ffi.U8Slice(0, 0, 0)
    assert L.am_count_matches(None, 0, C.byref(sl), C.byref(n)) == _ffi.AM_E_BADARG
    assert L.am_count_matches(None, 0, None, C.byref(n)) == _ffi.AM_E_BADARG
    c = C.c_void_p()
    assert L.am_comm_init(3, 2, None, -1, C.byref(c)) == _ffi.AM_E_BADARG                 # rank >= nranks


def test_lower_utf8_and_pins(golden, oracle, lower_dense):
    from alfred_margaret_b200 import utf8
    for v in golden["lower_code_point"]:
        assert utf8.lower_code_point(v["from"]) == v["to"], v["note"]
        assert utf8.lower_utf8(chr(v["from"])) == chr(v["to"]).encode("utf-8")
    s = "GROẞFRÄSMASCHINENÖFFNUNGSTÜR İK Ⱥ 𝄞💩 ǲ"
    assert utf8.lower_utf8(s) == oracle.lower_utf8(s, lower_dense)
    assert utf8.lower_utf8("") == b""
    rng = np.random.default_rng(1)
    cps = [int(c) for c in rng.integers(1, 0x11000, size=3000) if not 0xD800 <= c <= 0xDFFF]
    t = "".join(map(chr, cps))
    assert utf8.lower_utf8(t) == oracle.lower_utf8(t, lower_dense)


def test_skip_code_points_backwards(golden):
    from alfred_margaret_b200 import utf8
    for v in golden["skip_code_points_backwards"]:
        if v["expected"] == "error":
            with pytest.raises(ValueError):
                utf8.skip_code_points_backwards(v["text"], v["index"], v["n"])
        else:
            assert utf8.skip_code_points_backwards(v["text"], v["index"], v["n"]) == v["expected"], v["src"]
    # a Text slice with off != 0
    t = utf8.Text(b"xx" + "aİẞ💩ẞİa".encode("utf-8") + b"yy", 2, 16)
    assert utf8.skip_code_points_backwards(t, 15, 3) == 6


def test_shard_plan():
    from alfred_margaret_b200 import _ffi
    L = _ffi.lib()
    for total in (0, 1, 15, 16, 1000, (1 << 36) + 12345):
        for n in (1, 2, 4, 8, 7):
            prev_end = 0
            for r in range(n):
                w, b, e = C.c_uint64(), C.c_uint64(), C.c_uint64()
                assert L.am_shard_plan(total, 15, n, r, C.byref(w), C.byref(b), C.byref(e)) == 0
                assert b.value == prev_end and e.value >= b.value and w.value == max(0, b.value - 15)
                assert b.value % 16 == 0
                prev_end = e.value
            assert prev_end == total
    w = C.c_uint64()
    assert L.am_shard_plan(10, 1, 0, 0, C.byref(w), C.byref(w), C.byref(w)) == _ffi.AM_E_BADARG


def test_synth_host_generator():
    from alfred_margaret_b200 import synth
    a = synth.fill_host(0, 10000, 43)
    assert a.size == 10000 and set(a.tolist()) <= set(synth.AZ)
    assert np.array_equal(synth.fill_host(1234, 777, 43), a[1234:2011])       # pure function of the absolute index
    counts = np.bincount(a, minlength=128)[97:123]
    assert counts.min() > 250                                                   # roughly uniform
    needles = synth.random_needles(50, 42)
    assert len(set(needles)) == 50 and all(4 <= len(n) <= 16 for n in needles)
    full = synth.fill_host(0, 40000, 43); synth.plant_host(full, 0, 44, needles)
    part = synth.fill_host(8000, 9000, 43); synth.plant_host(part, 8000, 44, needles)
    assert np.array_equal(part, full[8000:17000])                              # shard-consistent planting
    planted = sum(full.tobytes().count(n) for n in needles)
    assert planted >= 9


def test_benchmark_file_format(golden, tmp_path):
    """The reference harness's file format (benchmark/haskell/app/Main.hs:26-40): needles, blank line, haystack."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ac_bench", os.path.join(ROOT, "alfred-margaret_b200", "tools", "ac_bench.py"))
    ac = importlib.util.module_from_spec(spec); spec.loader.exec_module(ac)
    v = golden["example_file"]
    p = tmp_path / "example.txt"
    p.write_bytes(("\n".join(v["needles"]) + "\n\n" + v["haystack"]).encode("utf-8"))
    needles, hay = ac.read_needle_haystack_file(str(p))
    assert needles == [n.encode() for n in v["needles"]] and hay == v["haystack"].encode("utf-8")
    p.write_bytes(b"a\nb")
    assert ac.read_needle_haystack_file(str(p)) == ([b"a", b"b"], b"")


def test_filter_has_no_false_negatives(oracle):
    """Host model of the fast path's filter (am_debug_host_filter: the same cell / bucket functions the kernel uses, on the
    host image): every position where the oracle finds a needle START must pass both levels -- for the stride-2 cells at
    either parity of the device address, for q < 4, for the exact second level and for the closed-4-gram / 5-gram
    bitmaps of large needle sets, and for needle sets whose prefixes share rows.  The filter may only add candidates."""
    from alfred_margaret_b200 import automaton, synth
    rng = np.random.default_rng(5)
    cases = [
        ("C2-like, exact second level", synth.random_needles(1000, 42), synth.AZ),
        ("large set, bitmap second level", synth.random_needles(6000, 43, 4, 12), synth.AZ),
        ("q = 3", synth.random_needles(300, 44, 3, 8, b"abcdef"), b"abcdef"),
        ("q = 2", synth.random_needles(40, 45, 2, 6, b"abc"), b"abc"),
        ("shared prefixes, duplicates", [b"abca", b"abcab", b"abcabc", b"bcab", b"abca", b"cabcabcab", b"aaaa", b"aaaab"], b"abc"),
        ("bytes with equal low 5 bits, UTF-8", ["aAb!", "!bAa", "Aa!b1", "åbcå", "💩ab", "ab💩", "ßßab"], "aA!b1åß💩"),
    ]
    cases.append(("q = 6: Bloom second level", synth.random_needles(5000, 47, 6, 12), synth.AZ))
    cases.append(("q = 8", synth.random_needles(4000, 48, 8, 16), synth.AZ))
    cases.append(("q = 6, C5-like density", synth.random_needles(30000, 49, 6, 16), synth.AZ))
    cases.append(("IgnoreCase: folded cells", synth.random_needles(1000, 46) + [b"a@b`", b"[x]{y}", b"  ab", b"12ab"], "IC"))
    cases.append(("IgnoreCase: folded cells, q = 6", synth.random_needles(3000, 50, 6, 12), "IC"))
    for name, needles, alpha in cases:
        nb = [n if isinstance(n, bytes) else n.encode("utf-8") for n in needles]
        if alpha == "IC":
            # the model is handed the ORIGINAL text; both filter levels see it folded (every ASCII byte | 0x20), never lowered
            raw = synth.fill_host(0, 1 << 18, 11, synth.AZ + synth.AZ.upper() + b" @`[]{}12")
            synth.plant_host(raw, 0, 12, [x.upper() for x in nb], block=512)
            m = automaton.AcMachine([(n, i) for i, n in enumerate(nb)], case_sensitivity=1, device=-2, force_kernel=2)
            want = oracle.Machine(nb).find_all(raw, cs=1, cap=1 << 22)   # (ASCII text: no table needed)
            assert len(want) > 100 and m.info()["kernel_kind"] == 2
            starts = np.unique(want["pos"] - np.array([len(nb[v]) for v in want["value"]], dtype=np.int64))
            for align in (0, 1):
                flags = m.host_filter_flags(raw, align)
                assert starts[(flags[starts] & 7) != 7].size == 0, (name, align)
                assert float((flags & 1).mean()) < 0.05
            continue
        if isinstance(alpha, bytes):
            hay = synth.fill_host(0, 1 << 18, 9, alpha)
            synth.plant_host(hay, 0, 10, nb, block=512)
        else:
            hay = np.frombuffer("".join(alpha[int(i)] for i in rng.integers(0, len(alpha), size=60000)).encode("utf-8"), dtype=np.uint8).copy()
        m = automaton.AcMachine([(n, i) for i, n in enumerate(nb)], device=-2)
        assert m.info()["kernel_kind"] == 2, name
        want = oracle.Machine(nb).find_all(hay, cap=1 << 22)
        assert len(want) > 100, name
        if name.startswith("q = 6, C5"):                        # the long q-gram keeps a large set selective (a 4-gram bitmap would pass ~7 %)
            f0 = m.host_filter_flags(hay, 0)
            assert float((f0 & 1).mean()) < 0.08 and float(((f0 & 3) == 3).mean()) < 0.01
        starts = np.unique(want["pos"] - np.array([len(nb[v]) for v in want["value"]], dtype=np.int64))
        for align in (0, 1):
            flags = m.host_filter_flags(hay, align)
            missed = starts[(flags[starts] & 7) != 7]
            assert missed.size == 0, (name, align, missed[:5])
            rate1, rate2 = float((flags & 1).mean()), float(((flags & 3) == 3).mean())
            assert rate2 <= rate1 <= 1.0
            if name.startswith("C2-like"):                      # the filter also has to FILTER: level 1 < 2 %, both levels < 0.4 %
                assert rate1 < 0.02 and rate2 < 0.004, (rate1, rate2)   # (0.2 % of the positions start a planted needle)


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """include/am_b200.h is the drop-in boundary: it must compile as plain C (no C++, no CUDA or torch types) and a C
    program must be able to link the library and drive it -- here on a host image (device = -2), without a GPU."""
    import os, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib_dir = os.path.join(root, "alfred-margaret_b200", "lib")
    src = tmp_path / "abi.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "am_b200.h"
int main(void) {
  const char *n[3] = {"tshirt", "shirts", "shorts"};
  am_u8slice needles[3];
  for (int i = 0; i < 3; i++) { needles[i].ptr = (const uint8_t *)n[i]; needles[i].off = 0; needles[i].len = (int64_t)strlen(n[i]); }
  am_options opts; memset(&opts, 0, sizeof opts); opts.device = -2;            /* host image only */
  am_automaton *a = 0;
  if (am_abi_version() != 2) return 2;
  if (am_automaton_build(needles, 3, 0, &opts, &a) != AM_OK) { printf("%s\n", am_last_error()); return 3; }
  uint64_t states = 0, maxlen = 0, halo = 0; int kind = 0;
  if (am_automaton_info(a, AM_CASE_SENSITIVE, &states, &maxlen, &halo, &kind) != AM_OK) return 4;
  const char *text = "short tshirts";
  am_u8slice hay = {(const uint8_t *)text, 0, 13};
  uint8_t flags[13];
  if (am_debug_host_filter(a, AM_CASE_SENSITIVE, &hay, 0, flags) != AM_OK) return 5;
  uint64_t cnt = 0;
  int rc = am_count_matches(a, AM_CASE_SENSITIVE, &hay, &cnt);                  /* no device: must refuse, not fall back */
  printf("%llu %llu %llu %d %d %d\n", (unsigned long long)states, (unsigned long long)maxlen, (unsigned long long)halo, kind, (flags[6] & 3) == 3, rc == AM_E_NODEVICE);
  am_automaton_free(a);
  return 0;
}
''')
    exe = tmp_path / "abi"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(root, "include"), str(src), "-o", str(exe),
                           "-L", lib_dir, "-lam_b200", "-Wl,-rpath," + lib_dir])
    out = subprocess.check_output([str(exe)], text=True).split()
    # 17 states (tshirt 6 + shirts 6 + shorts 4 new + root), longest needle 6, halo 5, filter kernel; "tshirts" starts at 6
    assert out == ["17", "6", "5", "2", "1", "1"], out


def test_filter_covers_length_changing_variants(oracle, lower_dense):
    """IgnoreCase on the lowered copy: a code point whose lower case has another UTF-8 length (K U+212A -> k, Å U+212B -> å,
    ẞ -> ß, İ -> i) stays as it is in the copy, and the automaton holds the needle VARIANTS that match it.  Host model of
    the filter on such a copy (all other code points already lower case, so the copy is the text itself): every start of
    an oracle `runLower` match must pass both filter levels -- i.e. the variants reached the bitmap and the second level."""
    from alfred_margaret_b200 import automaton
    rng = np.random.default_rng(17)
    pool = list("kaist") + list("åß")
    needles = sorted({"".join(rng.choice(pool, size=int(rng.integers(4, 9)))) for _ in range(400)})
    cps = list("kaist ") * 3 + list("åß") + list("KÅẞİ")
    hay_s = "".join(rng.choice(cps, size=120000))
    hay = np.frombuffer(hay_s.encode("utf-8"), dtype=np.uint8).copy()
    nb = [n.encode("utf-8") for n in needles]
    want = oracle.Machine(nb).find_all(hay, cs=1, lower=lower_dense, cap=1 << 22)
    assert len(want) > 500
    # start of a match = len_cps code points back from its end (makeMatch IgnoreCase, Replacer.hs:271-274)
    is_lead = (hay & 0xC0) != 0x80
    lead_index = np.flatnonzero(is_lead)                       # byte offset of every code point
    cp_of_byte = np.cumsum(is_lead) - 1
    starts = set()
    with_variant = 0
    for pos, v in zip(want["pos"].tolist(), want["value"].tolist()):
        end_cp = cp_of_byte[pos - 1] + 1
        s = int(lead_index[end_cp - len(needles[v])])
        starts.add(s)
        with_variant += any(c in "KÅẞİ" for c in hay[s:pos].tobytes().decode("utf-8"))
    assert with_variant > 50                                    # matches that only the variants can produce
    m = automaton.AcMachine([(n, i) for i, n in enumerate(nb)], case_sensitivity=1, device=-2, force_kernel=2)
    assert m.info()["kernel_kind"] == 2
    starts = np.array(sorted(starts), dtype=np.int64)
    for align in (0, 1):
        flags = m.host_filter_flags(hay, align)
        missed = starts[(flags[starts] & 7) != 7]
        assert missed.size == 0, (align, missed[:5])


def test_third_level_prefix_bitmap(oracle, lower_dense):
    """q = 4 images whose second level is the three shared-memory bitmaps carry a third level in global memory: the (folded)
    prefixes of min(length, 8) bytes of every needle variant (am_build.cpp, gp_hash), tested by verify_kernel before it decodes
    anything.  Host model, C3-like: lower-case needles with non-ASCII code points, IgnoreCase, on mixed-case UTF-8 text: every
    start of an oracle `runLower` match passes all three levels, and the third level rejects most of what the second lets by."""
    from alfred_margaret_b200 import automaton, synth, workloads
    needles = workloads.c3_needles(5000)
    hay = workloads.c3_unit(needles)[: 1 << 20].copy()
    while (hay[-1] & 0xC0) == 0x80 or hay[-1] >= 0xC0: hay = hay[:-1]          # end on a code point boundary
    want = oracle.Machine(needles).find_all(hay, cs=1, lower=lower_dense, cap=1 << 22)
    assert len(want) > 200
    is_lead = (hay & 0xC0) != 0x80
    lead_index = np.flatnonzero(is_lead)
    cp_of_byte = np.cumsum(is_lead) - 1
    n_cps = [len(n.decode("utf-8")) for n in needles]
    starts = np.unique(np.array([int(lead_index[cp_of_byte[pos - 1] + 1 - n_cps[v]]) for pos, v in zip(want["pos"].tolist(), want["value"].tolist())], dtype=np.int64))
    m = automaton.AcMachine([(n, i) for i, n in enumerate(needles)], case_sensitivity=1, device=-2, force_kernel=2)
    assert m.info()["kernel_kind"] == 2
    for align in (0, 1):
        flags = m.host_filter_flags(hay, align)
        missed = starts[(flags[starts] & 7) != 7]
        assert missed.size == 0, (align, missed[:5])
        two, three = float(((flags & 3) == 3).mean()), float(((flags & 7) == 7).mean())
        assert three < 0.4 * two, (two, three)
    # CaseSensitive image of a large set: same property on plain bytes
    nb = synth.random_needles(6000, 43, 4, 12)
    h2 = synth.fill_host(0, 1 << 18, 9, synth.AZ)
    synth.plant_host(h2, 0, 10, nb, block=512)
    w2 = oracle.Machine(nb).find_all(h2, cap=1 << 22)
    s2 = np.unique(w2["pos"] - np.array([len(nb[v]) for v in w2["value"]], dtype=np.int64))
    f2 = automaton.AcMachine([(n, i) for i, n in enumerate(nb)], device=-2).host_filter_flags(h2, 0)
    assert s2[(f2[s2] & 7) != 7].size == 0 and float(((f2 & 7) == 7).mean()) < float(((f2 & 3) == 3).mean())


def test_haskell_shim_imports_agree_with_the_header():
    """GHC is not installed here, so the shim under alfred-margaret_b200/haskell/ cannot be compiled; what CAN be checked
    is that every `foreign import ccall` names a symbol the header declares, with the same number of arguments, a pointer
    wherever the C parameter is one (every struct travels by pointer: GHC's FFI cannot pass one by value) and an integer
    status as the result -- and that the modules keep the reference's export lists (Automaton.hs:32-44, Searcher.hs:14-27,
    Replacer.hs:14-27, Splitter.hs:13-22)."""
    import re
    hdr = open(os.path.join(ROOT, "include", "am_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(int|void|uint64_t|size_t|const char \*)\s*\b(am_\w+)\s*\(([^;{]*?)\)\s*;", hdr):
        params = [x.strip() for x in m.group(3).split(",") if x.strip() and x.strip() != "void"]
        protos[m.group(2)] = (m.group(1), params)
    hs_dir = os.path.join(ROOT, "alfred-margaret_b200", "haskell", "Data", "Text", "AhoCorasick")
    ffi = open(os.path.join(hs_dir, "FFI.hs")).read()
    imports = re.findall(r'foreign import ccall\s+(?:safe|unsafe)?\s*"(&?)(am_\w+)"\s*\n\s*\w+\s*::\s*([^\n]+)', ffi)
    assert len(imports) >= 12
    for addr, name, sig in imports:
        assert name in protos, name
        ret, params = protos[name]
        if addr:                                               # `&am_x_free`: a finalizer, FunPtr (Ptr T -> IO ())
            assert sig.strip().startswith("FunPtr") and len(params) == 1 and ret == "void", name
            continue
        parts = [x.strip() for x in re.split(r"->", sig)]
        args, res = parts[:-1], parts[-1]
        assert len(args) == len(params), (name, args, params)
        for h, c in zip(args, params):
            assert ("Ptr" in h) == ("*" in c), (name, h, c)     # pointers where C has pointers, scalars where it has scalars
        assert res == {"int": "IO CInt", "void": "IO ()", "size_t": "IO CSize"}[ret], (name, res)
    exports = {
        "Automaton.hs": ["AcMachine", "build", "runText", "runLower", "runWithCase", "Match", "Next"],
        "Searcher.hs": ["Searcher", "build", "buildWithValues", "containsAny", "containsAll", "setCaseSensitivity", "caseSensitivity", "needles"],
        "Replacer.hs": ["Replacer", "build", "compose", "mapReplacement", "run", "runWithLimit", "setCaseSensitivity"],
        "Splitter.hs": ["Splitter", "build", "split", "splitIgnoreCase", "splitReverse", "splitReverseIgnoreCase"],
    }
    for fn, names in exports.items():
        src = open(os.path.join(hs_dir, fn)).read()
        head = src[: src.index(" where")]
        for n in names:
            assert re.search(r"\b%s\b" % n, head), (fn, n)
    assert not os.path.exists(os.path.join(ROOT, "alfred-margaret_b200", "haskell", "cbits"))   # no C glue: nothing is passed by value
    # the Storable instances and status constants of FFI.hs against the C side
    from alfred_margaret_b200 import _ffi
    import ctypes
    for hs_type, c_type in (("U8Slice", _ffi.U8Slice), ("AmMatch", _ffi.Match), ("AmLowerPair", _ffi.LowerPair)):
        m = re.search(r"instance Storable %s where\s*\n\s*sizeOf _ = (\d+)\s*\n\s*alignment _ = (\d+)" % hs_type, ffi)
        assert m and int(m.group(1)) == ctypes.sizeof(c_type) and int(m.group(2)) == ctypes.alignment(c_type), hs_type
    enum = dict(re.findall(r"\b(AM_\w+)\s*=\s*(-?\d+)", hdr))
    assert re.search(r"amOk = %s\b" % enum["AM_OK"], ffi) and re.search(r"amEOverflow = %s\b" % enum["AM_E_OVERFLOW"], ffi)
    assert re.search(r"caseToC CaseSensitive = %s\b" % enum["AM_CASE_SENSITIVE"], ffi) and re.search(r"caseToC IgnoreCase = %s\b" % enum["AM_IGNORE_CASE"], ffi)
