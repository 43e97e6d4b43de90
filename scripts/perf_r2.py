"""Round-2 kernel timings on one GPU (development tool): scan-kernel ms (CUDA events inside the library) per config.
usage: perf_r2.py [C2] [C3] [C5] [C4]   env: AM_LIB (variant build), AM_DEBUG_FLAGS (FK_DEBUG builds), AM_IC_ONE_PASS"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "alfred-margaret_b200")]
import numpy as np
import torch

from alfred_margaret_b200 import _ffi, automaton, replacer, synth, workloads

GIB = 1 << 30
L = _ffi.lib()
L.am_profile_enable(1)
st = torch.cuda.current_stream().cuda_stream
which = [a for a in sys.argv[1:]] or ["C2", "C3", "C5"]
tag = "dbg=%s onepass=%s lib=%s" % (os.environ.get("AM_DEBUG_FLAGS", "0"), os.environ.get("AM_IC_ONE_PASS", "1"), os.path.basename(os.environ.get("AM_LIB", "default")))


def scan_ms():
    ms = _ffi.C.c_float()
    L.am_profile_last_scan_ms(_ffi.C.byref(ms))
    return ms.value


def measure(name, m, ptr, n, case=None, reps=3):
    cnt = m.count_matches_dev(ptr, n, stream=st, case=case)
    ts = []
    for _ in range(reps):
        m.count_matches_dev(ptr, n, stream=st, case=case)
        ts.append(scan_ms())
    line = "%-34s count %8.3f ms  %7.1f GB/s  (n=%d)" % (name, min(ts), n / min(ts) / 1e6, cnt)
    if os.environ.get("AM_DEBUG_FLAGS", "0") == "0":
        out = torch.empty(2 * (cnt + 16), dtype=torch.int64, device="cuda")
        tf = []
        for _ in range(reps):
            t0 = time.perf_counter()
            m.find_all_dev(ptr, n, out.data_ptr(), cnt + 16, stream=st, case=case)
            torch.cuda.synchronize()
            tf.append((scan_ms(), (time.perf_counter() - t0) * 1e3))
        line += "   find_all kernel %8.3f ms  call %8.3f ms" % (min(x[0] for x in tf), min(x[1] for x in tf))
    print(tag, "|", line, flush=True)


if "C2" in which:
    needles = workloads.c2_needles()
    n = 4 * GIB
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    synth.fill_dev(dev.data_ptr(), n, 0, 43, stream=st); synth.plant_dev(dev.data_ptr(), n, 0, 44, needles, stream=st)
    measure("C2 1k needles CS 4 GiB", automaton.AcMachine([(x, i) for i, x in enumerate(needles)]), dev.data_ptr(), n)
    if "more" in which:
        for k in (10000, 40000):
            nd = synth.random_needles(k, 42)
            measure("C2-like %d needles CS 4 GiB" % k, automaton.AcMachine([(x, i) for i, x in enumerate(nd)]), dev.data_ptr(), n)
    del dev

if "C3" in which:
    needles = workloads.c3_needles()
    unit = workloads.c3_unit(needles)
    reps = 2 * GIB // unit.size
    n = reps * unit.size
    dev = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
    dev[:n].view(reps, unit.size).copy_(torch.from_numpy(unit).cuda().unsqueeze(0).expand(reps, unit.size))
    m = automaton.AcMachine([(x, i) for i, x in enumerate(needles)], case_sensitivity=1)
    measure("C3 10k needles IC 2 GiB utf8", m, dev.data_ptr(), n)
    measure("C3 same machine, CaseSensitive", m, dev.data_ptr(), n, case=0)
    asc = torch.empty(n, dtype=torch.uint8, device="cuda")
    synth.fill_dev(asc.data_ptr(), n, 0, 43, alphabet=synth.AZ + synth.AZ.upper() + b" .,", stream=st)
    measure("C3 needles IC, ASCII mixed case", m, asc.data_ptr(), n)
    del dev, asc

if "C5" in which:
    needles = workloads.c5_needles()
    n = 8 * GIB
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    synth.fill_dev(dev.data_ptr(), n, 0, 73, stream=st); synth.plant_dev(dev.data_ptr(), n, 0, 74, needles, stream=st)
    m = automaton.AcMachine([(x, i) for i, x in enumerate(needles)])
    measure("C5 100k needles CS 8 GiB", m, dev.data_ptr(), n)
    if "more" in which:
        for k in (20000, 400000):
            nd = synth.random_needles(k, 72, 6, 16)
            measure("C5-like %d needles CS 8 GiB" % k, automaton.AcMachine([(x, i) for i, x in enumerate(nd)]), dev.data_ptr(), n)
        nd = synth.random_needles(100000, 75, 8, 16)
        measure("100k needles 8-16 B (q = 8)", automaton.AcMachine([(x, i) for i, x in enumerate(nd)]), dev.data_ptr(), n)
    del dev

if "robust" in which:
    # the C2 automaton (and an English-word one) on texts that are NOT uniform random letters: throughput must degrade gracefully
    n = 1 * GIB
    rng = np.random.default_rng(7)
    needles = workloads.c2_needles()
    m = automaton.AcMachine([(x, i) for i, x in enumerate(needles)])
    def tile_to_dev(unit):
        reps = n // unit.size
        dev = torch.empty(reps * unit.size, dtype=torch.uint8, device="cuda")
        dev.view(reps, unit.size).copy_(torch.from_numpy(unit).cuda().unsqueeze(0).expand(reps, unit.size))
        return dev
    # (1) dense: a needle planted every 64 bytes (64x C2's match density)
    unit = synth.fill_host(0, 16 << 20, 43, synth.AZ); synth.plant_host(unit, 0, 44, needles, block=64)
    dev = tile_to_dev(unit); measure("robust: C2 needles, a plant per 64 B", m, dev.data_ptr(), dev.numel()); del dev
    # (2) skewed letter frequencies (English-like unigram distribution) + spaces
    freq = np.array([8.2,1.5,2.8,4.3,12.7,2.2,2.0,6.1,7.0,0.15,0.77,4.0,2.4,6.7,7.5,1.9,0.095,6.0,6.3,9.1,2.8,0.98,2.4,0.15,2.0,0.074, 18.0])
    alpha = np.frombuffer(synth.AZ + b" ", dtype=np.uint8)
    unit = alpha[rng.choice(27, size=16 << 20, p=freq / freq.sum())].copy()
    dev = tile_to_dev(unit); measure("robust: C2 needles, English letter frequencies", m, dev.data_ptr(), dev.numel()); del dev
    # (3) adversarial: the text is made of the needles' own 4-gram prefixes (every position passes the first level)
    pre = np.frombuffer(b"".join(x[:4] for x in needles), dtype=np.uint8)
    unit = np.tile(pre, (16 << 20) // pre.size + 1)[: 16 << 20].copy()
    dev = tile_to_dev(unit); measure("robust: text = the needles' 4-gram prefixes", m, dev.data_ptr(), dev.numel()); del dev
    # (4) adversarial: one letter repeated, against needles of that letter (survivor flood -> hand-over to the walk kernel)
    ma = automaton.AcMachine([(b"a" * k, k) for k in range(4, 12)])
    dev = torch.full((n,), ord("a"), dtype=torch.uint8, device="cuda")
    measure("robust: 'aaaa..' vs needles a^4..a^11 (count only)", ma, dev.data_ptr(), n // 16)
    del dev

if "C4" in which:
    needles, repls = workloads.c4_pairs()
    n = 2 * GIB
    dev = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
    synth.fill_dev(dev.data_ptr(), n, 0, 63, stream=st); synth.plant_dev(dev.data_ptr(), n, 0, 64, needles[:64], stream=st)
    r = replacer.build(0, list(zip(needles, repls)))
    for _ in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        p, ln = replacer.run_dev(r, dev.data_ptr(), n, stream=st)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        replacer.free_dev(p)
        ms, moved = replacer.last_profile()
        print(tag, "| C4 replacer 2 GiB: %.3f s  passes %d  full scans %d  out %d  device ms %.1f  bytes moved %.2f GB" % (dt, r.last_passes, r.last_rescans, ln, ms, moved / 1e9), flush=True)
