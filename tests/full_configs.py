"""BASELINE.json's five configs at their FULL sizes on one B200 (measurement + size-independent parity properties).

Not part of the pytest suite (it allocates up to 18 GiB and takes a few minutes); the down-scaled twins of every
config are.  One JSON line per config on stdout; the oracle (oracle/, test infrastructure) checks windows of the
full-size inputs, properties (sortedness, count == list length, shard sums, carried == rescanned Replacer) cover
the rest.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "alfred-margaret_b200"), os.path.join(ROOT, "oracle")]
import numpy as np
import torch

import am_oracle_py as oracle
from alfred_margaret_b200 import automaton, replacer, sharded, synth, utf8

GIB = 1 << 30
st = torch.cuda.current_stream().cuda_stream
which = set(sys.argv[1:]) or {"C1", "C2", "C3", "C4", "C5"}


def timeit(f, reps=3):
    f(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)


def machine(needles, cs=0, **kw):
    return automaton.AcMachine([(n, i) for i, n in enumerate(needles)], case_sensitivity=cs, **kw)


def same(rec, want, shift=0):
    return bool(len(rec) == len(want) and np.array_equal(rec["end_pos"].astype(np.int64) - shift, want["pos"])
                and np.array_equal(rec["needle_id"].astype(np.int64), want["value"]))


def find_all_dev(m, ptr, n, **kw):
    cnt = m.count_matches_dev(ptr, n, stream=st, **kw)
    out = torch.empty(2 * (cnt + 1), dtype=torch.int64, device="cuda")
    k = m.find_all_dev(ptr, n, out.data_ptr(), cnt + 1, stream=st, **kw)
    assert k == cnt
    return out, cnt


def emit(**kw):
    print(json.dumps(kw), flush=True)


# ---- C1: 3 needles, 1 MB ASCII, full list vs the oracle -------------------------------------------------------------
if "C1" in which:
    needles = ["tshirt", "shirts", "shorts"]
    rng = np.random.default_rng(1)
    sentences = ["short tshirts ", "sweatshirts and shirtshirts ", "long shirt "]
    hay = "".join(sentences[int(i)] for i in rng.integers(0, 3, size=80000))[:1000000].encode()
    want = oracle.Machine(needles).find_all(hay, cap=1 << 20)
    m = machine(needles)
    got = m.find_all(hay)
    emit(config="C1", haystack_bytes=len(hay), matches=len(want), full_list_equals_oracle=same(got, want), kernel=m.info()["kernel_kind"])

# ---- C2: 1 000 needles, 4 GiB ---------------------------------------------------------------------------------------------
if "C2" in which:
    needles = synth.random_needles(1000, 42)
    n = 4 * GIB
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    synth.fill_dev(dev.data_ptr(), n, 0, 43, stream=st); synth.plant_dev(dev.data_ptr(), n, 0, 44, needles, stream=st)
    m = machine(needles)
    out, cnt = find_all_dev(m, dev.data_ptr(), n)
    rec = out[: 2 * cnt].cpu().numpy().view(automaton.MATCH_DTYPE)
    W = 64 << 20
    om = oracle.Machine(needles)
    head = om.find_all(dev[:W].cpu().numpy(), threads=8, cap=1 << 20)
    tail_w = om.find_all(dev[n - W:].cpu().numpy(), threads=8, cap=1 << 20)
    tail_w = tail_w[tail_w["pos"] > 16]
    ok_head = same(rec[rec["end_pos"] <= W], head)
    ok_tail = same(rec[rec["end_pos"] > n - W + 16], tail_w, shift=n - W)
    halo = m.info()["halo_bytes"]
    ssum = 0
    for r in range(8):
        w, b, e = sharded.shard_plan(n, halo, 8, r)
        ssum += m.count_matches_dev(dev.data_ptr() + w, e - w, report_begin=b - w, pos_base=w, stream=st)
    ms_c = timeit(lambda: m.count_matches_dev(dev.data_ptr(), n, stream=st))
    ms_f = timeit(lambda: m.find_all_dev(dev.data_ptr(), n, out.data_ptr(), cnt + 1, stream=st))
    miss = torch.empty(n // 4, dtype=torch.uint8, device="cuda")
    synth.fill_dev(miss.data_ptr(), n // 4, 0, 45, alphabet=b"0123456789", stream=st)
    ms_any_miss = timeit(lambda: m.contains_any_dev(miss.data_ptr(), n // 4, stream=st)) if hasattr(m, "contains_any_dev") else None
    emit(config="C2", haystack_bytes=n, matches=cnt, sorted=bool(np.all(np.diff(rec["end_pos"].astype(np.int64)) >= 0)), first_64MiB_equals_oracle=ok_head,
         last_64MiB_equals_oracle=ok_tail, sum_of_8_shards_equals_whole=bool(ssum == cnt), count_GBps=n / ms_c / 1e6, find_all_GBps=n / ms_f / 1e6,
         contains_any_all_miss_GBps=(n // 4) / ms_any_miss / 1e6 if ms_any_miss else None, kernel=m.info()["kernel_kind"])
    del dev, out, miss

# ---- C3: 10 000 needles IgnoreCase, 8 GiB mixed-case UTF-8 -----------------------------------------------------------------
if "C3" in which:
    rng = np.random.default_rng(52)
    ascii_l = "abcdefghijklmnopqrstuvwxyz"
    extra = "éößåяωǳⱥ"
    nset = set()
    while len(nset) < 10000:
        k = int(rng.integers(4, 17))
        pool = ascii_l + (extra * 3 if rng.random() < 0.2 else "")
        nset.add("".join(pool[int(i)] for i in rng.integers(0, len(pool), size=k)))
    needles = sorted(nset)
    # symbol table: code points (70 % ASCII letters of either case, 10 % space / punctuation, 15 % two-byte, 4 % three-byte incl.
    # length-changing ones, 1 % four-byte) and re-cased needles (planted about once per 4 KiB)
    syms, wts = [], []
    def add(chars, total):
        for c in chars:
            syms.append(c.encode("utf-8")); wts.append(total / len(chars))
    add(ascii_l + ascii_l.upper(), 0.70); add(" .,;-", 0.10); add("éÉöÖßåÅяЯωΩǳǲǱ", 0.15); add("ẞKÅⱥ€", 0.04); add("𝄞💩", 0.01)
    plants = []
    for _ in range(512):
        nd = needles[int(rng.integers(0, len(needles)))]
        plants.append("".join((c.upper() if (rng.random() < 0.5 and len(c.upper()) == 1) else c) for c in nd).encode("utf-8"))
    for p in plants:
        syms.append(p); wts.append(1.0 / 3500 / len(plants))
    wts = np.array(wts); wts /= wts.sum()
    maxlen = max(len(s_) for s_ in syms)
    tab = np.zeros((len(syms), maxlen), dtype=np.uint8)
    for i, s_ in enumerate(syms):
        tab[i, : len(s_)] = np.frombuffer(s_, dtype=np.uint8)
    d_tab = torch.from_numpy(tab).cuda()
    d_len = torch.tensor([len(s_) for s_ in syms], dtype=torch.int64, device="cuda")
    d_cdf = torch.from_numpy(np.cumsum(wts)).cuda()
    n = int(float(os.environ.get("C3_BYTES", 8 * GIB)))
    dev = torch.empty(n + 256, dtype=torch.uint8, device="cuda")
    gen = torch.Generator(device="cuda"); gen.manual_seed(53)
    filled, CH = 0, 48 << 20                                    # symbols per chunk
    t0 = time.time()
    while filled < n:
        u = torch.rand(CH, device="cuda", dtype=torch.float64, generator=gen)
        idx = torch.searchsorted(d_cdf, u).clamp_(max=len(syms) - 1)
        lens = d_len[idx]
        ends = torch.cumsum(lens, 0)
        keep = ends <= (n - filled)                             # whole symbols only
        idx, lens, ends = idx[keep], lens[keep], ends[keep]
        if idx.numel() == 0:
            dev[filled:n] = ord(" "); filled = n; break
        starts = ends - lens + filled
        for k in range(maxlen):
            msk = lens > k
            if k >= 4 and not bool(msk.any()):
                break
            dev[starts[msk] + k] = d_tab[idx[msk], k]
        filled += int(ends[-1].item())
        if idx.numel() < CH:                                    # ran into the end: pad with spaces
            dev[filled:n] = ord(" "); filled = n
    torch.cuda.synchronize(); t_gen = time.time() - t0
    m = machine(needles, cs=1)
    out, cnt = find_all_dev(m, dev.data_ptr(), n)
    rec = out[: 2 * cnt].cpu().numpy().view(automaton.MATCH_DTYPE)
    W = min(48 << 20, n)
    head_host = dev[:W].cpu().numpy()
    while (head_host[W - 1] & 0xC0) == 0x80 or head_host[W - 1] >= 0xC0:   # cut on a code point boundary
        W -= 1
    lower_dense = oracle.lower_table_dense(utf8.host_lower_pairs())
    head = oracle.Machine(needles).find_all(head_host[:W], cs=1, lower=lower_dense, cap=1 << 22)
    ok_head = same(rec[rec["end_pos"] <= W], head)
    ms_c = timeit(lambda: m.count_matches_dev(dev.data_ptr(), n, stream=st), reps=2)
    ms_f = timeit(lambda: m.find_all_dev(dev.data_ptr(), n, out.data_ptr(), cnt + 1, stream=st), reps=2)
    nonascii = float((dev[: 256 << 20] >= 0x80).float().mean().item())
    emit(config="C3", haystack_bytes=n, needles=len(needles), matches=cnt, sorted=bool(np.all(np.diff(rec["end_pos"].astype(np.int64)) >= 0)),
         first_48MiB_equals_oracle=ok_head, count_GBps=n / ms_c / 1e6, find_all_GBps=n / ms_f / 1e6, non_ascii_byte_fraction=nonascii,
         kernel=m.info()["kernel_kind"], generate_s=t_gen)
    del dev, out

# ---- C4: Replacer, 5 000 pairs, 2 GiB ------------------------------------------------------------------------------------------
if "C4" in which:
    n = 2 * GIB
    rng = np.random.default_rng(64)
    needles = synth.random_needles(5000, 62, 4, 16)
    repls = [bytes(rng.integers(ord("A"), ord("Z") + 1, size=int(rng.integers(0, 25)), dtype=np.uint8)) for _ in needles]
    hay = synth.fill_host(0, n, 63); synth.plant_host(hay, 0, 64, needles[:64])
    r = replacer.build(0, list(zip(needles, repls)))
    t0 = time.time(); out_c = replacer.run(r, hay); t_c = time.time() - t0
    passes, rescans = r.last_passes, r.last_rescans
    os.environ["AM_REPLACER_RESCAN"] = "1"
    t0 = time.time(); out_r = replacer.run(r, hay); t_r = time.time() - t0
    del os.environ["AM_REPLACER_RESCAN"]
    emit(config="C4", haystack_bytes=n, pairs=len(needles), out_bytes=len(out_c), passes=passes, full_scans_carried=rescans, full_scans_literal=r.last_rescans,
         carried_equals_literal_form=bool(out_c == out_r), seconds_carried=t_c, seconds_literal=t_r)
    del hay, out_c, out_r

# ---- C5: 100 000 needles, 64 GiB as 8 shards of 8 GiB (one GPU: the shards are scanned one after the other) -------------------------
if "C5" in which:
    needles = synth.random_needles(100000, 72, 6, 16)
    m = machine(needles)
    halo = m.info()["halo_bytes"]
    total_len, P = 64 * GIB, 8
    om = oracle.Machine(needles)
    counts, ms_sum, ok_first, ok_last = [], 0.0, None, None
    buf = torch.empty(16 * GIB + 4096, dtype=torch.uint8, device="cuda")
    for r in range(P):
        w, b, e = sharded.shard_plan(total_len, halo, P, r)
        ln = e - w
        synth.fill_dev(buf.data_ptr(), ln, w, 73, stream=st); synth.plant_dev(buf.data_ptr(), ln, w, 74, needles, stream=st)
        cnt = m.count_matches_dev(buf.data_ptr(), ln, report_begin=b - w, pos_base=w, stream=st)
        counts.append(cnt)
        ms_sum += timeit(lambda: m.count_matches_dev(buf.data_ptr(), ln, report_begin=b - w, pos_base=w, stream=st), reps=1)
        if r in (0, P - 1):
            out = torch.empty(2 * (cnt + 1), dtype=torch.int64, device="cuda")
            k = m.find_all_dev(buf.data_ptr(), ln, out.data_ptr(), cnt + 1, report_begin=b - w, pos_base=w, stream=st)
            rec = out[: 2 * k].cpu().numpy().view(automaton.MATCH_DTYPE)
            W = 32 << 20
            if r == 0:
                want = om.find_all(buf[:W].cpu().numpy(), threads=8, cap=1 << 20)
                ok_first = same(rec[rec["end_pos"] <= W], want)
            else:
                want = om.find_all(buf[ln - W: ln].cpu().numpy(), threads=8, cap=1 << 20)
                want = want[want["pos"] > 16]
                ok_last = same(rec[rec["end_pos"] > e - W + 16], want, shift=e - W)
    # shards 2 and 3 again as ONE 16 GiB window: its count must equal the sum of the two shard counts (halo rule at scale)
    w2, b2, _ = sharded.shard_plan(total_len, halo, P, 2)
    _, _, e3 = sharded.shard_plan(total_len, halo, P, 3)
    ln = e3 - w2
    synth.fill_dev(buf.data_ptr(), ln, w2, 73, stream=st); synth.plant_dev(buf.data_ptr(), ln, w2, 74, needles, stream=st)
    merged = m.count_matches_dev(buf.data_ptr(), ln, report_begin=b2 - w2, pos_base=w2, stream=st)
    emit(config="C5", haystack_bytes=total_len, needles=len(needles), shards=P, matches=int(sum(counts)), shard_counts=counts,
         two_shards_as_one_window_equal_their_sum=bool(merged == counts[2] + counts[3]), first_32MiB_equals_oracle=ok_first, last_32MiB_equals_oracle=ok_last,
         one_gpu_count_GBps=total_len / ms_sum / 1e6, kernel=m.info()["kernel_kind"])
