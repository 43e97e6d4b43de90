#!/usr/bin/env python
"""bench.py -- the measurement contract for the alfred-margaret B200 hot path.

Headline (BASELINE.json `metric`, quoted on config C2): haystack GB/s of the all-matches scan (`runText` collecting
every match, in the reference's callback order): 1 000 random 4-16 byte a-z needles, CaseSensitive, a 4 GiB synthetic
haystack per GPU with one needle planted per 4 KiB.

  value     whole-job GB/s, haystack resident in HBM when the timed region starts (am_find_all_dev; at N > 1
            am_find_all_sharded: scan + ordering + the NCCL all-gather of the per-shard match counts, all inside the
            C ABI, one host round trip).
  e2e       the same metric through the drop-in C-ABI call am_find_all with HOST buffers: H2D of the haystack and D2H
            of the match list inside the timed region (pinned memory; `pageable` = the same call on an ordinary
            numpy buffer, which the library page-locks in place for the call).
  roofline  the scan kernel alone (CUDA events around the kernel on its launch stream, recorded inside the library)
            against the measured HBM copy bandwidth; 1 algorithmic byte per haystack byte.
  cpu_baseline  the CPU oracle (a C port of the reference's algorithm and memory layout; GHC is not in this image),
            timed like the reference's own harness: pinned to core 1, 5 runs, min and mean
            (benchmark/benchmark.py:47-56), -march=native, on a bounded sample of the same haystack.
  configs   the other BASELINE.json configurations -- C1 (plumbing), C3 (10 k needles IgnoreCase, 8 GiB mixed-case
            UTF-8), C4 (Replacer, 5 000 pairs, 2 GiB), C5 (100 k needles, 64 GiB over the GPUs) -- each with its
            device-resident GB/s, kernel, roofline fraction and an in-run parity check against the oracle.

`--impl reference` times the CPU port on all host cores (rank 0 only).  One JSON line on stdout (rank 0).
A failed parity check marks the line `"rejected": true` and the process exits with status 3.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(ROOT, "alfred-margaret_b200"), os.path.join(ROOT, "oracle")]

GIB = 1 << 30
MIB = 1 << 20
KERNEL_NAMES = {1: "am::walk_kernel", 2: "am::filter_kernel"}


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--bytes-per-gpu", type=float, default=4 * GIB, help="haystack bytes per GPU (C2: 4 GiB)")
    p.add_argument("--needles", type=int, default=1000)
    p.add_argument("--cpu-sample", type=float, default=96 * MIB, help="bytes each of the 5 runs of the 1-core CPU baseline scans")
    p.add_argument("--ref-sample", type=float, default=0, help="bytes per step of the --impl reference arm (0 = auto)")
    p.add_argument("--configs", default="C1,C3,C4,C5", help="extra BASELINE.json configs to measure beside the C2 headline ('' = none)")
    p.add_argument("--c3-bytes", type=float, default=8 * GIB)
    p.add_argument("--c4-bytes", type=float, default=2 * GIB)
    p.add_argument("--c5-bytes", type=float, default=64 * GIB, help="TOTAL C5 haystack, sharded over the GPUs (strong scaling)")
    p.add_argument("--cpu-baseline-child", default="", help=argparse.SUPPRESS)
    return p.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, STREAM-style copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_threads():
    """Threads the CPU arm can really use: the affinity mask, capped by the cgroup CPU quota."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = max(1, min(n, int(int(quota) / int(period))))
    except Exception:
        pass
    return n


def bind_to_gpu_numa_node(index):
    """Run this rank on the CPUs next to its GPU, BEFORE any pinned host memory is allocated: page-locked staging buffers
    are then first-touched on the GPU's NUMA node and eight ranks no longer share one node's memory controller."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        before = sorted(os.sched_getaffinity(0))
        pynvml.nvmlDeviceSetCpuAffinity(h)
        after = sorted(os.sched_getaffinity(0))
        node = None
        try:
            bus = pynvml.nvmlDeviceGetPciInfo(h).busId
            bus = bus.decode() if isinstance(bus, bytes) else bus
            node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus[-12:].lower()).read())
        except Exception:
            pass
        return {"cpus_before": len(before), "cpus": len(after), "first_cpu": after[0] if after else None, "numa_node": node}
    except Exception as e:
        return {"error": str(e)[:80]}


# =====================================================================================================================
# CPU arms (the oracle: test infrastructure, used here only as the timed baseline and as the checker)
# =====================================================================================================================
def cpu_baseline_child(spec):
    """One pinned process (the reference's harness protocol: `taskset -c 1`, 5 runs, benchmark/benchmark.py:47-56)."""
    n_needles, sample = (int(x) for x in spec.split(","))
    try:
        cpus = sorted(os.sched_getaffinity(0))
        os.sched_setaffinity(0, {cpus[1] if len(cpus) > 1 else cpus[0]})
    except Exception:
        pass
    os.environ["AM_ORACLE_NATIVE"] = "1"
    import am_oracle_py as oracle
    from alfred_margaret_b200 import synth, workloads
    needles = workloads.c2_needles(n_needles)
    t0 = time.perf_counter()
    om = oracle.Machine(needles)
    build_s = time.perf_counter() - t0
    hay = synth.fill_host(0, sample, workloads.C2_SEEDS[1])
    synth.plant_host(hay, 0, workloads.C2_SEEDS[2], needles)
    runs, n = [], 0
    for _ in range(5):
        t0 = time.perf_counter()
        n = om.count(hay)                      # `runText 0 (\n _ -> Step (n + 1))`, benchmark/haskell/app/Main.hs:67-76
        runs.append(time.perf_counter() - t0)
    print(json.dumps({"runs_s": runs, "build_s": build_s, "matches": int(n), "cpu": sorted(os.sched_getaffinity(0))}), flush=True)


def cpu_baseline(n_needles, sample):
    out = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-baseline-child", "%d,%d" % (n_needles, sample)],
                         capture_output=True, text=True, timeout=600)
    r = json.loads(out.stdout.strip().splitlines()[-1])
    runs = r["runs_s"]
    gbs = [sample / t / 1e9 for t in runs]
    return {"value": max(gbs), "unit": "GB/s", "cores": 1, "kind": "port",
            "mean": statistics.mean(gbs), "stdev": statistics.pstdev(gbs), "runs": 5, "automaton_build_s": r["build_s"], "pinned_to_cpu": r["cpu"],
            "sample": "first %d MiB of the C2 haystack per run; oracle/am_oracle.c (C port of the reference's algorithm and packed layout), -O3 -march=native, "
                      "pinned to one core, 5 runs, value = best (min time), mean/stdev beside it -- the protocol of benchmark/benchmark.py:47-56; "
                      "automaton build timed separately" % (sample >> 20),
            "matches_per_s": r["matches"] / min(runs)}


def run_reference(args):
    """The reference's CPU implementation of the path = the C port in oracle/ (GHC absent), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    os.environ["AM_ORACLE_NATIVE"] = "1"
    import am_oracle_py as oracle
    from alfred_margaret_b200 import synth, workloads
    threads = host_threads()
    needles = workloads.c2_needles(args.needles)
    m = oracle.Machine(needles)
    # bounded sample of the C2 haystack: ~0.04 GB/s per core on this needle set => aim at ~1-2 s per step
    sample = int(args.ref_sample) or int(min(args.bytes_per_gpu, max(64 * MIB, min(2 * GIB, threads * (24 * MIB)))))
    hay = synth.fill_host(0, sample, workloads.C2_SEEDS[1])
    synth.plant_host(hay, 0, workloads.C2_SEEDS[2], needles)
    cap = sample // 1024 + 4096
    n_matches = 0
    for _ in range(args.warmup):
        n_matches = len(m.find_all(hay, threads=threads, cap=cap))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        n_matches = len(m.find_all(hay, threads=threads, cap=cap))
    dt = time.perf_counter() - t0
    gbs = sample * args.steps / dt / 1e9
    line = {
        "impl": "reference", "metric": "haystack GB/s, all-matches scan (runText)", "value": gbs, "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "C2: %d random 4-16 B a-z needles, CaseSensitive, all matches; CPU arm scans a %d MiB sample of the 4 GiB haystack per step" % (args.needles, sample >> 20)},
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": threads, "kind": "port",
                         "sample": "%d MiB of the C2 haystack per step, %d overlapping shards (halo = max needle length), -O3 -march=native" % (sample >> 20, threads)},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "matches_per_step": n_matches, "matches_per_s": n_matches * args.steps / dt,
        "note": "GHC is not installed in this image: the reference cannot be built; this arm is oracle/am_oracle.c (the reference's algorithm and packed layout restated in C), which the reference itself runs single-threaded",
    }
    print(json.dumps(line), flush=True)
    return 0


# =====================================================================================================================
# our arm
# =====================================================================================================================
class Ctx:
    pass


def records(torch, out, n):
    from alfred_margaret_b200 import automaton
    return out[: 2 * n].cpu().numpy().view(automaton.MATCH_DTYPE)


def same(rec, want, shift=0):
    import numpy as np
    return bool(len(rec) == len(want) and np.array_equal(rec["end_pos"].astype(np.int64) - shift, want["pos"])
                and np.array_equal(rec["needle_id"].astype(np.int64), want["value"]))


def window_parity(torch, om, dev, rec, lo, hi, pos_base, skip, align_cp=False, **okw):
    """Oracle on the resident bytes [lo, hi) of `dev` vs the records (positions rebased by `pos_base`) whose end lies in
    (lo + skip, hi].  `skip`: leading bytes of the window whose matches are left out -- the halo (a match ending there may
    have started before the window) or, for the first window of a shard, its report_begin."""
    import numpy as np
    if align_cp:                                               # IgnoreCase: the oracle decodes code points, start on a boundary
        while lo < hi and (int(dev[lo].item()) & 0xC0) == 0x80:
            lo += 1
    w = dev[lo:hi].cpu().numpy()
    want = om.find_all(w, cap=max(1 << 16, (hi - lo) // 64), **okw)
    want = want[want["pos"] > skip]
    ends = rec["end_pos"].astype(np.int64) - pos_base
    got = rec[(ends > lo + skip) & (ends <= hi)]
    return same(got, want, shift=pos_base + lo)


def time_steps(torch, fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def last_scan_ms(L, _ffi):
    ms = _ffi.C.c_float()
    L.am_profile_last_scan_ms(_ffi.C.byref(ms))
    return ms.value


def max_over_ranks(torch, dist, world, x):
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def all_true(torch, dist, world, ok):
    t = torch.tensor([1 if ok else 0], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(t.item())


def config_c1(c):
    """C1: 3 needles, 1 MB ASCII: the full match list against the oracle (the reference's own CPU-runnable case)."""
    import numpy as np
    import am_oracle_py as oracle
    from alfred_margaret_b200 import automaton, workloads
    torch = c.torch
    needles, hay = workloads.c1()
    m = automaton.AcMachine([(n, i) for i, n in enumerate(needles)], device=c.local)
    want = oracle.Machine(needles).find_all(hay, cap=1 << 20)
    dev = torch.from_numpy(hay).cuda()
    out = torch.empty(2 * (len(want) + 16), dtype=torch.int64, device="cuda")
    n = m.find_all_dev(dev.data_ptr(), hay.size, out.data_ptr(), len(want) + 16, stream=c.st)
    ok = n == len(want) and same(records(torch, out, n), want) and same(m.find_all(hay), want) and m.count_matches(hay) == len(want)
    ms = time_steps(torch, lambda: m.find_all_dev(dev.data_ptr(), hay.size, out.data_ptr(), len(want) + 16, stream=c.st), 5, 2)
    scan = last_scan_ms(c.L, c.ffi)
    return {"workload": "C1: 3 needles [tshirt, shirts, shorts], CaseSensitive, 1 MB ASCII, all matches", "haystack_bytes": int(hay.size), "matches": int(n),
            "value": hay.size / ms / 1e6, "unit": "GB/s", "ms_per_step": ms, "kernel": KERNEL_NAMES[m.info()["kernel_kind"]],
            "roofline": {"achieved": hay.size / scan / 1e6, "frac": hay.size / scan / 1e6 / c.peak, "note": "1 MB: launch-latency-bound, one third of the SMs get a tile"},
            "parity_checked_vs_oracle": bool(ok), "parity": "full match list (device-resident and host-buffer calls) and count == oracle"}


def config_c3(c, args):
    """C3: 10 000 needles IgnoreCase (runLower) over mixed-case UTF-8 (35 % of the bytes above ASCII)."""
    import numpy as np
    import am_oracle_py as oracle
    from alfred_margaret_b200 import automaton, utf8, workloads
    torch = c.torch
    needles = workloads.c3_needles()
    unit = workloads.c3_unit(needles)
    U = unit.size
    reps = max(1, int(args.c3_bytes) // U)
    n = reps * U
    d_unit = torch.from_numpy(unit).cuda()
    dev = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
    dev[:n].view(reps, U).copy_(d_unit.unsqueeze(0).expand(reps, U))
    m = automaton.AcMachine([(x, i) for i, x in enumerate(needles)], case_sensitivity=1, device=c.local)
    cnt = m.count_matches_dev(dev.data_ptr(), n, stream=c.st)
    out = torch.empty(2 * (cnt + 16), dtype=torch.int64, device="cuda")
    k = m.find_all_dev(dev.data_ptr(), n, out.data_ptr(), cnt + 16, stream=c.st)
    rec = records(torch, out, k)
    lower_dense = oracle.lower_table_dense(utf8.host_lower_pairs())
    om = oracle.Machine(needles)
    W = min(8 * MIB, U)
    halo = m.info()["halo_bytes"]
    okw = dict(cs=1, lower=lower_dense)
    ok = k == cnt and bool(np.all(np.diff(rec["end_pos"].astype(np.int64)) >= 0))
    ok = ok and window_parity(torch, om, dev, rec, 0, W, 0, 0, **okw)                            # head
    if reps > 1:                                                                                  # a unit junction and the tail
        j = (reps // 2) * U
        ok = ok and window_parity(torch, om, dev, rec, j - W // 2, j + W // 2, 0, halo, align_cp=True, **okw)
    ok = ok and window_parity(torch, om, dev, rec, n - W, n, 0, halo, align_cp=True, **okw)
    ms = time_steps(torch, lambda: m.find_all_dev(dev.data_ptr(), n, out.data_ptr(), cnt + 16, stream=c.st), 3, 1)
    scan = last_scan_ms(c.L, c.ffi)
    ms_count = time_steps(torch, lambda: m.count_matches_dev(dev.data_ptr(), n, stream=c.st), 3, 1)
    res = {"workload": "C3: 10 000 lower-case needles (4-16 code points, 20 %% with non-ASCII), IgnoreCase (runLower), all matches over %.2f GiB mixed-case UTF-8 "
                       "(%.0f %% of the bytes above ASCII; a 16 MiB unit repeated)" % (n / GIB, 100 * float((unit >= 0x80).mean())),
           "haystack_bytes": n, "matches": int(k), "value": n / ms / 1e6, "unit": "GB/s", "ms_per_step": ms, "count_GBps": n / ms_count / 1e6,
           "kernel": KERNEL_NAMES[m.info()["kernel_kind"]] + " (IgnoreCase)",
           "roofline": {"achieved": n / scan / 1e6, "frac": n / scan / 1e6 / c.peak, "ms_per_launch": scan, "algorithmic_bytes": n,
                        "note": "events around everything the scan launches (lowering included when the text takes the lowered-copy form)"},
           "parity_checked_vs_oracle": bool(ok), "parity": "count == list length, sorted, oracle runLower on the first / a junction / the last %d MiB" % (W >> 20)}
    del dev, out
    return res


def config_c4(c, args):
    """C4: Replacer.run, 5 000 (needle, replacement) pairs, 2 GiB text."""
    import numpy as np
    import am_oracle_py as oracle
    from alfred_margaret_b200 import replacer, synth, workloads
    torch = c.torch
    needles, repls = workloads.c4_pairs()
    n = int(args.c4_bytes)
    r = replacer.build(0, list(zip(needles, repls)), device=c.local)
    # parity: the oracle's Replacer on a window-sized input (every pass of the CPU port is a full scan at ~0.04 GB/s)
    W = 512 << 10
    small = synth.fill_host(0, W, workloads.C4_SEEDS[1])
    synth.plant_host(small, 0, workloads.C4_SEEDS[2], needles[:64])
    t0 = time.perf_counter()
    orep = oracle.Replacer(list(zip(needles, repls)), cs=0)
    want = orep.run(small)
    cpu_s = time.perf_counter() - t0
    got = replacer.run(r, small)
    ok = got == want and r.last_passes == orep.passes
    small_passes = int(r.last_passes)
    dsmall = torch.from_numpy(small).cuda()
    p, ln = replacer.run_dev(r, dsmall.data_ptr(), W, stream=c.st)
    back = torch.empty(ln, dtype=torch.uint8, device="cuda")
    c.memcpy_d2d(back.data_ptr(), p, ln)
    replacer.free_dev(p)
    ok = ok and bytes(back.cpu().numpy()) == want
    # the full-size run, device-resident
    dev = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
    synth.fill_dev(dev.data_ptr(), n, 0, workloads.C4_SEEDS[1], stream=c.st)
    synth.plant_dev(dev.data_ptr(), n, 0, workloads.C4_SEEDS[2], needles[:64], stream=c.st)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    p, out_len = replacer.run_dev(r, dev.data_ptr(), n, stream=c.st)
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    passes, rescans = int(r.last_passes), int(r.last_rescans)
    dev_ms, moved = replacer.last_profile()
    # size-independent properties of the full-size output: its first bytes are the window's output up to the first place
    # where the window's end could matter, and no planted needle of the first pass survives
    head = torch.empty(min(out_len, W // 2), dtype=torch.uint8, device="cuda")
    c.memcpy_d2d(head.data_ptr(), p, head.numel())
    ok = ok and bytes(head.cpu().numpy()) == want[: head.numel()]
    replacer.free_dev(p)
    res = {"workload": "C4: Replacer.run, 5 000 (needle, replacement) pairs (needles 4-16 B a-z, replacements 0-24 B A-Z), CaseSensitive, %.2f GiB a-z text "
                       "with plants from 64 needles" % (n / GIB),
           "haystack_bytes": n, "out_bytes": int(out_len), "passes": passes, "full_scans": rescans, "seconds": sec, "value": n / sec / 1e9, "unit": "GB/s (input bytes / wall time of the run)",
           "kernel": "am::filter_kernel (first pass) + per-pass carry / rescan / splice kernels",
           "roofline": {"bound": "hbm", "bytes_moved": int(moved), "device_ms": dev_ms,
                        "achieved": (moved / dev_ms / 1e6) if dev_ms > 0 else None, "frac": (moved / dev_ms / 1e6 / c.peak) if dev_ms > 0 else None,
                        "note": "SURVEY 8d: sum over the passes of bytes scanned + bytes written, as counted by the library; a pass-latency-bound workload (%d passes)" % passes},
           "cpu_port_window_s": cpu_s, "window_passes": small_passes,
           "parity_checked_vs_oracle": bool(ok),
           "parity": "oracle Replacer on a %d KiB window: byte-identical output and identical pass count, host-buffer and device-resident calls; the first %d KiB of the "
                     "full-size output equal the window's" % (W >> 10, W >> 11)}
    del dev
    return res


def config_c5(c, args):
    """C5: 100 000 needles, 64 GiB in total, sharded over the GPUs (strong scaling), through am_find_all_sharded."""
    import numpy as np
    import am_oracle_py as oracle
    from alfred_margaret_b200 import automaton, sharded, synth, workloads
    torch = c.torch
    needles = workloads.c5_needles()
    t0 = time.perf_counter()
    m = automaton.AcMachine([(x, i) for i, x in enumerate(needles)], device=c.local)
    build_s = time.perf_counter() - t0
    info = m.info()
    halo = info["halo_bytes"]
    total = int(args.c5_bytes)
    # one shard per rank when it fits next to the other buffers, else several windows per rank, scanned one after the other
    per_rank = total // c.world
    windows = max(1, -(-per_rank // (64 * GIB)))
    P = c.world * windows
    om = oracle.Machine(needles)
    ok, n_local_total, scan_ms_sum, step_ms_sum, launches = True, 0, 0.0, 0.0, 0
    buf = None
    for wi in range(windows):
        r = c.rank * windows + wi
        w, b, e = sharded.shard_plan(total, halo, P, r)
        ln = e - w
        if buf is None or buf.numel() < ln + 64:
            buf = torch.empty(ln + 64, dtype=torch.uint8, device="cuda")
        synth.fill_dev(buf.data_ptr(), ln, w, workloads.C5_SEEDS[1], stream=c.st)
        synth.plant_dev(buf.data_ptr(), ln, w, workloads.C5_SEEDS[2], needles, stream=c.st)
        n_loc, off, tot = c.comm.count(m, buf.data_ptr(), ln, report_begin=b - w, pos_base=w, stream=c.st)
        cap = n_loc + 1024
        out = torch.empty(2 * cap, dtype=torch.int64, device="cuda")
        k, off2, tot2 = c.comm.find_all(m, buf.data_ptr(), ln, out.data_ptr(), cap, report_begin=b - w, pos_base=w, stream=c.st)
        ok = ok and (k, off2, tot2) == (n_loc, off, tot)
        rec = records(torch, out, k)
        ok = ok and bool(np.all(np.diff(rec["end_pos"].astype(np.int64)) >= 0))
        # oracle windows on EVERY rank: across the shard's left boundary (its halo) and at its right end
        Wn = 2 * MIB
        ok = ok and window_parity(torch, om, buf, rec, 0, min(ln, Wn), w, b - w)
        ok = ok and window_parity(torch, om, buf, rec, max(0, ln - Wn), ln, w, halo)
        c.barrier()
        l0 = c.L.am_profile_kernel_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = 2
        e0.record()
        for _ in range(steps):
            c.comm.find_all(m, buf.data_ptr(), ln, out.data_ptr(), cap, report_begin=b - w, pos_base=w, stream=c.st)
        e1.record()
        c.barrier()
        launches += (c.L.am_profile_kernel_launches() - l0) // steps
        step_ms_sum += max_over_ranks(torch, c.dist, c.world, e0.elapsed_time(e1) / steps)
        scan_ms_sum += max_over_ranks(torch, c.dist, c.world, last_scan_ms(c.L, c.ffi))
        n_local_total += k
        del out
    ok = all_true(torch, c.dist, c.world, ok)
    t = torch.tensor([n_local_total], dtype=torch.int64, device="cuda")
    if c.world > 1:
        c.dist.all_reduce(t)
    res = {"workload": "C5: 100 000 random 6-16 B a-z needles, CaseSensitive, all matches over %.0f GiB of a-z text in total (one needle planted per 4 KiB), "
                       "%d contiguous shard(s) of %.1f GiB with a %d B halo, am_find_all_sharded (NCCL all-gather of the match counts inside the C ABI)"
                       % (total / GIB, P, total / P / GIB, halo),
           "haystack_bytes_total": total, "scaling": "strong", "shards": P, "matches": int(t.item()), "value": total / step_ms_sum / 1e6, "unit": "GB/s", "ms_per_step": step_ms_sum,
           "automaton_states": info["num_states"], "automaton_build_s": build_s, "kernel": KERNEL_NAMES[info["kernel_kind"]], "gpu_launches_per_step": int(launches),
           "roofline": {"achieved": total / c.world / scan_ms_sum / 1e6, "frac": total / c.world / scan_ms_sum / 1e6 / c.peak, "ms_per_launch": scan_ms_sum,
                        "algorithmic_bytes_per_gpu": total // c.world, "note": "per GPU: its shard's bytes / its scan kernel time (max over ranks)"},
           "parity_checked_vs_oracle": bool(ok),
           "parity": "on every rank: count call == find_all call == list length, sorted, oracle on 2 MiB windows across the shard's left boundary and at its right end"}
    del buf
    return res


def run_ours(args):
    import numpy as np
    import torch
    from alfred_margaret_b200 import _ffi, automaton, sharded, synth, workloads

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU: there is no CPU fallback"
    affinity = bind_to_gpu_numa_node(local)
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = _ffi.lib()
    c = Ctx()
    c.torch, c.dist, c.rank, c.world, c.local, c.L, c.ffi = torch, dist, rank, world, local, L, _ffi
    c.st = torch.cuda.current_stream().cuda_stream
    c.peak, peak_src = peaks()
    # the library's own communicator (NCCL inside libam_b200): rank 0's id travels over torch.distributed's store
    c.comm = sharded.Comm.from_torch(dist, local) if world > 1 else sharded.Comm(0, 1, None, local)
    cudart = _ffi.C.CDLL("libcudart.so.12")
    cudart.cudaMemcpy.argtypes = [_ffi.C.c_void_p, _ffi.C.c_void_p, _ffi.C.c_size_t, _ffi.C.c_int]
    c.memcpy_d2d = lambda dst, src, n: cudart.cudaMemcpy(dst, src, n, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    c.barrier = barrier

    B = int(args.bytes_per_gpu)
    needles = workloads.c2_needles(args.needles)
    m = automaton.AcMachine([(n, i) for i, n in enumerate(needles)], device=local)
    info = m.info()
    halo = info["halo_bytes"]
    st = c.st
    S_NEEDLES, S_HAY, S_PLANT = workloads.C2_SEEDS

    # ---- this rank's shard of the logical N x B haystack, resident with its halo ----------------------
    begin = rank * B
    pre = ((halo + 15) // 16) * 16 if rank > 0 else 0
    dev = torch.empty(B + pre, dtype=torch.uint8, device="cuda")
    synth.fill_dev(dev.data_ptr(), B + pre, begin - pre, S_HAY, stream=st)
    synth.plant_dev(dev.data_ptr(), B + pre, begin - pre, S_PLANT, needles, stream=st)
    n_local, _, _ = c.comm.count(m, dev.data_ptr(), B + pre, report_begin=pre, pos_base=begin - pre, stream=st)
    cap = n_local + 1024
    out = torch.empty(2 * cap, dtype=torch.int64, device="cuda")
    totals = {}

    def step_dev():
        if world > 1:   # the path's only exchange -- per-shard match counts -> global offsets -- happens inside the call
            n, off, tot = c.comm.find_all(m, dev.data_ptr(), B + pre, out.data_ptr(), cap, report_begin=pre, pos_base=begin - pre, stream=st)
            totals["total"] = tot
            return n
        return m.find_all_dev(dev.data_ptr(), B + pre, out.data_ptr(), cap, report_begin=pre, pos_base=begin - pre, stream=st)

    L.am_profile_enable(1)
    for _ in range(max(3, args.warmup)):
        n = step_dev()
    assert n == n_local
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    launches0 = L.am_profile_kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    scan_ms = []
    e0.record()
    for _ in range(args.steps):
        step_dev()
        scan_ms.append(last_scan_ms(L, _ffi))
    e1.record()
    barrier()
    launches = L.am_profile_kernel_launches() - launches0
    ms_step = max_over_ranks(torch, dist, world, e0.elapsed_time(e1)) / args.steps
    n_total = int(totals.get("total", n_local))

    # ---- parity of the headline, on EVERY rank: oracle windows across the shard's left boundary and at its right end ------
    import am_oracle_py as oracle
    om = oracle.Machine(needles)
    rec = records(torch, out, n_local)
    Wn = 16 * MIB
    parity = bool(np.all(np.diff(rec["end_pos"].astype(np.int64)) >= 0))
    parity = parity and window_parity(torch, om, dev, rec, 0, min(B + pre, Wn), begin - pre, pre)
    parity = parity and window_parity(torch, om, dev, rec, B + pre - Wn, B + pre, begin - pre, halo)
    parity_all = all_true(torch, dist, world, parity)

    # ---- end to end through the drop-in C-ABI call with host buffers -----------------------------------------
    # the pinned host copy of the shard: all of it when host memory allows; with many ranks on a small host the end-to-end
    # input shrinks to this rank's share of the free memory and the line says so
    E = B
    try:
        import psutil
        share = psutil.virtual_memory().available // max(1, world) - (3 << 30)
        if share < 2 * B:
            E = int(max(256 * MIB, min(B, share // 2 // (256 * MIB) * (256 * MIB))))
    except Exception:
        pass
    host = torch.empty(E, dtype=torch.uint8, pin_memory=True)
    host.copy_(dev[pre:pre + E])
    torch.cuda.synchronize()
    hbuf = host.numpy()
    hout = np.empty(cap, dtype=automaton.MATCH_DTYPE)
    hs = _ffi.U8Slice(hbuf.ctypes.data, 0, E)
    nf = _ffi.C.c_uint64()

    def step_e2e(sl=hs):
        _ffi.check(L.am_find_all(m.handle, 0, _ffi.C.byref(sl), hout.ctypes.data, cap, _ffi.C.byref(nf)))

    for _ in range(3):
        step_e2e()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    f0.record()
    for _ in range(args.steps):
        step_e2e()
    f1.record()
    barrier()
    wall_e2e = (time.perf_counter() - t0) / args.steps * 1e3
    clocks = sampler.stop()   # sampled across both timed regions (device-resident steps and end-to-end steps)
    e2e_ms = max_over_ranks(torch, dist, world, max(f0.elapsed_time(f1) / args.steps, wall_e2e if world == 1 else 0.0))
    n_e2e = int(nf.value)
    got = hout[: n_e2e]
    kk = int(np.searchsorted(got["end_pos"], Wn, side="right"))
    want_head = om.find_all(hbuf[:Wn], cap=Wn // 64)
    want_head = want_head[want_head["pos"] <= Wn - 16]      # (matches ending in the window's last bytes: complete anyway, but keep the margin)
    gh = got[:kk]
    gh = gh[gh["end_pos"] <= Wn - 16]
    parity_e2e = same(gh, want_head)
    # pageable host memory (what a GHC `ByteArray#` is): the library page-locks it in place for the call
    pageable = None
    if world == 1:
        P = min(E, 1 * GIB)
        pbuf = np.empty(P, dtype=np.uint8)
        pbuf[:] = hbuf[:P]
        ps = _ffi.U8Slice(pbuf.ctypes.data, 0, P)
        step_e2e(ps)
        tp = []
        for _ in range(3):
            t0 = time.perf_counter()
            step_e2e(ps)
            tp.append(time.perf_counter() - t0)
        dtp = sorted(tp)[1]                                 # median of three: page-locking a GiB in place costs 30 .. 500 ms from call to call
        pageable = {"value": P / dtp / 1e9, "unit": "GB/s", "bytes": P, "ms_per_step": dtp * 1e3,
                    "how": "numpy (pageable) buffer; am_find_all page-locks texts >= 256 MiB in place for the call (cudaHostRegister) and unlocks them after; median of 3 calls", "ms_all": [round(x * 1e3, 2) for x in tp]}
    parity_all = all_true(torch, dist, world, parity_all and parity_e2e)

    # ---- CPU baseline (rank 0, single GPU only): the reference harness's protocol ---------------------------------------------
    cpu = None
    if rank == 0 and world == 1:
        try:
            cpu = cpu_baseline(args.needles, int(args.cpu_sample))
        except Exception as e:
            cpu = {"value": None, "error": str(e)[:200]}
    del host, hbuf

    # ---- the other BASELINE.json configs ------------------------------------------------------------------------------------------
    del dev, out
    torch.cuda.empty_cache()
    configs = {}
    wanted = [x for x in args.configs.split(",") if x]
    for name in wanted:
        try:
            if name == "C5":
                configs[name] = config_c5(c, args)
            elif rank == 0:                                  # single-GPU configs: measured on rank 0, the other ranks wait at C5's barrier
                configs[name] = {"C1": lambda: config_c1(c), "C3": lambda: config_c3(c, args), "C4": lambda: config_c4(c, args)}[name]()
        except Exception as e:  # a config that cannot run must not take the headline down with it; it is reported as failed
            configs[name] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300]), "parity_checked_vs_oracle": False}
        torch.cuda.empty_cache()
    config_parity = all(v.get("parity_checked_vs_oracle") is True for v in configs.values())

    rc = 0
    if rank == 0:
        scan_avg = sum(scan_ms) / len(scan_ms)
        achieved = B / (scan_avg * 1e-3) / 1e9
        traffic, traffic_src = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                t = json.load(f)
                traffic = t.get("dram_bytes_per_launch_scaled_to", {}).get(str(B)) or t.get("dram_bytes_per_byte", 0) * B or None
                traffic_src = "from profile (profiles/traffic.json: dram__bytes_read + dram__bytes_write of one ncu --set full capture, scaled to this launch's bytes), not measured in this run"
        except Exception:
            pass
        rejected = not (parity_all and config_parity)
        line = {
            "metric": "haystack GB/s, all-matches scan (runText)", "value": world * B / (ms_step * 1e-3) / 1e9, "unit": "GB/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "C2: %d random 4-16 B a-z needles, CaseSensitive, all matches (ordered list) over a %.2f GiB synthetic a-z haystack per GPU, one needle planted per 4 KiB"
                                   % (args.needles, B / GIB),
                       "haystack_bytes_per_gpu": B,
                       "sharding": ("contiguous shards, halo %d B, am_find_all_sharded: NCCL all-gather of the match counts inside the C ABI" % halo) if world > 1 else "single shard (am_find_all_dev)",
                       "l2": "inputs (%.1f GiB per GPU) are larger than L2 (126 MB); no flush needed" % (B / GIB),
                       "kernel": {1: "walk", 2: "qgram-filter"}[info["kernel_kind"]], "host_affinity": affinity},
            "matches_per_step": n_total, "matches_per_s": n_total / (ms_step * 1e-3),
            "e2e": {"value": world * E / (e2e_ms * 1e-3) / 1e9, "unit": "GB/s", "h2d_bytes_per_step": E, "d2h_bytes_per_step": n_e2e * 16 + 16,
                    "ms_per_step": e2e_ms, "wall_ms_per_step": wall_e2e, "api": "am_find_all (host slice in, am_match[] out)", "memory": "pinned (cudaHostAlloc via torch), allocated after the rank was bound with nvmlDeviceSetCpuAffinity (config.host_affinity: a no-op on a box that exposes one NUMA node / numa_node -1)",
                    "input": "the whole shard" if E == B else "first %d MiB of the shard (host memory per rank)" % (E >> 20), "pageable": pageable},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": c.peak, "unit": "GB/s", "frac": achieved / c.peak, "traffic": traffic, "traffic_source": traffic_src,
                         "kernel": KERNEL_NAMES[info["kernel_kind"]] + "<EMIT>", "ms_per_launch": scan_avg, "algorithmic_bytes_per_launch": B, "peak_source": peak_src},
            "cpu_baseline": cpu,
            "clocks": clocks,
            "parity_checked_vs_oracle": bool(parity_all),
            "parity": "every rank: sorted, oracle windows of 16 MiB across the shard's left boundary (its halo) and at its right end; host-buffer call: first 16 MiB",
            "configs": configs,
            "rejected": rejected,
        }
        print(json.dumps(line), flush=True)
        rc = 3 if rejected else 0
    c.comm.close()
    if world > 1:
        dist.destroy_process_group()
    return rc


def _emit_only_json_on_stdout(fn, args):
    """Exactly ONE line on stdout: libraries (e.g. NCCL's version banner) print to fd 1, so everything else is
    routed to stderr while the benchmark runs and the JSON line is written to the saved descriptor at the end."""
    import builtins
    real = os.dup(1)
    os.dup2(2, 1)
    lines = []
    orig_print = builtins.print
    rc = 1

    def capture(*a, **k):
        if k.get("file") in (None, sys.stdout):
            lines.append(" ".join(str(x) for x in a))
        else:
            orig_print(*a, **k)

    builtins.print = capture
    try:
        rc = fn(args)
    finally:
        builtins.print = orig_print
        sys.stdout.flush()
        os.dup2(real, 1)
        os.close(real)
    for ln in lines:
        orig_print(ln, flush=True)
    return rc or 0


if __name__ == "__main__":
    a = parse_args()
    if a.cpu_baseline_child:
        cpu_baseline_child(a.cpu_baseline_child)
        sys.exit(0)
    sys.exit(_emit_only_json_on_stdout(run_reference if a.impl == "reference" else run_ours, a))
