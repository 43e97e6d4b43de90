#!/usr/bin/env python
"""bench.py -- the measurement contract for the alfred-margaret B200 hot path.

Metric (BASELINE.json): haystack GB/s of the all-matches scan (`runText` collecting every match, in the
reference's callback order) on config C2: 1 000 random 4-16 byte a-z needles, CaseSensitive, a 4 GiB
synthetic haystack per GPU with one needle planted per 4 KiB.

  value     whole-job GB/s, haystack resident in HBM when the timed region starts (am_find_all_dev:
            scan kernel + radix sort of the match keys + unpack; at N > 1 plus the NCCL all-gather of
            the per-shard match counts).
  e2e       the same metric through the drop-in C-ABI call am_find_all with HOST (pinned) buffers:
            H2D of the haystack and D2H of the match list inside the timed region.
  roofline  the scan kernel alone (CUDA events around the kernel on its launch stream, recorded inside
            the library) against the measured HBM copy bandwidth; 1 algorithmic byte per haystack byte.
  cpu_baseline  the CPU oracle (a C port of the reference's algorithm and memory layout; GHC is not in
            this image) on one host core over a bounded sample of the same haystack.

`--impl reference` times that CPU port on all host cores (rank 0 only).
One JSON line on stdout (rank 0).
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
SEED_NEEDLES, SEED_HAY, SEED_PLANT = 42, 43, 44


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--bytes-per-gpu", type=float, default=4 * GIB, help="haystack bytes per GPU (C2: 4 GiB)")
    p.add_argument("--needles", type=int, default=1000)
    p.add_argument("--cpu-sample", type=float, default=256 << 20, help="bytes the 1-core CPU baseline scans")
    p.add_argument("--ref-sample", type=float, default=0, help="bytes per step of the --impl reference arm (0 = auto)")
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


def make_needles(n):
    from alfred_margaret_b200 import synth
    return synth.random_needles(n, SEED_NEEDLES)


def run_reference(args):
    """The reference's CPU implementation of the path = the C port in oracle/ (GHC absent), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    import am_oracle_py as oracle
    from alfred_margaret_b200 import synth
    threads = host_threads()
    needles = make_needles(args.needles)
    m = oracle.Machine(needles)
    # bounded sample of the C2 haystack: ~0.025 GB/s per core on this needle set => aim at ~1-2 s per step
    sample = int(args.ref_sample) or int(min(args.bytes_per_gpu, max(64 << 20, min(2 * GIB, threads * (24 << 20)))))
    hay = synth.fill_host(0, sample, SEED_HAY)
    synth.plant_host(hay, 0, SEED_PLANT, needles)
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
                         "sample": "%d MiB of the C2 haystack per step, %d overlapping shards (halo = max needle length)" % (sample >> 20, threads)},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "matches_per_step": n_matches, "matches_per_s": n_matches * args.steps / dt,
        "note": "GHC is not installed in this image: the reference cannot be built; this arm is oracle/am_oracle.c (the reference's algorithm and packed layout restated in C), which the reference itself runs single-threaded",
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import numpy as np
    import torch
    from alfred_margaret_b200 import _ffi, automaton, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU: there is no CPU fallback"
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = _ffi.lib()
    B = int(args.bytes_per_gpu)
    needles = make_needles(args.needles)
    m = automaton.AcMachine([(n, i) for i, n in enumerate(needles)], device=local)
    info = m.info()
    halo = info["halo_bytes"]
    st = torch.cuda.current_stream().cuda_stream

    # ---- this rank's shard of the logical N x B haystack, resident with its halo ----------------------
    begin = rank * B
    pre = ((halo + 15) // 16) * 16 if rank > 0 else 0
    dev = torch.empty(B + pre, dtype=torch.uint8, device="cuda")
    synth.fill_dev(dev.data_ptr(), B + pre, begin - pre, SEED_HAY, stream=st)
    synth.plant_dev(dev.data_ptr(), B + pre, begin - pre, SEED_PLANT, needles, stream=st)
    n_local = m.count_matches_dev(dev.data_ptr(), B + pre, report_begin=pre, pos_base=begin - pre, stream=st)
    cap = n_local + 1024
    out = torch.empty(2 * cap, dtype=torch.int64, device="cuda")
    counts = torch.zeros(world, dtype=torch.int64, device="cuda")
    mine = torch.zeros(1, dtype=torch.int64, device="cuda")

    def step_dev():
        n = m.find_all_dev(dev.data_ptr(), B + pre, out.data_ptr(), cap, report_begin=pre, pos_base=begin - pre, stream=st)
        if world > 1:  # the path's only exchange: per-shard match counts -> global offsets
            mine.fill_(n)
            dist.all_gather_into_tensor(counts, mine)
        return n

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

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
        ms = _ffi.C.c_float()
        L.am_profile_last_scan_ms(_ffi.C.byref(ms))
        scan_ms.append(ms.value)
    e1.record()
    barrier()
    launches = L.am_profile_kernel_launches() - launches0
    ms_total = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_step = ms_total.item() / args.steps
    n_total = int(counts.sum().item()) if world > 1 else n_local

    # ---- end to end through the drop-in C-ABI call with host buffers -----------------------------------------
    # the pinned host copy of the shard: all of it when host memory allows (it does at N <= 4 on a 62 GB box); with many
    # ranks on a small host the end-to-end input shrinks to this rank's share of the free memory and the line says so
    E = B
    try:
        import psutil
        share = psutil.virtual_memory().available // max(1, world) - (3 << 30)
        if share < B:
            E = int(max(256 << 20, min(B, share // (256 << 20) * (256 << 20))))
    except Exception:
        pass
    host = torch.empty(E, dtype=torch.uint8, pin_memory=True)
    host.copy_(dev[pre:pre + E])
    torch.cuda.synchronize()
    hbuf = host.numpy()
    hout = np.empty(cap, dtype=automaton.MATCH_DTYPE)
    hs = _ffi.U8Slice(hbuf.ctypes.data, 0, E)
    nf = _ffi.C.c_uint64()

    def step_e2e():
        _ffi.check(L.am_find_all(m.handle, hs, hout.ctypes.data, cap, _ffi.C.byref(nf)))

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
    e2e_ms = torch.tensor([max(f0.elapsed_time(f1) / args.steps, 0.0)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_ms = e2e_ms.item()
    n_e2e = int(nf.value)

    # ---- CPU baseline + parity spot check (rank 0, single GPU only) ---------------------------------------------
    cpu = None
    parity = None
    if rank == 0 and world == 1:
        import am_oracle_py as oracle
        S = int(min(args.cpu_sample, E))
        sample = hbuf[:S]
        om = oracle.Machine(needles)
        t0 = time.perf_counter()
        want = om.find_all(sample, cap=S // 512 + 4096)
        dt = time.perf_counter() - t0
        cpu = {"value": S / dt / 1e9, "unit": "GB/s", "cores": 1, "kind": "port",
               "sample": "first %d MiB of the same haystack, oracle/am_oracle.c (C port of the reference's algorithm and layout), 1 core" % (S >> 20),
               "matches_per_s": len(want) / dt}
        got = hout[: n_e2e]
        k = int(np.searchsorted(got["end_pos"], S, side="right"))
        parity = bool(k == len(want) and np.array_equal(got["end_pos"][:k].astype(np.int64), want["pos"])
                      and np.array_equal(got["needle_id"][:k].astype(np.int64), want["value"]))

    if rank == 0:
        peak, peak_src = peaks()
        scan_avg = sum(scan_ms) / len(scan_ms)
        achieved = B / (scan_avg * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                t = json.load(f)
                traffic = t.get("dram_bytes_per_launch_scaled_to", {}).get(str(B)) or t.get("dram_bytes_per_byte", 0) * B or None
        except Exception:
            pass
        line = {
            "metric": "haystack GB/s, all-matches scan (runText)", "value": world * B / (ms_step * 1e-3) / 1e9, "unit": "GB/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "C2: %d random 4-16 B a-z needles, CaseSensitive, all matches (ordered list) over a %.2f GiB synthetic a-z haystack per GPU, one needle planted per 4 KiB"
                                   % (args.needles, B / GIB),
                       "haystack_bytes_per_gpu": B, "sharding": "contiguous shards, halo %d B, NCCL all-gather of match counts" % halo if world > 1 else "single shard",
                       "l2": "inputs (%.1f GiB per GPU) are larger than L2 (126 MB); no flush needed" % (B / GIB),
                       "kernel": {1: "walk", 2: "qgram-filter"}[info["kernel_kind"]]},
            "matches_per_step": n_total, "matches_per_s": n_total / (ms_step * 1e-3),
            "e2e": {"value": world * E / (e2e_ms * 1e-3) / 1e9, "unit": "GB/s", "h2d_bytes_per_step": E, "d2h_bytes_per_step": n_e2e * 16 + 16,
                    "ms_per_step": e2e_ms, "wall_ms_per_step": wall_e2e, "api": "am_find_all (host slice in, am_match[] out)",
                    "input": "the whole shard" if E == B else "first %d MiB of the shard (host memory per rank)" % (E >> 20)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "kernel": "am::filter_kernel<EMIT>", "ms_per_launch": scan_avg, "algorithmic_bytes_per_launch": B, "peak_source": peak_src},
            "cpu_baseline": cpu,
            "clocks": clocks,
            "parity_checked_vs_oracle": parity,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _emit_only_json_on_stdout(fn, args):
    """Exactly ONE line on stdout: libraries (e.g. NCCL's version banner) print to fd 1, so everything else is
    routed to stderr while the benchmark runs and the JSON line is written to the saved descriptor at the end."""
    import builtins
    real = os.dup(1)
    os.dup2(2, 1)
    lines = []
    orig_print = builtins.print

    def capture(*a, **k):
        if k.get("file") in (None, sys.stdout):
            lines.append(" ".join(str(x) for x in a))
        else:
            orig_print(*a, **k)

    builtins.print = capture
    try:
        fn(args)
    finally:
        builtins.print = orig_print
        sys.stdout.flush()
        os.dup2(real, 1)
        os.close(real)
    for ln in lines:
        orig_print(ln, flush=True)


if __name__ == "__main__":
    a = parse_args()
    _emit_only_json_on_stdout(run_reference if a.impl == "reference" else run_ours, a)
