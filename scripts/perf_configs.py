"""Timing of the secondary configs (development aid): walk kernel on big needle sets and IgnoreCase."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "alfred-margaret_b200")]
import numpy as np, torch
from alfred_margaret_b200 import automaton, synth, utf8
st = torch.cuda.current_stream().cuda_stream
def timeit(f, reps=3):
    f(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
n = 1 << 30
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
for label, nn, lo, hi, cs, kind in (("C2 walk", 1000, 4, 16, 0, 1), ("10k CS auto", 10000, 4, 16, 0, 0), ("10k CS filter-forced", 10000, 4, 16, 0, 2), ("20k CS filter-forced", 20000, 4, 16, 0, 2), ("20k CS walk", 20000, 4, 16, 0, 1), ("40k CS filter-forced", 40000, 4, 16, 0, 2), ("40k CS walk", 40000, 4, 16, 0, 1),
                                    ("100k CS auto", 100000, 6, 16, 0, 0), ("100k CS filter-forced", 100000, 6, 16, 0, 2),
                                    ("1k IC auto", 1000, 4, 16, 1, 0), ("10k IC auto", 10000, 4, 16, 1, 0), ("10k IC walk", 10000, 4, 16, 1, 1)):
    needles = synth.random_needles(nn, 42, lo, hi)
    synth.fill_dev(dev.data_ptr(), n, 0, 43, alphabet=(synth.AZ + synth.AZ.upper() if cs else synth.AZ)); synth.plant_dev(dev.data_ptr(), n, 0, 44, needles)
    t0 = time.time()
    m = automaton.AcMachine([(x, i) for i, x in enumerate(needles)], case_sensitivity=cs, force_kernel=kind)
    tb = time.time() - t0
    cnt = m.count_matches_dev(dev.data_ptr(), n, stream=st)
    ms = timeit(lambda: m.count_matches_dev(dev.data_ptr(), n, stream=st))
    print("%-22s kernel=%d states=%d build %.2fs  count=%d  %.3f ms/GiB -> %.0f GB/s" % (label, m.info()["kernel_kind"], m.info()["num_states"], tb, cnt, ms, n / ms / 1e6), flush=True)
