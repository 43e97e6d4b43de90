"""End-to-end timing of the host-buffer entry points on C2 (development aid): pinned 4 GiB haystack in host memory."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "alfred-margaret_b200")]
import numpy as np, torch
from alfred_margaret_b200 import automaton, synth
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 4 << 30
needles = synth.random_needles(1000, 42)
m = automaton.AcMachine([(x, i) for i, x in enumerate(needles)])
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
def timeit(f, reps=3):
    f(); ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); r = f(); ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3, r
for label, alphabet, plant in (("C2 text (matches everywhere)", synth.AZ, True), ("all-miss text (digits)", b"0123456789", False)):
    synth.fill_dev(dev.data_ptr(), n, 0, 43, alphabet=alphabet)
    if plant: synth.plant_dev(dev.data_ptr(), n, 0, 44, needles)
    host.copy_(dev); torch.cuda.synchronize()
    h = host.numpy()
    for name, f in (("am_count_matches", lambda: m.count_matches(h)), ("am_contains_any", lambda: m.contains_any(h))):
        ms, r = timeit(f)
        print("%-30s %-18s -> %-8s %8.2f ms  (%.1f GB/s end to end)" % (label, name, r, ms, n / ms / 1e6), flush=True)
