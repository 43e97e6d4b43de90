"""One or two launches of a scan for ncu (development aid)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "alfred-margaret_b200")]
import torch
from alfred_margaret_b200 import automaton, synth
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 30
mode = sys.argv[2] if len(sys.argv) > 2 else "count"
kind = int(sys.argv[3]) if len(sys.argv) > 3 else 0
nn = int(sys.argv[4]) if len(sys.argv) > 4 else 1000
needles = synth.random_needles(nn, 42)
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
synth.fill_dev(dev.data_ptr(), n, 0, 43); synth.plant_dev(dev.data_ptr(), n, 0, 44, needles)
m = automaton.AcMachine([(x, i) for i, x in enumerate(needles)], force_kernel=kind)
for _ in range(2):
    if mode == "count":
        print(m.count_matches_dev(dev.data_ptr(), n))
    else:
        cnt = m.count_matches_dev(dev.data_ptr(), n)
        out = torch.empty(2 * (cnt + 16), dtype=torch.int64, device="cuda")
        print(m.find_all_dev(dev.data_ptr(), n, out.data_ptr(), cnt + 16))
