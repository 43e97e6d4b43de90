"""Profiling driver (development aid): one scan per kernel of interest, for `ncu -k regex:... -c N`.
usage: prof_r2.py c5walk | c5 | c3 | c2"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "alfred-margaret_b200")]
import torch
from alfred_margaret_b200 import automaton, synth, workloads
GIB = 1 << 30
st = torch.cuda.current_stream().cuda_stream
what = sys.argv[1]
if what in ("c5", "c5walk"):
    needles = workloads.c5_needles()
    n = 2 * GIB
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    synth.fill_dev(dev.data_ptr(), n, 0, 73, stream=st); synth.plant_dev(dev.data_ptr(), n, 0, 74, needles, stream=st)
    m = automaton.AcMachine([(x, i) for i, x in enumerate(needles)], force_kernel=1 if what == "c5walk" else 0)
    for _ in range(3):
        print(what, m.count_matches_dev(dev.data_ptr(), n, stream=st))
elif what == "c3":
    needles = workloads.c3_needles()
    unit = workloads.c3_unit(needles)
    reps = 2 * GIB // unit.size
    n = reps * unit.size
    dev = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
    dev[:n].view(reps, unit.size).copy_(torch.from_numpy(unit).cuda().unsqueeze(0).expand(reps, unit.size))
    m = automaton.AcMachine([(x, i) for i, x in enumerate(needles)], case_sensitivity=1)
    for _ in range(3):
        print(what, m.count_matches_dev(dev.data_ptr(), n, stream=st))
else:
    needles = workloads.c2_needles()
    n = 2 * GIB
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    synth.fill_dev(dev.data_ptr(), n, 0, 43, stream=st); synth.plant_dev(dev.data_ptr(), n, 0, 44, needles, stream=st)
    m = automaton.AcMachine([(x, i) for i, x in enumerate(needles)])
    cnt = m.count_matches_dev(dev.data_ptr(), n, stream=st)
    out = torch.empty(2 * (cnt + 16), dtype=torch.int64, device="cuda")
    for _ in range(3):
        print(what, m.find_all_dev(dev.data_ptr(), n, out.data_ptr(), cnt + 16, stream=st))
