"""IgnoreCase scan launches for ncu (development aid): lower_kernel + filter kernel, then forced walk."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "alfred-margaret_b200")]
import torch
from alfred_margaret_b200 import automaton, synth
n = 1 << 30
needles = synth.random_needles(1000, 42)
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
synth.fill_dev(dev.data_ptr(), n, 0, 43, alphabet=synth.AZ + synth.AZ.upper()); synth.plant_dev(dev.data_ptr(), n, 0, 44, needles)
for kind in (0, 1):
    m = automaton.AcMachine([(x, i) for i, x in enumerate(needles)], case_sensitivity=1, force_kernel=kind)
    for _ in range(2):
        print(kind, m.count_matches_dev(dev.data_ptr(), n))
