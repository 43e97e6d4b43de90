"""Replacer.run timing on BASELINE.json config 4 (development aid)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "alfred-margaret_b200")]
import numpy as np
from alfred_margaret_b200 import replacer, synth
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2 << 30
rng = np.random.default_rng(64)
needles = synth.random_needles(5000, 62, 4, 16)
repls = [bytes(rng.integers(ord("A"), ord("Z") + 1, size=int(rng.integers(0, 25)), dtype=np.uint8)) for _ in needles]
t0 = time.time(); hay = synth.fill_host(0, n, 63); synth.plant_host(hay, 0, 64, needles[:64]); print("generated %.1f s" % (time.time() - t0), flush=True)
r = replacer.build(0, list(zip(needles, repls)))
for _ in range(2):
    t0 = time.time(); out = replacer.run(r, hay); dt = time.time() - t0
    print("Replacer.run %d B -> %d B, %d passes (%d full scans), %.3f s (%.2f ms/pass)" % (n, len(out), r.last_passes, r.last_rescans, dt, dt / r.last_passes * 1e3), flush=True)
