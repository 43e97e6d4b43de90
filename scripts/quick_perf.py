"""Quick device-resident timing of the scan (development aid; bench.py is the contract): whole-call times by CUDA events."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "alfred-margaret_b200")]
import torch
from alfred_margaret_b200 import automaton, synth
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 30
nn = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
needles = synth.random_needles(nn, 42)
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
synth.fill_dev(dev.data_ptr(), n, 0, 43); synth.plant_dev(dev.data_ptr(), n, 0, 44, needles)
st = torch.cuda.current_stream().cuda_stream
def timeit(f, reps=7):
    f(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts), sum(ts) / len(ts)
m = automaton.AcMachine([(x, i) for i, x in enumerate(needles)])
cnt = m.count_matches_dev(dev.data_ptr(), n, stream=st)
out = torch.empty(2 * (cnt + 16), dtype=torch.int64, device="cuda")
for name, f in (("count", lambda: m.count_matches_dev(dev.data_ptr(), n, stream=st)),
                ("find_all", lambda: m.find_all_dev(dev.data_ptr(), n, out.data_ptr(), cnt + 16, stream=st))):
    mn, av = timeit(f)
    print("%s %-9s n=%.2f GiB needles=%d matches=%d  min %.3f ms  avg %.3f ms  -> %.1f GB/s" % (os.environ.get("TAG", "r2"), name, n / 2**30, nn, cnt, mn, av, n / mn / 1e6), flush=True)
