"""The N > 1 host path on CPU: world_size-2 (and 3) gloo process groups.  The shard planner, the halo rule
and the match-count all-gather are the product's; the scan engine plugged in here is the CPU oracle (test
infrastructure), so the test checks that per-rank lists concatenate to exactly the single-shard list."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [os.path.join(root, "alfred-margaret_b200"), os.path.join(root, "oracle")]
    import am_oracle_py as oracle
    from alfred_margaret_b200 import sharded, synth
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    needles = synth.random_needles(200, 42, 2, 9, b"abcd")
    n = (1 << 20) + 37
    text = synth.fill_host(0, n, 43, b"abcd")
    synth.plant_host(text, 0, 44, needles)
    om = oracle.Machine(needles)
    halo = max(len(x) for x in needles) - 1

    def scan(w, e, report_begin, pos_base):
        ms = om.find_all(text[w:e])
        ms = ms[ms["pos"] > report_begin].copy()
        ms["pos"] += pos_base
        return ms

    matches, off, total = sharded.find_all_sharded(scan, n, halo, rank, world, dist)
    np.save(os.path.join(out_dir, "m%d.npy" % rank), matches)
    np.save(os.path.join(out_dir, "o%d.npy" % rank), np.array([off, total]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_concatenation_equals_single_shard(tmp_path, oracle, world):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    from alfred_margaret_b200 import synth
    needles = synth.random_needles(200, 42, 2, 9, b"abcd")
    n = (1 << 20) + 37
    text = synth.fill_host(0, n, 43, b"abcd")
    synth.plant_host(text, 0, 44, needles)
    want = oracle.Machine(needles).find_all(text)
    parts, offs = [], []
    for r in range(world):
        parts.append(np.load(os.path.join(str(tmp_path), "m%d.npy" % r)))
        offs.append(np.load(os.path.join(str(tmp_path), "o%d.npy" % r)))
    got = np.concatenate(parts)
    assert len(got) == len(want) and np.array_equal(got["pos"], want["pos"]) and np.array_equal(got["value"], want["value"])
    acc = 0
    for r in range(world):
        assert offs[r][0] == acc and offs[r][1] == len(want)
        acc += len(parts[r])
