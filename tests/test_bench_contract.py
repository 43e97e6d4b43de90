"""bench.py's reference arm runs on the CPU (rank 0, oracle port): check the one-line JSON contract here, without a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1", "--ref-sample", str(8 << 20)],
                         capture_output=True, text=True, timeout=300, check=True).stdout
    lines = [ln for ln in out.splitlines() if ln.strip()]
    assert len(lines) == 1, lines                             # exactly one line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GB/s" and d["higher_is_better"] is True and d["dtype"] == "u8"
    assert d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["matches_per_step"] > 0


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_workloads_are_reproducible_and_well_formed():
    """The five configurations' generators (workloads.py) are pure functions of their seeds: the oracle windows of bench.py and of
    the full-size tests only mean something if host and device see the same bytes run after run.  C3's unit is valid UTF-8 that
    ends on a code point boundary (units are laid back to back), holds all four UTF-8 lengths and re-cased needles."""
    import hashlib
    import os, sys
    sys.path.insert(0, os.path.join(ROOT, "alfred-margaret_b200"))
    from alfred_margaret_b200 import workloads
    n1, h1 = workloads.c1()
    assert n1 == [b"tshirt", b"shirts", b"shorts"] and h1.size == 1000000
    c2 = workloads.c2_needles()
    assert len(c2) == 1000 and min(map(len, c2)) == 4 and max(map(len, c2)) == 16 and c2 == workloads.c2_needles()
    c5 = workloads.c5_needles(2000)
    assert min(map(len, c5)) >= 6 and max(map(len, c5)) <= 16 and c5 == workloads.c5_needles(2000)
    nd, rp = workloads.c4_pairs(300)
    assert len(nd) == len(rp) == 300 and all(0 <= len(r) <= 24 for r in rp) and (nd, rp) == workloads.c4_pairs(300)
    c3 = workloads.c3_needles(1500)
    assert c3 == workloads.c3_needles(1500) and all(x.decode("utf-8") == x.decode("utf-8").lower() for x in c3)
    assert any(max(x) >= 0x80 for x in c3)                      # needles with non-ASCII code points
    u1 = workloads.c3_unit(c3, size=1 << 20)
    u2 = workloads.c3_unit(c3, size=1 << 20)
    assert u1.size == 1 << 20 and hashlib.sha256(u1.tobytes()).digest() == hashlib.sha256(u2.tobytes()).digest()
    text = u1.tobytes().decode("utf-8")                         # raises if the unit is not valid UTF-8 up to its very end
    lens = {len(c.encode("utf-8")) for c in text[:200000]}
    assert lens == {1, 2, 3, 4}
    assert any(c.isupper() for c in text[:1000]) and any(c.islower() for c in text[:1000])
