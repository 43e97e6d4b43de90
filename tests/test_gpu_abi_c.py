"""The drop-in boundary driven from plain C on the GPU: every hot-path symbol of include/am_b200.h, by pointer, the way
GHC's `foreign import ccall` would call it (tests/c/abi_driver.c)."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_host_drives_every_symbol(tmp_path):
    lib_dir = os.path.join(ROOT, "alfred-margaret_b200", "lib")
    exe = tmp_path / "abi_driver"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include",
                           os.path.join(ROOT, "tests", "c", "abi_driver.c"), "-o", str(exe), "-L", lib_dir, "-lam_b200", "-Wl,-rpath," + lib_dir,
                           "-L", "/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath,/usr/local/cuda/lib64"])
    out = subprocess.run([str(exe)], text=True, capture_output=True)
    assert out.returncode == 0, out.stdout + out.stderr
    got = dict(kv.split("=", 1) for line in out.stdout.splitlines() if "=" in line for kv in ([line] if line.count("=") == 1 else line.split(" ")))
    assert got["info"] == "17,6,5,2"
    assert got["find_all"] == "10:0,11:1,22:1,26:0,27:1"                     # README.md:94-100
    assert got["overflow_needs"] == "5"
    assert got["contains_any"] == "0,1" and got["count"] == "0,2"            # AhoCorasickSpec.hs:169-179
    assert got["contains_all"] == "1,0"
    assert got["dev"] == "5,1,5" and got["first"] == "1010:0"                # pos_base is added to every position
    assert got["slice_off"] == "5" and sum(map(int, got["shards"].split("+"))) == 5 and got["any"] == "1" and got["allreduce"] == "41"
    assert got["replace_ic"] == "BAZ BAZ" and got["passes"] == "2"           # foo -> BAR, then bar -> BAZ (AhoCorasickSpec.hs:117-118)
    assert got["replace_cs"] == "Foo BAR"                                    # CaseSensitive: "BAR" is not "bar"
    assert got["exceeded"] == "1"
    assert got["replace_dev"] == "BAZ BAZ"
    assert got["replace_stored"] == "un bolt"
    assert got["lower_len"] == "4" and got["skip"] == "0"
    assert "case sensitivity" in got["error"]
