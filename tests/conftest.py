import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "alfred-margaret_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json"), encoding="utf-8") as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure): compiled on demand from oracle/am_oracle.c."""
    import am_oracle_py
    am_oracle_py.build_lib()
    return am_oracle_py


@pytest.fixture(scope="session")
def lower_dense(oracle):
    from alfred_margaret_b200 import utf8
    return oracle.lower_table_dense(utf8.host_lower_pairs())
