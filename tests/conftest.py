import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pt():
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(ROOT, "phylo_tools_b200", "libptgpu.so")):
        g.build()
    import phylo_tools_b200
    return phylo_tools_b200


@pytest.fixture(scope="session")
def oracle():
    from oracle import pt_oracle
    pt_oracle.lib()
    return pt_oracle


def load_golden(name):
    with open(os.path.join(GOLD, name)) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def gold_tc():
    return load_golden("tc_ref_verdicts.json")


@pytest.fixture(scope="session")
def gold_newick():
    return load_golden("newick_ref.json")


@pytest.fixture(scope="session")
def gold_lca():
    return load_golden("lca_ref.json")
