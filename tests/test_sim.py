"""The reducer's rule logic (csrc/tc_core.cuh -- the code the CUDA kernel executes) run through the single-threaded
simulator tests/host_sim/sim_tc.cpp and compared with the oracle.  The simulator permutes the order in which the
iterations of every parallel-for run: a verdict that depends on the permutation would expose an ordering race."""
import os
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIM_SRC = os.path.join(ROOT, "tests", "host_sim", "sim_tc.cpp")
SIM_BIN = os.path.join(ROOT, "tests", "host_sim", "build", "sim_tc")


@pytest.fixture(scope="module")
def sim():
    os.makedirs(os.path.dirname(SIM_BIN), exist_ok=True)
    deps = [SIM_SRC, os.path.join(ROOT, "phylo_tools_b200", "csrc", "tc_core.cuh")]
    if not os.path.exists(SIM_BIN) or any(os.path.getmtime(d) > os.path.getmtime(SIM_BIN) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", SIM_SRC, "-o", SIM_BIN])
    return SIM_BIN


def run_file(sim, pairs, perm):
    with tempfile.NamedTemporaryFile("w", suffix=".nw", delete=False) as f:
        for h, g in pairs:
            f.write(h + "\n" + g + "\n")
    try:
        out = subprocess.run([sim, "file", f.name, str(perm)], capture_output=True, text=True, check=True).stdout.strip()
    finally:
        os.unlink(f.name)
    return out


@pytest.mark.parametrize("perm", [0, 1, 5, 23])
def test_simulated_kernel_logic_on_golden_instances(sim, gold_tc, perm):
    items = gold_tc["items"]
    out = run_file(sim, [(it["host"], it["guest"]) for it in items], perm)
    assert len(out) == len(items)
    bad = [(i, it["tag"], out[i], it["truth"]) for i, it in enumerate(items) if out[i] != str(it["truth"])]
    assert not bad, bad[:5]


def test_simulated_kernel_logic_generated(sim, pt, oracle):
    for (l, r, seed) in [(7, 3, 1), (9, 4, 2), (14, 6, 3)]:
        for kind in (0, 1, 2):
            p = pt.Packer()
            p.add_generated(60, l, r, seed * 1000 + kind, kind)
            truth = "".join(str(oracle.tc_bruteforce_newick(*p.newick(i))[0]) for i in range(60))
            for perm in (0, 3):
                got = subprocess.run([sim, "gen", "60", str(l), str(r), str(seed * 1000 + kind), str(kind), str(perm)], capture_output=True, text=True, check=True).stdout.strip()
                assert got == truth, (l, r, kind, perm)


def test_simulated_config2_positives(sim):
    # BASELINE config 2 shape: l=100, r=20; positives are displayed by construction
    got = subprocess.run([sim, "gen", "40", "100", "20", "45568", "0", "0"], capture_output=True, text=True, check=True).stdout.strip()
    assert got == "1" * 40
