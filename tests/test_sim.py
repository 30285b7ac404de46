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
_PT_HDR = os.path.join(ROOT, "phylo_tools_b200", "include", "pt")
HOST_HEADERS = [os.path.join(ROOT, "include", "pt_gpu.h")] + [os.path.join(_PT_HDR, f) for f in sorted(os.listdir(_PT_HDR))]


@pytest.fixture(params=["small-core", "always-branch"], autouse=True)
def finishing_mode(request, monkeypatch):
    """every test runs twice: with the small-core enumeration that finishes a stuck instance on the spot (the default), and with it
    switched off, so that branching (emit_child / keep_only / the pool of children) stays under test"""
    if request.param == "always-branch":
        monkeypatch.setenv("SIM_NO_CORE", "1")
    else:
        monkeypatch.delenv("SIM_NO_CORE", raising=False)


@pytest.fixture(scope="module")
def sim():
    os.makedirs(os.path.dirname(SIM_BIN), exist_ok=True)
    deps = [SIM_SRC, os.path.join(ROOT, "phylo_tools_b200", "csrc", "tc_core.cuh")] + HOST_HEADERS
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


# ---- general DAG hosts: junctions, non-binary nodes, multi-parent leaves, unlabelled leaves, labels on internal nodes -------
def _fuzz_instance(rng):
    import numpy as np
    n = int(rng.integers(2, 15))
    edges = set()
    for v in range(1, n):
        edges.add((int(rng.integers(0, v)), v))
    for _ in range(int(rng.integers(0, 6))):
        v = int(rng.integers(1, n)); u = int(rng.integers(0, v))
        edges.add((u, v))
    edges = sorted(edges)
    out = np.zeros(n, dtype=int)
    for a, _ in edges:
        out[a] += 1
    hl = np.full(n, -1, dtype=np.int32)
    nxt = 0
    for v in range(n):
        if out[v] == 0 and rng.random() < 0.85:
            hl[v] = nxt; nxt += 1
        elif out[v] > 0 and rng.random() < 0.05:
            hl[v] = 100 + v                      # label on an internal node: must be ignored
    # guest: tree displayed by a random switching (plain python), then perturbed
    par = {}
    for v in range(1, n):
        ps = [a for a, b in edges if b == v]
        par[v] = ps[int(rng.integers(0, len(ps)))]
    kids = {v: [] for v in range(n)}
    for v, p in par.items():
        kids[p].append(v)
    def build(v):
        if out[v] == 0:
            return ("L", int(hl[v])) if hl[v] >= 0 else None
        ch = [c for c in (build(k) for k in kids[v]) if c is not None]
        if not ch:
            return None
        return ch[0] if len(ch) == 1 else ("N", ch)
    t = build(0)
    labels = [int(x) for x in hl if 0 <= x < 100]
    mode = int(rng.integers(0, 6))
    tp, tl = [], []
    def emit(node, parent):
        i = len(tp); tp.append(parent)
        if node[0] == "L":
            tl.append(node[1])
        else:
            tl.append(-1)
            order = list(node[1]); rng.shuffle(order)
            for c in order:
                emit(c, i)
        return i
    if t is None:
        tp, tl = [-1], [-1]
    else:
        if mode == 5 and rng.random() < 0.5:     # unary root on top + an unlabelled dangling leaf: both must be cleaned away
            tp.append(-1); tl.append(-1); emit(t, 0); tp.append(0); tl.append(-1)
        else:
            emit(t, -1)
    tl = np.array(tl, dtype=np.int32); tp = np.array(tp, dtype=np.int32)
    leaves = [i for i in range(len(tl)) if tl[i] >= 0]
    if mode == 1 and len(leaves) >= 2:           # swap two labels
        a, b = rng.choice(leaves, 2, replace=False); tl[a], tl[b] = tl[b], tl[a]
    elif mode == 2 and leaves:                   # guest-only label
        tl[leaves[0]] = 99
    elif mode == 3 and len(leaves) >= 3:         # a label only the host has: drop one guest leaf (leaves a unary node behind)
        tl[leaves[-1]] = -1
    elif mode == 4 and len(leaves) >= 4:         # rotate three labels
        a, b, c = rng.choice(leaves, 3, replace=False); tl[a], tl[b], tl[c] = tl[b], tl[c], tl[a]
    return n, edges, hl, tp, tl


def test_simulated_kernel_logic_general_dags(sim, oracle):
    import numpy as np
    rng = np.random.default_rng(20261019)
    insts = [_fuzz_instance(rng) for _ in range(1500)]
    truth = []
    with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
        for n, edges, hl, tp, tl in insts:
            te = [(int(tp[i]), i) for i in range(len(tp)) if tp[i] >= 0]
            truth.append(oracle.tc_bruteforce(edges, n, hl, te, len(tp), tl)[0])
            f.write("%d %d %d\n" % (n, len(edges), len(tp)))
            f.write("".join("%d %d\n" % e for e in edges))
            f.write(" ".join(map(str, hl.tolist())) + "\n" + " ".join(map(str, tp.tolist())) + "\n" + " ".join(map(str, tl.tolist())) + "\n")
    try:
        assert min(truth) >= 0 and 0.15 < sum(truth) / len(truth) < 0.9
        for perm in (0, 1, 9):
            got = subprocess.run([sim, "arrays", f.name, str(perm)], capture_output=True, text=True, check=True).stdout.strip()
            bad = [(i, got[i], truth[i]) for i in range(len(truth)) if got[i] != str(truth[i])]
            assert not bad, (perm, bad[:5], insts[bad[0][0]])
    finally:
        os.unlink(f.name)


# ---- multi-threaded simulator: real barriers and atomics (what the warp / CTA executors do) -------------------------------
SIM_MT_SRC = os.path.join(ROOT, "tests", "host_sim", "sim_tc_mt.cpp")
SIM_MT_BIN = os.path.join(ROOT, "tests", "host_sim", "build", "sim_tc_mt")


@pytest.fixture(scope="module")
def sim_mt():
    os.makedirs(os.path.dirname(SIM_MT_BIN), exist_ok=True)
    deps = [SIM_MT_SRC, os.path.join(ROOT, "phylo_tools_b200", "csrc", "tc_core.cuh")] + HOST_HEADERS
    if not os.path.exists(SIM_MT_BIN) or any(os.path.getmtime(d) > os.path.getmtime(SIM_MT_BIN) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-pthread", SIM_MT_SRC, "-o", SIM_MT_BIN])
    return SIM_MT_BIN


@pytest.mark.parametrize("threads,int32", [(3, 0), (8, 1)])
def test_group_execution_with_real_barriers(sim_mt, gold_tc, threads, int32):
    """T threads run the solver SPMD over one shared workspace; a barrier not reached by every thread deadlocks (the
    simulator aborts after 20 s), non-uniform return codes abort, verdicts must equal the pinned truth"""
    items = gold_tc["items"][:300]
    with tempfile.NamedTemporaryFile("w", suffix=".nw", delete=False) as f:
        for it in items:
            f.write(it["host"] + "\n" + it["guest"] + "\n")
    try:
        r = subprocess.run([sim_mt, "file", f.name, str(threads), str(int32)], capture_output=True, text=True, timeout=600)
    finally:
        os.unlink(f.name)
    assert r.returncode == 0, r.stderr[-500:]
    out = r.stdout.strip()
    assert [c for c in out] == [str(it["truth"]) for it in items]


def test_simulated_kernel_logic_at_headline_size(sim):
    """the reference's verdicts and the exhaustive truth at l=100, r=20 (tests/golden/tc_ref_l100.json, tc_exh_l100.json) against
    the kernel logic, on the CPU: the first 40 instances of every config, read through the uint8 layout the bench uploads"""
    from conftest import load_golden
    env = dict(os.environ, SIM_LAYOUT="1")
    for cfg in load_golden("tc_ref_l100.json")["configs"]:
        if cfg["leaves"] != 100:
            continue
        got = subprocess.run([sim, "gen", "40", "100", "20", str(cfg["seed"]), str(cfg["kind"]), "3"], capture_output=True, text=True, check=True, env=env).stdout.strip()
        settled = cfg["settled"]
        bad = [(i, r, got[i]) for i, r in enumerate(cfg["reference"][:40]) if r in "01" and got[i] != str(settled.get(str(i), r))]
        assert not bad, (cfg["kind"], bad)
    for cfg in load_golden("tc_exh_l100.json")["configs"]:
        got = subprocess.run([sim, "gen", "40", "100", "20", str(cfg["seed"]), str(cfg["kind"]), "0"], capture_output=True, text=True, check=True, env=env).stdout.strip()
        assert got == cfg["truth"][:40], cfg["kind"]


@pytest.mark.parametrize("layout", ["2", "1"])
def test_simulated_layouts_agree(sim, layout):
    """int16 and uint8 (out-degree) input layouts, children emitted in the same layout: same verdicts as int32"""
    for kind in (0, 1, 2):
        ref = subprocess.run([sim, "gen", "80", "12", "6", str(900 + kind), str(kind), "0"], capture_output=True, text=True, check=True).stdout.strip()
        got = subprocess.run([sim, "gen", "80", "12", "6", str(900 + kind), str(kind), "5"], capture_output=True, text=True, check=True,
                             env=dict(os.environ, SIM_LAYOUT=layout)).stdout.strip()
        assert got == ref and "E" not in got
