"""The `tc` / `gen` command-line replacements and the C++ facade: they build on CPU (link test), and on the GPU box the
CLI protocol of examples/tc.cpp is checked (verdict = last line of stdout, exit codes)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "phylo_tools_b200", "bin")
FACADE = os.path.join(ROOT, "tests", "cpp", "build", "test_facade")


@pytest.fixture(scope="module")
def tools(pt):
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "phylo_tools_b200", "csrc"), "tools"])   # no-op when up to date
    os.makedirs(os.path.dirname(FACADE), exist_ok=True)
    src = os.path.join(ROOT, "tests", "cpp", "test_facade.cpp")
    deps = [src, os.path.join(ROOT, "include", "pt_gpu.h"), os.path.join(ROOT, "phylo_tools_b200", "libptgpu.so")]
    hdr_dir = os.path.join(ROOT, "phylo_tools_b200", "include", "pt")
    deps += [os.path.join(hdr_dir, f) for f in os.listdir(hdr_dir)]
    if not os.path.exists(FACADE) or os.path.getmtime(FACADE) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", src, "-o", FACADE, "-L" + os.path.join(ROOT, "phylo_tools_b200"), "-lptgpu",
                               "-Wl,-rpath," + os.path.join(ROOT, "phylo_tools_b200")])
    return BIN


def test_tools_build_and_usage_errors(tools, tmp_path):
    r = subprocess.run([os.path.join(tools, "tc"), str(tmp_path / "missing.nw")], capture_output=True, text=True)
    assert r.returncode == 1 and "cannot be opened for reading" in r.stderr       # examples/tc.cpp:41-44
    r = subprocess.run([os.path.join(tools, "tc"), "-r", "0", "5", "1"], capture_output=True, text=True)
    assert r.returncode == 1 and "cannot construct tree" in r.stderr               # examples/tc.cpp:31-34
    r = subprocess.run([os.path.join(tools, "gen"), "-l", "5", "-r", "2", "--with-tree"], capture_output=True, text=True)
    lines = r.stdout.strip().split("\n")
    assert r.returncode == 0 and len(lines) == 2 and lines[0].count("#H") == 4 and lines[1].count(",") == 4
    assert os.path.exists(FACADE)


@pytest.mark.gpu
def test_tc_cli_protocol(tools, tmp_path):
    tc = os.path.join(tools, "tc")
    f = tmp_path / "test.nw"
    f.write_text("((a,(((b)#H1,(c)#H2),(#H1,#H2))),d);\n((a,b),(c,d));\n")    # data/test.nw
    r = subprocess.run([tc, str(f)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip().split("\n")[-1] == "not displayed"
    f.write_text("((a,b),(c,d));\n((a,(b)#H1),(#H1,(c,d)));\n")                # tree first: the network becomes the host
    r = subprocess.run([tc, "-v", str(f)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip().split("\n")[-1] == "displayed"
    f.write_text("((a,(b)#H1),(#H1,c));\n((a,(b)#H1),(#H1,c));\n")
    r = subprocess.run([tc, str(f)], capture_output=True, text=True)
    assert r.stdout.strip().split("\n")[-1] == "sorry, can't check network-network containment yet..."
    r = subprocess.run([tc, "-r", "30", "60", "10"], capture_output=True, text=True)  # self-test mode: always displayed
    assert r.returncode == 0 and r.stdout.strip().split("\n")[-1] == "displayed"
    g = subprocess.run([os.path.join(tools, "gen"), "-l", "40", "-r", "8", "--with-tree", "--seed", "5", str(tmp_path / "g.nw")], capture_output=True, text=True)
    assert g.returncode == 0
    r = subprocess.run([tc, str(tmp_path / "g.nw")], capture_output=True, text=True)
    assert r.stdout.strip().split("\n")[-1] == "displayed"
    b = tmp_path / "batch.nw"
    b.write_text("((a,b),(c,d));\n((a,b),(c,d));\n((a,b),(c,d));\n((a,c),(b,d));\n")
    r = subprocess.run([tc, "--batch", str(b)], capture_output=True, text=True)
    assert r.stdout.strip().split("\n") == ["displayed", "not displayed"]


@pytest.mark.gpu
def test_cpp_facade(tools):
    r = subprocess.run([FACADE], capture_output=True, text=True)
    assert r.returncode == 0 and "facade ok" in r.stdout, r.stderr[-2000:]


@pytest.mark.gpu
def test_bridges_and_components_in_reference_order(tools, tmp_path):
    """pt/bridges.hpp over pt_net_bridges: bridges in the post-order of the reference BridgeFinder (utils/bridges.hpp:73-103) and
    the components its BiconnectedComponents iterator yields (utils/biconnected_comps.hpp:34-96), pinned by
    tests/golden/bcc_ref.json (generated from the reference itself, oracle/make_golden.py make_bcc)"""
    import json
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "bcc_ref.json")))
    items = [it for it in gold["items"] if "bridges" in it]
    assert len(items) >= 40
    f = tmp_path / "nets.nw"
    f.write_text("".join(it["newick"] + "\n" for it in items))
    r = subprocess.run([FACADE, "bcc", str(f)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    blocks = [b.strip("\n").split("\n") for b in r.stdout.split("--\n")][:len(items)]
    for it, b in zip(items, blocks):
        bridges = [list(map(int, x.split()[1:])) for x in b if x.startswith("B ")]
        comps = [[list(map(int, e.split(","))) for e in x.split()[1:]] for x in b if x.startswith("C")]
        assert bridges == it["bridges"], it["newick"]
        assert comps == it["components"], it["newick"]


@pytest.mark.gpu
def test_iso_cli_protocol(tools, tmp_path):
    """the `iso` replacement (examples/iso.cpp): verdict = last line of stdout, data/nets.nw of the reference is isomorphic"""
    iso = os.path.join(tools, "iso")
    f = tmp_path / "nets.nw"
    f.write_text("(a,(((b)#H1,(c)#H2),(#H1,#H2)),d);\n(d,(((b)#H1,(c)#H2),(#H1,#H2)),a);\n")   # data/nets.nw
    r = subprocess.run([iso, str(f)], capture_output=True, text=True)
    lines = r.stdout.strip().split("\n")
    assert r.returncode == 0 and lines[0] == "reading networks..." and lines[-2] == "checking isomorphism..." and lines[-1] == "isomorph!"
    g = tmp_path / "one.nw"
    h = tmp_path / "two.nw"
    g.write_text("((a,b),(c,d));\n"); h.write_text("((a,c),(b,d));\n")
    r = subprocess.run([iso, str(g), str(h)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip().split("\n")[-1] == "not isomorph!"
    r = subprocess.run([iso, "-il", str(g), str(h)], capture_output=True, text=True)    # leaf labels need not match: same shape
    assert r.returncode == 0 and r.stdout.strip().split("\n")[-1] == "isomorph!"
    r = subprocess.run([iso, str(tmp_path / "missing.nw")], capture_output=True, text=True)
    assert r.returncode == 1 and "cannot be opened for reading" in r.stderr
    b = tmp_path / "batch.nw"
    b.write_text(f.read_text() + g.read_text() + h.read_text())
    r = subprocess.run([iso, "--batch", str(b)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip().split("\n") == ["isomorph!", "not isomorph!"]
