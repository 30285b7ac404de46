"""oracle/make_golden_dp.py -- regenerates the fixtures of the general tree-in-tree DP IN THE BUILD CONTAINER (needs the reference
through oracle/_ref/ref_tc; the committed JSON travels to the GPU box).

  tests/golden/whomul_ref.json  (multi-labelled host tree, guest tree) -> who_displays(u) of the REFERENCE for every guest node
        (TreeInTreeContainment<ROMulTree<>, ROTree<>>, utils/containment.hpp:27-233, via `ref_tc whomul`); data/test_tree.nw first
  tests/golden/whowide_ref.json the same for guests (and hosts) with WIDE nodes: 65 ... 300 children on one node (stars, wide multifurcations,
        single- and multi-labelled hosts) -- beyond one 64-bit child-set word
  tests/golden/hda_ref.json     (host network, guest tree) -> the MUL-tree the REFERENCE unzips below the host's root
        (TreeInComponent::unzip_retis :262-312) and highest_displayed_ancestor(v) of every guest node (:328-344), via `ref_tc hda`
  tests/golden/comps_ref.json   network -> comp_root / visible_leaf per node and the component DAG of the REFERENCE's
        TreeComponentInfos (utils/tree_components.hpp:41-234), via `ref_tc comps`
  tests/golden/vcr_ref.json     (host network, guest tree) -> the REFERENCE's VisibleComponentRule (utils/containment.hpp:603-725), via `ref_tc vcr`
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import phylo_tools_b200 as pt  # noqa: E402  (host-side generator only)

GOLD = os.path.join(ROOT, "tests", "golden")
REF_TC = os.path.join(ROOT, "oracle", "_ref", "ref_tc")


def run_lines(mode, lines):
    with tempfile.NamedTemporaryFile("w", suffix=".nw", delete=False) as f:
        f.write("\n".join(lines) + "\n")
    out = subprocess.run([REF_TC, mode, f.name], capture_output=True, text=True).stdout
    os.unlink(f.name)
    return [b.strip("\n").split("\n") for b in out.split("--\n")]


def name(i):
    s = ""
    while True:
        s = chr(ord("a") + i % 26) + s
        if i < 26:
            return s
        i //= 26


def rand_tree(rng, leaves, max_deg):
    nodes = list(leaves)
    rng.shuffle(nodes)
    while len(nodes) > 1:
        k = int(min(len(nodes), rng.integers(2, max_deg + 1)))
        idx = sorted(rng.choice(len(nodes), k, replace=False).tolist(), reverse=True)
        kids = [nodes.pop(i) for i in idx]
        nodes.insert(int(rng.integers(0, len(nodes) + 1)), "(" + ",".join(kids) + ")")
    return nodes[0] + ";"


def restricted(rng, host_nw):
    """the tree a MUL-tree displays when one leaf per label is kept (chosen at random), unary nodes suppressed"""
    from oracle.newick import parse_newick
    n, edges, lab = parse_newick(host_nw)
    kids = {v: [] for v in range(n)}
    for a, b in edges:
        kids[a].append(b)
    by_label = {}
    for v, name_ in lab.items():
        if not kids[v] and name_:
            by_label.setdefault(name_, []).append(v)
    keep = {vs[int(rng.integers(0, len(vs)))] for vs in by_label.values()}

    def build(v):
        if not kids[v]:
            return lab[v] if v in keep else None
        parts = [x for x in (build(c) for c in kids[v]) if x is not None]
        if not parts:
            return None
        return parts[0] if len(parts) == 1 else "(" + ",".join(parts) + ")"
    return build(0) + ";"


def make_whomul():
    rng = np.random.default_rng(20261020)
    pairs = [("((((a,b),(c,d)),(e,d)),((a,b),(e,d)));", "(e,((a,c),(b,d)));")]   # data/test_tree.nw
    for trial in range(220):
        l = int(rng.integers(3, 9))
        labs = [name(i) for i in range(l)]
        mult = int(rng.integers(1, 4))
        host_leaves = labs + [labs[int(rng.integers(0, l))] for _ in range(int(rng.integers(0, mult * l + 1)))]
        host = rand_tree(rng, host_leaves, 2 if trial % 3 else 4)
        mode = trial % 4
        if mode == 0:      # guest = the host restricted to one (random) leaf per label: displayed by construction
            guest = restricted(rng, host)
        elif mode == 1:
            guest = rand_tree(rng, labs, 2 if trial % 2 else 3)
        elif mode == 2:
            guest = rand_tree(rng, labs[: max(2, l - 1)], 3)
        else:
            guest = rand_tree(rng, labs, 4)
        pairs.append((host, guest))
    blocks = run_lines("whomul", [x for p in pairs for x in p])[:len(pairs)]
    items = []
    for (h, g), b in zip(pairs, blocks):
        if any(x in ("CRASH", "EXC") for x in b):
            items.append(dict(host=h, guest=g, reference="CRASH"))
            continue
        items.append(dict(host=h, guest=g, reference=[[int(x) for x in line.split(":")[1].split()] for line in b if ":" in line]))
    crashed = sum(1 for it in items if it["reference"] == "CRASH")
    json.dump(dict(source="oracle/_ref/ref_tc whomul (reference TreeInTreeContainment<ROMulTree, ROTree>::who_displays, containment.hpp:72-230)", crashed=crashed, items=items),
              open(os.path.join(GOLD, "whomul_ref.json"), "w"), indent=0)
    disp = sum(1 for it in items if it["reference"] != "CRASH" and it["reference"][0])
    print("whomul_ref.json:", len(items), "pairs,", crashed, "reference crashes,", disp, "with a displayed guest root")


def wide_tree(rng, leaves, wide, max_deg):
    """random tree over `leaves` in which one join takes `wide` subtrees at once (a node with `wide` children)"""
    nodes = list(leaves)
    rng.shuffle(nodes)
    done_wide = False
    while len(nodes) > 1:
        if not done_wide and len(nodes) >= wide and rng.integers(0, 3) == 0 or (not done_wide and len(nodes) == wide):
            k = wide; done_wide = True
        else:
            k = int(min(len(nodes), rng.integers(2, max_deg + 1)))
            if not done_wide and len(nodes) - k + 1 < wide:
                k = wide if len(nodes) >= wide else k; done_wide = done_wide or k == wide
        idx = sorted(rng.choice(len(nodes), k, replace=False).tolist(), reverse=True)
        kids = [nodes.pop(i) for i in idx]
        nodes.insert(int(rng.integers(0, len(nodes) + 1)), "(" + ",".join(kids) + ")")
    return nodes[0] + ";"


def make_whowide():
    rng = np.random.default_rng(20261021)
    pairs = []
    for trial in range(36):
        wide = int(rng.choice([65, 66, 70, 100, 128, 129, 200, 300]))
        l = wide + int(rng.integers(0, 40))
        labs = [name(i) for i in range(l)]
        kind = trial % 6
        if kind == 0:      # star guest in a star host (displayed), or in a host whose root is wider still
            host = "(" + ",".join(rng.permutation(labs).tolist()) + ");"
            guest = "(" + ",".join(rng.permutation(labs[:wide]).tolist()) + ");"
        elif kind == 1:    # wide guest = the host itself, children shuffled (displayed)
            host = wide_tree(rng, labs, wide, 3)
            guest = host
        elif kind == 2:    # wide guest against a binary host (not displayed at the wide node)
            host = rand_tree(rng, labs, 2)
            guest = wide_tree(rng, labs, wide, 3)
        elif kind == 3:    # multi-labelled host with a wide node, guest = one leaf per label
            host_leaves = labs + [labs[int(rng.integers(0, l))] for _ in range(int(rng.integers(1, l // 2)))]
            host = wide_tree(rng, host_leaves, wide + 20, 3)
            guest = restricted(rng, host)
        elif kind == 4:    # multi-labelled host, independent wide guest
            host_leaves = labs + [labs[int(rng.integers(0, l))] for _ in range(int(rng.integers(1, l // 2)))]
            host = wide_tree(rng, host_leaves, wide + 20, 4)
            guest = wide_tree(rng, labs, wide, 4)
        else:              # star host with every label twice, star guest: the matching has to choose
            host = "(" + ",".join(rng.permutation(labs + labs).tolist()) + ");"
            guest = "(" + ",".join(rng.permutation(labs).tolist()) + ");"
        pairs.append((host, guest))
    blocks = run_lines("whomul", [x for p in pairs for x in p])[:len(pairs)]
    items = []
    for (h, g), b in zip(pairs, blocks):
        if any(x in ("CRASH", "EXC") for x in b):
            items.append(dict(host=h, guest=g, reference="CRASH"))
            continue
        items.append(dict(host=h, guest=g, reference=[[int(x) for x in line.split(":")[1].split()] for line in b if ":" in line]))
    crashed = sum(1 for it in items if it["reference"] == "CRASH")
    json.dump(dict(source="oracle/_ref/ref_tc whomul on pairs with a node of 65 ... 320 children (reference TreeInTreeContainment::who_displays, containment.hpp:72-230)", crashed=crashed, items=items),
              open(os.path.join(GOLD, "whowide_ref.json"), "w"), indent=0)
    disp = sum(1 for it in items if it["reference"] != "CRASH" and it["reference"][0])
    print("whowide_ref.json:", len(items), "pairs,", crashed, "reference crashes,", disp, "with a displayed guest root")


def gen_pairs(count, l, r, seed, kind):
    p = pt.Packer()
    p.add_generated(count, l, r, seed, kind)
    return [p.newick(i) for i in range(count)]


def make_hda():
    pairs = [("((a,(((b)#H1,(c)#H2),(#H1,#H2))),d);", "((a,b),(c,d));"), ("((a,(b)#H1),(#H1,c));", "(a,(b,c));"), ("((a,(b)#H1),(#H1,c));", "((a,b),c);")]
    for (l, r, cnt) in [(4, 1, 10), (6, 2, 14), (8, 3, 14), (10, 4, 12), (12, 5, 10), (16, 6, 8)]:
        for kind in (0, 1, 2):
            pairs += gen_pairs(cnt, l, r, 8800 + 100 * l + kind, kind)
    blocks = run_lines("hda", [x for p in pairs for x in p])[:len(pairs)]
    items = []
    for (h, g), b in zip(pairs, blocks):
        if any(x in ("CRASH", "EXC") for x in b) or not b or not b[0].startswith("U "):
            items.append(dict(host=h, guest=g, reference="CRASH"))
            continue
        k = int(b[0].split()[1])
        tree = [(int(x.split()[0]), x.split()[1]) for x in b[1:1 + k]]
        hda = {x.split(":")[0]: int(x.split(":")[1]) for x in b[1 + k:] if ":" in x}
        items.append(dict(host=h, guest=g, mul_parent=[t[0] for t in tree], mul_label=[t[1] for t in tree], hda=hda))
    crashed = sum(1 for it in items if it.get("reference") == "CRASH")
    json.dump(dict(source="oracle/_ref/ref_tc hda (reference TreeInComponent: unzip_retis containment.hpp:262-312, highest_displayed_ancestor :328-344, at the host's root)",
                   crashed=crashed, items=items), open(os.path.join(GOLD, "hda_ref.json"), "w"), indent=0)
    print("hda_ref.json:", len(items), "pairs,", crashed, "reference crashes")


def make_comps():
    lines = ["((a,(((b)#H1,(c)#H2),(#H1,#H2))),d);", "((a,(b)#H1),(#H1,c));", "(((a,b),((c)#H1,d)),(#H1,e));", "((((a,b))#H1,c),(#H1,d));",
             "(((((a,b))#H2)#H1,c),((#H1,d),(#H2,e)));"]
    for (l, r, cnt) in [(5, 2, 15), (8, 3, 15), (12, 5, 12), (20, 8, 10), (40, 12, 6), (100, 20, 4)]:
        lines += [h for h, _ in gen_pairs(cnt, l, r, 9900 + l, 0)]
    blocks = run_lines("comps", lines)[:len(lines)]
    items = []
    for s, b in zip(lines, blocks):
        if any(x in ("CRASH", "EXC") for x in b):
            items.append(dict(newick=s, reference="CRASH"))
            continue
        nodes = [list(map(int, x.split())) for x in b if x and not x.startswith("D")]
        dag = [list(map(int, x.split()[1:])) for x in b if x.startswith("D ")]
        items.append(dict(newick=s, comp_root=[x[1] for x in nodes], visible_leaf=[x[2] for x in nodes], dag=dag))
    crashed = sum(1 for it in items if it.get("reference") == "CRASH")
    json.dump(dict(source="oracle/_ref/ref_tc comps (reference TreeComponentInfos, utils/tree_components.hpp:41-234)", crashed=crashed, items=items),
              open(os.path.join(GOLD, "comps_ref.json"), "w"), indent=0)
    print("comps_ref.json:", len(items), "networks,", crashed, "reference crashes,", sum(len(it.get("dag", [])) for it in items), "component-DAG arcs")




def make_edgelist():
    """edge-list texts -> (num_nodes, edges, labels) of the REFERENCE's parse_edgelist (io/edgelist.hpp:41-86, via `ref_tc parse_el`),
    malformed texts included ("ERR" = MalformedEdgeVec)"""
    texts = ["a b\na c\nb d\nb e\n", "root x\nroot y\nx l1\nx r\ny r\nr l2\n", "1 2\n2 3\n3 4\n", "a\tb\nb\tc\n", "u v\n", "a b\nc d\nb c\n",
             "a  b\n a c\n", "a b", "a b\n\n\nb c\n", "a b c\n", "a\n", "", "x y \nz w\n", "n1 n2\nn1 n3\nn2 n4\nn3 n4\nn4 n5\n"]
    rng = np.random.default_rng(5)
    for _ in range(6):
        n = int(rng.integers(4, 30))
        lines = []
        for v in range(1, n):
            lines.append("v%d v%d" % (int(rng.integers(0, v)), v))
        rng.shuffle(lines)
        texts.append("\n".join(lines) + "\n")
    items = []
    for t in texts:
        with tempfile.NamedTemporaryFile("w", suffix=".el", delete=False) as f:
            f.write(t)
        out = subprocess.run([REF_TC, "parse_el", f.name], capture_output=True, text=True)
        os.unlink(f.name)
        b = out.stdout.split("--\n")[0].strip("\n").split("\n")
        if out.returncode != 0 or not b or b[0] == "ERR" or b[0] == "":
            items.append(dict(text=t, reference="ERR" if out.returncode == 0 else "CRASH"))
            continue
        n, m = map(int, b[0].split())
        edges = [list(map(int, x.split())) for x in b[1:1 + m]]
        labels = {x.split(" ", 2)[1]: x.split(" ", 2)[2] for x in b[1 + m:]}
        items.append(dict(text=t, num_nodes=n, edges=edges, labels=labels))
    json.dump(dict(source="oracle/_ref/ref_tc parse_el (reference io/edgelist.hpp parse_edgelist)", items=items), open(os.path.join(GOLD, "edgelist_ref.json"), "w"), indent=0)
    print("edgelist_ref.json:", len(items), "texts,", sum(1 for it in items if "reference" in it), "malformed / crashed")


def make_vcr():
    """(host network, guest tree) -> the REFERENCE's VisibleComponentRule on the fresh pair (utils/containment.hpp:603-725, via `ref_tc vcr`):
    the order in which it visits the component roots and their component-DAG children (its hash containers decide), is_half_eligible
    and eligibility of every root, and for every eligible root its visible leaf and the guest node treat_comp_root would match it with"""
    pairs = [("((a,(((b)#H1,(c)#H2),(#H1,#H2))),d);", "((a,b),(c,d));"), ("((a,(b)#H1),(#H1,c));", "(a,(b,c));"),
             ("((((a,b),(c)#H1),(#H1,d)),((e,f),g));", "(((a,b),(c,d)),((e,f),g));"),
             # two tree components above one root (x below a reticulation with parents in both): the reference's counters double
             ("(((a,((x,y))#H1),b),((c,#H1),d));", "((a,b),((c,d),(x,y)));")]
    for (l, r, cnt) in [(6, 2, 14), (8, 3, 16), (10, 4, 16), (14, 6, 14), (20, 8, 12), (30, 12, 8), (60, 20, 4)]:
        for kind in (0, 1):
            pairs += gen_pairs(cnt, l, r, 7700 + 100 * l + kind, kind)
    blocks = run_lines("vcr", [x for p in pairs for x in p])[:len(pairs)]
    items = []
    for (h, g), b in zip(pairs, blocks):
        if any(x in ("CRASH", "EXC", "FAIL", "UNCLEAN") for x in b) or not b or not b[0]:
            items.append(dict(host=h, guest=g, reference="CRASH"))
            continue
        order = [[int(x.split(":")[0].split()[1])] + [int(y) for y in x.split(":")[1].split()] for x in b if x.startswith("O ")]
        half = {x.split()[1]: [int(x.split()[2]), int(x.split()[3])] for x in b if x.startswith("H ")}
        res = [[int(y) for y in x.split()[1:]] for x in b if x.startswith("R ") and "NOLABEL" not in x]
        items.append(dict(host=h, guest=g, edgeless=any(x.startswith("EDGELESS") for x in b), order=order, half=half, treated=res))
    crashed = sum(1 for it in items if it.get("reference") == "CRASH")
    json.dump(dict(source="oracle/_ref/ref_tc vcr (reference VisibleComponentRule: is_half_eligible containment.hpp:652-687, get_eligible_component_root :689-704, treat_comp_root :623-644)",
                   crashed=crashed, items=items), open(os.path.join(GOLD, "vcr_ref.json"), "w"), indent=0)
    print("vcr_ref.json:", len(items), "pairs,", crashed, "reference crashes,", sum(len(it.get("treated", [])) for it in items), "eligible roots treated,",
          sum(1 for it in items if it.get("half") and any(v[0] == 0 for v in it["half"].values())), "pairs with a root that is not half-eligible")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "vcr":
        make_vcr()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "whowide":
        make_whowide()
        sys.exit(0)
    make_whomul()
    make_whowide()
    make_hda()
    make_comps()
    make_edgelist()
    make_vcr()
