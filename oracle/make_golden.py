"""oracle/make_golden.py -- regenerates tests/golden/*.json IN THE BUILD CONTAINER (needs /root/reference through the
binaries in oracle/_ref/, built by `make -C oracle`).  The GPU box has neither; the committed JSON files travel.

  newick_ref.json      Newick strings -> (num_nodes, edges, labels) as produced by the REFERENCE reader
                       (io/newick.hpp via oracle/_ref/ref_tc parse)
  tc_ref_verdicts.json (host, guest) pairs -> the REFERENCE tc verdict ('1','0','E' exception,'S' crash) and the
                       exhaustive-switching truth from oracle/pt_oracle.c.  Includes data/test.nw and the five reference
                       false positives of SURVEY 4.3.
  lca_ref.json         random trees + queries -> answers of the REFERENCE rmq/lca.hpp (oracle/_ref/ref_lca lca) and
                       RMQ arrays -> (index, value) answers of the reference sparse_rmq / pm_rmq
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import phylo_tools_b200 as pt  # noqa: E402  (host-side generator only; no GPU needed)
from oracle import pt_oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
REF_TC = os.path.join(ROOT, "oracle", "_ref", "ref_tc")
REF_LCA = os.path.join(ROOT, "oracle", "_ref", "ref_lca")

FALSE_POSITIVES = [  # the reference prints "displayed", exhaustive switching says no
    ("((((((d,e))#H6)#H3,(a,b)),#H3),(#H6,c));", "((c,(e,(a,b))),d);"),
    ("(((((((c,d),(e,f)))#H4)#H3,(a,b)),#H3),#H4);", "(((f,(d,c)),(b,a)),e);"),
    ("((((b,c),a),(((d,e))#H6,(f)#H8)),(#H6,#H8));", "(((d,f),(a,(c,b))),e);"),
    ("((((c,d))#H2,((((e,f),a),(b)#H6),#H6)),#H2);", "((((e,a),b),(c,d)),f);"),
    ("((((((a)#H10)#H9,(((f)#H14,c),b)))#H6,#H10),((((d,e),(i,j)),(#H6,(#H9,(g,h)))),#H14));", "(g,(((d,e),(j,i)),((a,h),(b,(c,f)))));"),
]
DATA_TEST = ("((a,(((b)#H1,(c)#H2),(#H1,#H2))),d);", "((a,b),(c,d));")  # data/test.nw


def gen_pairs(count, l, r, seed, kind):
    p = pt.Packer()
    p.add_generated(count, l, r, seed, kind)
    return [p.newick(i) for i in range(count)]


def run_lines(binary, mode, lines):
    with tempfile.NamedTemporaryFile("w", suffix=".nw", delete=False) as f:
        f.write("\n".join(lines) + "\n")
    out = subprocess.run([binary, mode, f.name], capture_output=True, text=True).stdout
    os.unlink(f.name)
    return out


def make_newick():
    lines = []
    for (l, r) in [(5, 2), (10, 4), (16, 7), (30, 8), (100, 20)]:
        for h, g in gen_pairs(6, l, r, 1234, 0):
            lines += [h, g]
    for f in ("test.nw", "nets.nw", "test_tree.nw"):
        lines += [x.strip() for x in open("/root/reference/data/" + f) if x.strip()]
    lines += ["(a:1.5,(b:2,c:0.1e-3)x:3)root;", " ( a , b ) ; ", "((a,b)#H1,(#H1,c));", "(a,(b,c,d),e);", "((x y,z),w);", "(A,(B)#LGT12,(#LGT12,C));",
              "((a,b),(c,d))r;", "(((a)#H1,(#H1)#H2),#H2);"]
    out = run_lines(REF_TC, "parse", lines)
    blocks = [b.strip("\n").split("\n") for b in out.split("--\n")][:len(lines)]
    items = []
    for s, b in zip(lines, blocks):
        assert b[0] != "ERR", s
        n, m = map(int, b[0].split())
        edges = [list(map(int, x.split())) for x in b[1:1 + m]]
        labels = {}
        for x in b[1 + m:]:
            _, v, name = x.split(" ", 2)
            labels[v] = name
        items.append(dict(newick=s, num_nodes=n, edges=edges, labels=labels))
    json.dump(dict(source="oracle/_ref/ref_tc parse (reference io/newick.hpp)", items=items), open(os.path.join(GOLD, "newick_ref.json"), "w"), indent=0)
    print("newick_ref.json:", len(items))


def make_tc():
    pairs = [DATA_TEST] + FALSE_POSITIVES + [("((a,b),(c,d));", "((a,b),(c,d));"), ("((a,b),(c,d));", "((a,c),(b,d));")]
    tags = ["data/test.nw"] + ["survey_false_positive"] * 5 + ["tree_in_tree_pos", "tree_in_tree_neg"]
    for (l, r, cnt) in [(4, 1, 20), (6, 2, 40), (8, 3, 40), (10, 4, 40), (12, 5, 30), (16, 7, 30)]:
        for kind in (0, 1, 2):
            ps = gen_pairs(cnt, l, r, 7000 + 100 * l + r, kind)
            pairs += ps
            tags += ["gen l=%d r=%d kind=%d" % (l, r, kind)] * len(ps)
    lines = [x for p in pairs for x in p]
    ref = run_lines(REF_TC, "verdicts", lines).strip()
    assert len(ref) == len(pairs), (len(ref), len(pairs))
    items = []
    for (h, g), tag, rv in zip(pairs, tags, ref):
        truth, ndisp, first = O.tc_bruteforce_newick(h, g, stop_at_first=False)
        items.append(dict(host=h, guest=g, tag=tag, reference=rv, truth=int(truth), n_displaying=int(ndisp)))
    agree = sum(1 for it in items if it["reference"] in "01" and int(it["reference"]) == it["truth"])
    wrong = [i for i, it in enumerate(items) if it["reference"] in "01" and int(it["reference"]) != it["truth"]]
    crashed = sum(1 for it in items if it["reference"] in "ES")
    json.dump(dict(source="oracle/_ref/ref_tc verdicts (reference containment.hpp) + oracle/pt_oracle.c exhaustive switching",
                   summary=dict(instances=len(items), reference_agrees=agree, reference_wrong=wrong, reference_crashed=crashed), items=items),
              open(os.path.join(GOLD, "tc_ref_verdicts.json"), "w"), indent=0)
    print("tc_ref_verdicts.json:", len(items), "agree", agree, "wrong", wrong, "crashed", crashed)


def make_lca():
    rng = np.random.default_rng(2026)
    trees = []
    for n in (1, 2, 9, 33, 100, 1000, 4097):
        par = np.full(n, -1, dtype=np.int64)
        for v in range(1, n):
            par[v] = rng.integers(0, v)
        perm = rng.permutation(n)
        newpar = np.full(n, -1, dtype=np.int64)
        for v in range(n):
            newpar[perm[v]] = perm[par[v]] if par[v] >= 0 else -1
        q = 200
        u, v = rng.integers(0, n, q), rng.integers(0, n, q)
        inp = "%d %d\n%s\n%s\n" % (n, q, " ".join(map(str, newpar)), "\n".join("%d %d" % (a, b) for a, b in zip(u, v)))
        out = subprocess.run([REF_LCA, "lca"], input=inp, capture_output=True, text=True).stdout.split()
        trees.append(dict(parent=newpar.tolist(), u=u.tolist(), v=v.tolist(), lca=list(map(int, out))))
    # the tree of rmq/lca.cpp:24-43 with its 4 known answers (ids as assigned there: c0 d1 e2 b3 h4 g5 i6 f7 a8)
    known = dict(parent=[3, 3, 3, 8, 5, 7, 7, 8, -1], u=[8, 3, 0, 4], v=[8, 7, 2, 6], lca=[8, 8, 3, 7], source="rmq/lca.cpp:46-49")
    rmqs = []
    for n, mode in ((10, "rmq"), (1000, "rmq"), (5000, "rmq"), (1000, "pmrmq")):
        if mode == "rmq":
            a = rng.integers(0, 50, n)
        else:
            a = np.cumsum(np.concatenate([[0], rng.choice([-1, 1], n - 1)]))
        q = 300
        l = rng.integers(0, n, q)
        r = np.minimum(n, l + 1 + rng.integers(0, 100, q))
        inp = "%d %d\n%s\n%s\n" % (n, q, " ".join(map(str, a)), "\n".join("%d %d" % (x, y) for x, y in zip(l, r)))
        out = subprocess.run([REF_LCA, mode], input=inp, capture_output=True, text=True).stdout.split()
        rmqs.append(dict(mode=mode, values=a.tolist(), l=l.tolist(), r=r.tolist(), index=list(map(int, out[0::2])), value=list(map(int, out[1::2]))))
    json.dump(dict(source="oracle/_ref/ref_lca (reference rmq/lca.hpp, sparse_rmq.hpp, pm_rmq.hpp)", trees=trees, known=known, rmq=rmqs),
              open(os.path.join(GOLD, "lca_ref.json"), "w"), indent=0)
    print("lca_ref.json:", len(trees), "trees,", len(rmqs), "rmq arrays")


def _rand_tree_newick(rng, labels, max_deg):
    """random rooted tree on the given leaf labels, inner nodes with 2..max_deg children"""
    nodes = list(labels)
    rng.shuffle(nodes)
    while len(nodes) > 1:
        k = int(min(len(nodes), rng.integers(2, max_deg + 1)))
        idx = sorted(rng.choice(len(nodes), k, replace=False).tolist(), reverse=True)
        kids = [nodes.pop(i) for i in idx]
        nodes.insert(int(rng.integers(0, len(nodes) + 1)), "(" + ",".join(kids) + ")")
    return nodes[0] + ";"


def make_who():
    """tree/tree pairs -> who_displays(u) of the REFERENCE for every guest node (oracle/_ref/ref_tc who)"""
    rng = np.random.default_rng(4711)
    pairs = [("((a,b),(c,d));", "((a,b),(c,d));"), ("((a,b),(c,d));", "((a,c),(b,d));"), ("(((a,b),c),(d,e));", "((a,b),((d,e),c));")]
    names = [pt_name(i) for i in range(40)]
    for trial in range(160):
        l = int(rng.integers(3, 14))
        labs = names[:l]
        max_deg = 2 if trial % 3 else 4
        host = _rand_tree_newick(rng, labs, max_deg)
        mode = trial % 5
        if mode == 0:
            guest = host                                     # identical
        elif mode == 1:
            guest = _rand_tree_newick(rng, labs, max_deg)    # random tree on the same labels
        elif mode == 2:
            guest = _rand_tree_newick(rng, labs[: max(2, l - 2)], max_deg)   # host has extra labels
        elif mode == 3:
            guest = _rand_tree_newick(rng, labs + ["zz"], max_deg)           # guest has a label the host lacks
        else:                                                # host with one subtree re-hung: partly displayed
            guest = _rand_tree_newick(rng, labs, 2)
            host = guest.replace("(" + labs[0] + ",", "(" + labs[1] + ",", 1) if False else host
        pairs.append((host, guest))
    out = run_lines(REF_TC, "who", [x for p in pairs for x in p])
    blocks = [b.strip("\n").split("\n") for b in out.split("--\n")][:len(pairs)]
    items = []
    for (h, g), b in zip(pairs, blocks):
        if b and b[0] in ("CRASH", "EXC"):
            items.append(dict(host=h, guest=g, reference=b[0]))
            continue
        who = []
        for line in b:
            u, rest = line.split(":")
            who.append([int(x) for x in rest.split()])
        items.append(dict(host=h, guest=g, reference=who))
    crashed = sum(1 for it in items if not isinstance(it["reference"], list))
    json.dump(dict(source="oracle/_ref/ref_tc who (reference TreeInTreeContainment::who_displays, containment.hpp:72)", crashed=crashed, items=items),
              open(os.path.join(GOLD, "who_ref.json"), "w"), indent=0)
    print("who_ref.json:", len(items), "pairs,", crashed, "reference crashes")


def make_bridges():
    """networks -> the bridge set of the REFERENCE BridgeFinder (oracle/_ref/ref_tc bridges)"""
    lines = ["((a,(((b)#H1,(c)#H2),(#H1,#H2))),d);", "((a,b),(c,d));", "((a,(b)#H1),(#H1,c));"]
    for (l, r, cnt) in [(5, 2, 15), (10, 4, 15), (20, 8, 10), (50, 10, 10), (100, 20, 6)]:
        lines += [h for h, _ in gen_pairs(cnt, l, r, 9100 + l, 0)]
    out = run_lines(REF_TC, "bridges", lines)
    blocks = [b.strip("\n").split("\n") for b in out.split("--\n")][:len(lines)]
    items = []
    for s, b in zip(lines, blocks):
        if b and b[0] in ("CRASH", "EXC"):
            items.append(dict(newick=s, reference=b[0]))
        else:
            items.append(dict(newick=s, reference=[list(map(int, x.split())) for x in b if x.strip()]))
    crashed = sum(1 for it in items if not isinstance(it["reference"], list))
    json.dump(dict(source="oracle/_ref/ref_tc bridges (reference utils/bridges.hpp BridgeFinder)", crashed=crashed, items=items),
              open(os.path.join(GOLD, "bridges_ref.json"), "w"), indent=0)
    print("bridges_ref.json:", len(items), "networks,", crashed, "reference crashes")


def make_bcc():
    """networks -> bridges in the reference's post-order and the components its BiconnectedComponents iterator yields
    (oracle/_ref/ref_tc bcc): pins the ORDER that pt/bridges.hpp reproduces on top of pt_net_bridges"""
    lines = ["((a,(((b)#H1,(c)#H2),(#H1,#H2))),d);", "((a,b),(c,d));", "((a,(b)#H1),(#H1,c));", "(((a,b),((c)#H1,d)),(#H1,e));"]
    for (l, r, cnt) in [(5, 2, 12), (10, 4, 12), (20, 8, 8), (50, 10, 6)]:
        lines += [h for h, _ in gen_pairs(cnt, l, r, 9300 + l, 0)]
    out = run_lines(REF_TC, "bcc", lines)
    blocks = [b.strip("\n").split("\n") for b in out.split("--\n")][:len(lines)]
    items = []
    for s, b in zip(lines, blocks):
        if any(x in ("CRASH", "EXC") for x in b):
            items.append(dict(newick=s, reference="CRASH"))
            continue
        bridges = [list(map(int, x.split()[1:])) for x in b if x.startswith("B ")]
        comps = [[list(map(int, e.split(","))) for e in x.split()[1:]] for x in b if x.startswith("C")]
        items.append(dict(newick=s, bridges=bridges, components=comps))
    crashed = sum(1 for it in items if it.get("reference") == "CRASH")
    json.dump(dict(source="oracle/_ref/ref_tc bcc (reference utils/bridges.hpp + utils/biconnected_comps.hpp)", crashed=crashed, items=items),
              open(os.path.join(GOLD, "bcc_ref.json"), "w"), indent=0)
    print("bcc_ref.json:", len(items), "networks,", crashed, "reference crashes")


def make_iso():
    """pairs of networks -> verdict of the REFERENCE IsomorphismMapper (oracle/_ref/ref_tc iso, default flags = leaf labels) and the
    truth by definition (oracle/pt_oracle.c orc_iso).  Variants of generated networks: children / hybrid numbering shuffled
    (isomorphic), two leaf labels swapped, one reticulation arc moved, and unrelated networks of the same size."""
    import random
    from oracle.newick import parse_newick, write_newick
    rnd = random.Random(20261020)
    pairs = [("(a,(((b)#H1,(c)#H2),(#H1,#H2)),d);", "(d,(((b)#H1,(c)#H2),(#H1,#H2)),a);", "data/nets.nw"),
             ("((a,b),(c,d));", "((d,c),(b,a));", "tree mirrored"), ("((a,b),(c,d));", "((a,c),(b,d));", "tree, other pairs"),
             ("((a,(b)#H1),(#H1,c));", "((c,(b)#H1),(#H1,a));", "one reticulation mirrored"), ("((a,(b)#H1),(#H1,c));", "((b,(a)#H1),(#H1,c));", "leaf under the reticulation changed")]
    for (l, r, cnt) in [(4, 1, 6), (6, 2, 10), (9, 3, 10), (12, 4, 8), (16, 6, 6)]:
        nets = [h for h, _ in gen_pairs(cnt + 1, l, r, 9500 + l, 0)]
        for i in range(cnt):
            n, e, lab = parse_newick(nets[i])
            e2 = list(e); rnd.shuffle(e2)
            pairs.append((nets[i], write_newick(n, e2, lab), "gen l=%d r=%d shuffled" % (l, r)))
            leaves = sorted(v for v in lab if lab[v] != "")
            a, b = rnd.sample(leaves, 2)
            lab2 = dict(lab); lab2[a], lab2[b] = lab[b], lab[a]
            pairs.append((nets[i], write_newick(n, e2, lab2), "gen l=%d r=%d two labels swapped" % (l, r)))
            pairs.append((nets[i], nets[i + 1], "gen l=%d r=%d unrelated" % (l, r)))
    lines = [x for a, b, _ in pairs for x in (a, b)]
    ref = run_lines(REF_TC, "iso", lines).strip()
    assert len(ref) == len(pairs), (len(ref), len(pairs))
    items = []
    for (a, b, tag), rv in zip(pairs, ref):
        items.append(dict(a=a, b=b, tag=tag, reference=rv, truth=O.iso_newick(a, b)))
    agree = sum(1 for it in items if it["reference"] == str(it["truth"]))
    json.dump(dict(source="oracle/_ref/ref_tc iso (reference utils/isomorphism.hpp, flags = FLAG_MAP_LEAF_LABELS) + oracle/pt_oracle.c orc_iso",
                   reference_agrees=agree, items=items), open(os.path.join(GOLD, "iso_ref.json"), "w"), indent=0)
    print("iso_ref.json:", len(items), "pairs,", agree, "where the reference agrees with the definition;", sum(it["truth"] for it in items), "isomorphic")


def pt_name(i):
    s = ""
    while True:
        s = chr(ord("a") + i % 26) + s
        if i < 26:
            return s
        i //= 26


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    make_newick()
    make_tc()
    make_lca()
    make_who()
    make_bridges()
    make_bcc()
    make_iso()
