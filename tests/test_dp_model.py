"""The device algorithm of the general tree-in-tree DP (tests/dp_model.py: dataflow order, closed form for single candidates, stack
sweep + W-word child sets + Kuhn matching) against the definition-based oracle and the reference's own lists -- on the CPU."""
import numpy as np

import dp_model
from conftest import load_golden


def _lists(off, adj):
    return [sorted(adj[off[u]:off[u + 1]].tolist()) for u in range(len(off) - 1)]


def _check_gold(oracle, name, min_checked):
    gold = load_golden(name)
    checked = general = 0
    for it in gold["items"]:
        hp, hl, gp, gl = oracle.tree_arrays_from_newick(it["host"], it["guest"])
        got = dp_model.who_lists(list(hp), list(hl), list(gp), list(gl))
        # ascending in host preorder as the kernel appends them (an antichain popped from the stack)
        H = dp_model.Host(list(hp))
        for lst in got:
            assert [H.pos[v] for v in lst] == sorted(H.pos[v] for v in lst)
        got = [sorted(x) for x in got]
        ooff, oadj = oracle.mul_who_displays(hp, hl, gp, gl)
        assert got == _lists(ooff, oadj), (it["host"], it["guest"])
        if it["reference"] != "CRASH":
            assert got == [sorted(r) for r in it["reference"]], (it["host"], it["guest"])
            checked += 1
        general += sum(len(x) > 1 for x in got)
    assert checked >= min_checked
    return general


def test_model_equals_reference_on_multilabelled_hosts(oracle):
    assert _check_gold(oracle, "whomul_ref.json", 150) >= 100


def test_model_equals_reference_on_wide_nodes(oracle):
    """nodes of 65 ... 320 children: more than one 64-bit word per child set, the sorted closed form beyond 8 children"""
    assert _check_gold(oracle, "whowide_ref.json", 28) >= 500


def test_matching_needs_augmenting_paths():
    """3 guest children, 3 induced children a, b, c: child 0 fits {a, b}, child 1 fits {a}, child 2 fits {b, c} -- the greedy choice
    0 -> a has to be undone by an augmenting path"""
    cov = [[0b011], [0b101], [0b100]]       # a: children 0, 1   b: children 0, 2   c: child 2
    assert dp_model.perfect_matching(3, 3, 1, cov, [0, 1, 2])
    assert not dp_model.perfect_matching(3, 3, 1, [[0b011], [0b011], [0b000]], [0, 1, 2])   # child 2 fits nowhere
    assert not dp_model.perfect_matching(3, 2, 1, cov, [0, 1])                               # fewer induced children than guest children

    def m(*bits):           # a 2-word child set
        x = 0
        for b in bits:
            x |= 1 << b
        return [x & ((1 << 64) - 1), x >> 64]
    # across the word boundary (66 guest children, W = 2): children 63, 64, 65 compete for three induced children
    ok = [m(63, 64), m(63, 65), m(65)] + [m(i) for i in range(63)]       # 64 -> first, 63 -> second, 65 -> third
    assert dp_model.perfect_matching(66, 66, 2, ok, list(range(66)))
    bad = [m(63, 64), m(64), m(65)] + [m(i) for i in range(62)] + [m(62, 63)]   # fine: 64 -> second, 63 -> first or last
    assert dp_model.perfect_matching(66, 66, 2, bad, list(range(66)))
    none = [m(63), m(63), m(65)] + [m(i) for i in range(63)]              # child 64 fits nowhere: the coverage test already says no
    assert not dp_model.perfect_matching(66, 66, 2, none, list(range(66)))
    hall = [m(63, 64), m(65), m(65)] + [m(i) for i in range(63)]          # covered, but 63 and 64 share ONE induced child (Hall)
    assert not dp_model.perfect_matching(66, 66, 2, hall, list(range(66)))


def test_model_on_random_pairs(oracle):
    rng = np.random.default_rng(77)

    def tree(labels, deg):
        parent = [-1] * len(labels)
        label = list(labels)
        roots = list(range(len(labels)))
        while len(roots) > 1:
            k = int(min(len(roots), rng.integers(2, deg + 1)))
            pick = sorted(rng.choice(len(roots), k, replace=False).tolist(), reverse=True)
            new = len(parent)
            parent.append(-1); label.append(-1)
            for i in pick:
                parent[roots.pop(i)] = new
            roots.append(new)
        # renumber so that the root is node 0 (parents before children is not required)
        return parent, label
    multi = 0
    for trial in range(120):
        l = int(rng.integers(3, 14))
        mult = int(rng.integers(1, 4))
        host_labels = list(range(l)) + [int(rng.integers(0, l)) for _ in range(int(rng.integers(0, mult * l + 1)))]
        rng.shuffle(host_labels)
        hp, hl = tree(host_labels, int(rng.integers(2, 6)))
        gp, gl = tree(list(rng.permutation(l)), int(rng.integers(2, 6)))
        got = [sorted(x) for x in dp_model.who_lists(hp, hl, gp, gl)]
        ooff, oadj = oracle.mul_who_displays(hp, hl, gp, gl)
        assert got == _lists(ooff, oadj), (hp, hl, gp, gl)
        multi += sum(len(x) > 1 for x in got)
    assert multi >= 50
