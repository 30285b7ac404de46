"""Host-side logic and the C-ABI surface (no GPU): the product's C++ Newick reader against the reference reader's output,
the packer, the generator, and that libptgpu.so exports every symbol include/pt_gpu.h declares."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_every_declared_symbol(pt):
    hdr = open(os.path.join(ROOT, "include", "pt_gpu.h")).read()
    declared = sorted(set(re.findall(r"^(?:const char\*|int64_t|int)\s+(pt_[a-z0-9_]+)\s*\(", hdr, flags=re.M)))
    assert len(declared) >= 24
    lib = C.CDLL(os.path.join(ROOT, "phylo_tools_b200", "libptgpu.so"))
    for name in declared:
        assert hasattr(lib, name), name
    assert pt.lib().pt_abi_version() == 3


def test_cpp_newick_reader_matches_reference(pt, gold_newick):
    for it in gold_newick["items"]:
        n, edges, labels = pt.parse_newick(it["newick"])
        assert n == it["num_nodes"]
        assert [list(e) for e in edges] == it["edges"]
        assert {str(k): v for k, v in labels.items()} == it["labels"]


def test_malformed_newick(pt):
    for bad in ["((a,b),(c,d))", "(a,b)x#H;", "((a)#H1,(#H1,#H1));", "(a,b));"]:
        with pytest.raises(pt.MalformedNewick):
            pt.parse_newick(bad)
    # like the reference, the reader stops once the root's subtree is complete: an unmatched '(' on the left is ignored
    assert pt.parse_newick("((a,b),(c,d);") == (3, [(0, 1), (0, 2)], {1: "d", 2: "c"})


def test_packer_compact_arrays(pt):
    """index_bytes = 2: the same batch with int16 data arrays; refused when an instance needs wider ids"""
    p = pt.Packer()
    p.add_generated(20, 30, 5, 77, 0)
    a, c = p.arrays(), p.arrays(compact=True)
    assert c["index_bytes"] == 2 and a["index_bytes"] == 4
    for k in ("host_succ_off", "host_succ_adj", "host_label", "guest_parent", "guest_label"):
        assert c[k].dtype == np.int16 and (c[k].astype(np.int32) == a[k]).all()
    for k in ("host_node_off", "host_arc_off", "guest_node_off"):
        assert (c[k] == a[k]).all()
    big = pt.Packer()
    big.add_generated(1, 20000, 10, 5, 0)   # 40 019 host nodes
    with pytest.raises(pt.PtError):
        big.arrays(compact=True)


def test_packer_layout_and_roundtrip(pt, oracle):
    p = pt.Packer()
    p.add_generated(5, 10, 3, 99, 0)
    p.add_newick("((a,(((b)#H1,(c)#H2),(#H1,#H2))),d);", "((a,b),(c,d));")
    p.add_newick("((a,b),(c,d));", "(((a)#H1,b),(#H1,c));")  # tree first, network second: the network becomes the host (tc.cpp:110-111)
    a = p.arrays()
    B = a["num_instances"]
    assert B == 7
    for i in range(B):
        hb, he = a["host_node_off"][i], a["host_node_off"][i + 1]
        n, m = he - hb, a["host_arc_off"][i + 1] - a["host_arc_off"][i]
        off = a["host_succ_off"][hb + i: hb + i + n + 1]
        assert off[0] == 0 and off[-1] == m and (np.diff(off) >= 0).all()
        adj = a["host_succ_adj"][a["host_arc_off"][i]: a["host_arc_off"][i + 1]]
        assert ((adj >= 0) & (adj < n)).all()
        gb, ge = a["guest_node_off"][i], a["guest_node_off"][i + 1]
        par = a["guest_parent"][gb:ge]
        assert (par == -1).sum() == 1 and (par < ge - gb).all()
        # pred arrays are the transpose of succ
        poff = a["host_pred_off"][hb + i: hb + i + n + 1]
        assert poff[-1] == m
        # the Newick text rebuilt from the flat arrays describes the same instance: same oracle verdict as the arrays
        h, g = p.newick(i)
        tails = np.repeat(np.arange(n), np.diff(off))
        edges = list(zip(tails.tolist(), adj.tolist()))
        gedges = [(int(par[t]), t) for t in range(ge - gb) if par[t] >= 0]
        v1 = oracle.tc_bruteforce(edges, n, a["host_label"][hb:he], gedges, ge - gb, a["guest_label"][gb:ge])[0]
        v2 = oracle.tc_bruteforce_newick(h, g)[0]
        assert v1 == v2
    assert a["max_host_nodes"] == max(np.diff(a["host_node_off"]))
    # instance 6: host must be the network (4 leaves a b c + ... ) although it came second
    assert a["host_arc_off"][7] - a["host_arc_off"][6] == 7


def test_generator_shapes(pt):
    # utils/generator.hpp:19-23: n = 2r + 2l - 1, m = n - 1 + r; binary; positives are displayed by construction
    p = pt.Packer()
    p.add_generated(20, 100, 20, 0xB200, 0)
    a = p.arrays()
    assert (np.diff(a["host_node_off"]) == 239).all() and (np.diff(a["host_arc_off"]) == 258).all()
    assert (np.diff(a["guest_node_off"]) == 199).all()
    for i in range(20):
        hb = a["host_node_off"][i]
        off = a["host_succ_off"][hb + i: hb + i + 240]
        deg = np.diff(off)
        assert set(deg.tolist()) <= {0, 1, 2} and (deg == 0).sum() == 100 and (deg == 1).sum() == 20
        lab = a["host_label"][hb:hb + 239]
        assert sorted(lab[lab >= 0].tolist()) == list(range(100))
    # same seed -> same batch
    q = pt.Packer()
    q.add_generated(20, 100, 20, 0xB200, 0)
    b = q.arrays()
    assert all((a[k] == b[k]).all() for k in ("host_succ_adj", "guest_parent", "guest_label"))
    h, g = p.newick(0)
    assert h.count("#H") == 40 and g.count(",") == 99


def test_sequential_taxon_names(pt):
    # utils/generator.hpp:31-46: a..z, ba, bb, ...
    p = pt.Packer()
    p.add_generated(1, 30, 0, 1, 0)
    h, _ = p.newick(0)
    names = set(re.findall(r"[a-z]+", h))
    assert {"a", "z", "ba", "bd"} <= names and "aa" not in names


def test_random_tree(pt):
    par = pt.random_tree(1000, 7)
    assert len(par) == 1999 and (par == -1).sum() == 1
    deg = np.bincount(par[par >= 0], minlength=1999)
    assert set(deg.tolist()) == {0, 2} and (deg[:1000] == 0).all()


def test_no_cpu_fallback(pt):
    """without a GPU every compute entry point must fail loudly with PT_ERR_NO_DEVICE"""
    if pt.device_count() > 0:
        pytest.skip("a GPU is present")
    p = pt.Packer()
    p.add_generated(2, 5, 1, 1, 0)
    with pytest.raises(pt.PtError) as e:
        pt.Batch(p.arrays())
    assert e.value.code == -4
    with pytest.raises(pt.PtError) as e:
        pt.Lca(np.array([-1, 0, 0], dtype=np.int32))
    assert e.value.code == -4
    with pytest.raises(pt.PtError) as e:
        pt.Rmq([3, 1, 2])
    assert e.value.code == -4


def test_batch_reader_is_bit_identical_to_the_per_instance_reader(pt, gold_tc, gold_newick):
    """pt_packer_add_newick_text (parallel, straight into flat arrays) against pt_packer_add_newick on the same lines:
    every array equal, for the pinned instances, generated ones, length annotations and white space"""
    pairs = [(it["host"], it["guest"]) for it in gold_tc["items"]]
    p = pt.Packer()
    p.add_generated(300, 40, 9, 77, 0)
    pairs += [p.newick(i) for i in range(300)]
    pairs += [("((a:1.5,b:2)x:3,(c,d)y);", "((a,b),(c,d));"), (" ( a , b ) ; ", "(b,a);"), ("((a,b),(c,d));", "(((a)#H1,b),(#H1,(c,d)));")]
    slow, fast = pt.Packer(), pt.Packer()
    for h, g in pairs:
        slow.add_newick(h, g)
    text = "\n\n".join(h + "\r\n" + g for h, g in pairs) + "\n"
    assert fast.add_newick_text(text, threads=3) == len(pairs)
    a, b = slow.arrays(), fast.arrays()
    for k in a:
        assert (a[k] == b[k]).all() if isinstance(a[k], np.ndarray) else a[k] == b[k], k
    one = pt.Packer()
    assert one.add_newick_text(text, threads=1) == len(pairs)
    c = one.arrays()
    assert all((a[k] == c[k]).all() for k in a if isinstance(a[k], np.ndarray))
    # a malformed pair is reported by index and nothing is added
    bad = pt.Packer()
    with pytest.raises(pt.MalformedNewick) as e:
        bad.add_newick_text("(a,b);\n(a,b);\n(a,b));\n(a,b);\n")
    assert "pair 1" in str(e.value) and len(bad) == 0
    with pytest.raises(pt.PtError):
        bad.add_newick_text("(a,b);\n((a)#H1,(#H1,b));\n((a)#H1,(#H1,b));\n((a)#H1,(#H1,b));\n")   # second pair: guest is not a tree


def test_iso_arrays_layout(pt):
    """pt.iso_arrays: the flat layout of pt_iso_desc (graph g = 2 * pair + side; succ / pred CSR with local ids)"""
    a = pt.iso_arrays([(3, [(0, 1), (0, 2)], [-1, 0, 1]), (4, [(0, 3), (0, 1), (1, 2)], [-1, -1, 1, 0])])
    assert a["num_pairs"] == 1 and a["max_nodes"] == 4
    assert a["node_off"].tolist() == [0, 3, 7] and a["arc_off"].tolist() == [0, 2, 5]
    assert a["succ_off"].tolist() == [0, 2, 2, 2, 0, 2, 3, 3, 3] and a["succ_adj"].tolist() == [1, 2, 3, 1, 2]
    assert a["pred_off"].tolist() == [0, 0, 1, 2, 0, 0, 1, 2, 3] and a["pred_adj"].tolist() == [0, 0, 0, 1, 0]
    assert a["label"].tolist() == [-1, 0, 1, -1, -1, 1, 0]


def test_edgelist_reader_equals_reference(pt, tmp_path):
    """parse_edgelist (io/edgelist.hpp:41-86): ids by first appearance, EVERY node labelled with its name, the reference's own
    acceptance of blanks / tabs / empty lines and its MalformedEdgeVec cases -- tests/golden/edgelist_ref.json, produced by the
    reference reader itself (oracle/_ref/ref_tc parse_el); then read_edgelists (io/io.hpp:52-62) through the `tc` tool's reader:
    an edge list labels its inner nodes too, which is why the reference's tc cannot use it"""
    import pytest
    from conftest import load_golden
    gold = load_golden("edgelist_ref.json")
    ok = bad = 0
    for it in gold["items"]:
        if it.get("reference") == "ERR":
            with pytest.raises(pt.MalformedEdgeVec):
                pt.parse_edgelist(it["text"])
            bad += 1
            continue
        n, edges, labels = pt.parse_edgelist(it["text"])
        assert n == it["num_nodes"], it["text"]
        assert [list(e) for e in edges] == it["edges"], it["text"]
        assert {str(k): v for k, v in labels.items()} == it["labels"], it["text"]
        assert len(labels) == n      # every node carries its name
        ok += 1
    assert ok >= 12 and bad >= 3
