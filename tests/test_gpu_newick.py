"""pt_batch_create_from_newick (extended-Newick text parsed ON THE GPU into a device-resident batch) against the host batch reader
(pt_packer_add_newick_text = the reference reader's numbering, pinned by tests/golden/newick_ref.json): the arrays must be
bit-identical, the verdicts the pinned truths, malformed input must name the first bad pair."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

# PT_STRESS=<k>: k more seeds for the two randomised reader tests (a one-off soak, not part of the regular run)
EXTRA_SEEDS = [9000 + i for i in range(int(os.environ.get("PT_STRESS", "0") or 0))]

KEYS = ["host_node_off", "host_arc_off", "guest_node_off", "host_succ_off", "host_succ_adj", "host_label", "guest_parent", "guest_label"]


def _same_arrays(pt, text):
    b = pt.Batch.from_newick_text(text)
    d = b.download()
    p = pt.Packer()
    k = p.add_newick_text(text)
    assert k == b.num_instances
    a = p.arrays(compact=d["index_bytes"] == 2)
    for key in KEYS:
        assert d[key].dtype == a[key].dtype and d[key].shape == a[key].shape and (d[key] == a[key]).all(), key
    for key in ("max_host_nodes", "max_host_arcs", "max_guest_nodes", "max_labels"):
        assert d[key] == max(a[key], 1), key
    return b, p


def test_device_reader_matches_host_reader_on_golden_instances(pt, gold_tc):
    items = gold_tc["items"]
    text = "".join(it["host"] + "\n" + it["guest"] + "\n" for it in items)
    b, _ = _same_arrays(pt, text)
    v, s, _ = b.contain()
    b.free()
    assert (s == 0).all() and [int(x) for x in v] == [it["truth"] for it in items]


def test_device_reader_on_generated_batches_and_formatting(pt):
    p = pt.Packer()
    p.add_generated(3000, 100, 20, 0xB200, 0); p.add_generated(500, 30, 6, 99, 1); p.add_generated(500, 12, 3, 98, 2)
    pairs = [p.newick(i) for i in range(len(p))]
    text = "".join(h + "\n" + g + "\n" for h, g in pairs)
    b, _ = _same_arrays(pt, text)
    v, s, _ = b.contain()
    b.free()
    pv, ps, _ = pt.Batch(p.arrays()).contain()
    assert (s == 0).all() and (v == pv).all() and v[:3000].all()
    # blank lines, CRLF, indentation, tree first (the network becomes the host), branch lengths, no newline at the end
    odd = ("\n\n  ((a:1.5,b:2):0.1,(c,d):3e-2);\r\n\t((a,(b)#H1:0.5),(#H1,(c,d)));   \n\n"
           "((x,(y)#H7),(#H7,z));\n((x,y),z);\n   \n"
           "(a,b,c);\n(c,b,a);")
    b2, _ = _same_arrays(pt, odd)
    v2, s2, _ = b2.contain()
    b2.free()
    assert b2.num_instances == 3 and (s2 == 0).all() and v2.tolist() == [True, True, True]
    # an odd number of lines: the last one is ignored, like the host reader
    b3, _ = _same_arrays(pt, "((a,b),c);\n((a,b),c);\n(a,b);\n")
    assert b3.num_instances == 1
    b3.free()
    # empty text
    b4 = pt.Batch.from_newick_text("")
    assert b4.num_instances == 0
    v4, s4, _ = b4.contain()
    assert len(v4) == 0
    b4.free()


def test_device_reader_text_already_on_the_device(pt):
    import torch
    p = pt.Packer()
    p.add_generated(2000, 50, 10, 4711, 0)
    text = "".join(h + "\n" + g + "\n" for h, g in (p.newick(i) for i in range(len(p)))).encode()
    t = torch.frombuffer(bytearray(text), dtype=torch.uint8).cuda()
    b = pt.Batch.from_newick_text(t)
    v, s, _ = b.contain()
    b.free()
    assert b.num_instances == 2000 and (s == 0).all() and v.all()


def test_device_reader_reports_the_first_bad_pair(pt):
    good = "((a,b),(c,d));\n((a,b),(c,d));\n"
    for bad, what in [("((a,b),(c,d))\n((a,b),(c,d));\n", "expected ';'"), ("((a,b)#H,(c,d));\n((a,b),(c,d));\n", "hybrid number"),
                      ("((a,b),(c,d));\n(a,b));\n", "unexpected start"), ("((a)#H1,(#H1,#H1));\n(a,b);\n", "double edge")]:
        with pytest.raises(pt.MalformedNewick) as ei:
            pt.Batch.from_newick_text(good * 5 + bad + good * 3)
        assert "pair 5" in str(ei.value) and what in str(ei.value), str(ei.value)
    with pytest.raises(pt.PtError) as ei:   # two networks: the guest has to be a tree (examples/tc.cpp:135-137)
        pt.Batch.from_newick_text(good + "((a,(b)#H1),(#H1,c));\n((a,(b)#H1),(#H1,c));\n")
    assert "pair 1" in str(ei.value) and "guest must be a tree" in str(ei.value)


@pytest.mark.parametrize("seed", [4242] + EXTRA_SEEDS)
def test_device_reader_fuzz_against_host_reader(pt, seed):
    """mutated Newick lines (characters deleted, duplicated, replaced by structure characters): the device reader must accept
    exactly what the host reader accepts, with identical arrays, and reject the rest -- never crash or hang"""
    import random
    rnd = random.Random(seed)
    p = pt.Packer()
    p.add_generated(60, 8, 3, 321 + seed % 1000, 0)
    base = [p.newick(i) for i in range(len(p))]
    alphabet = "(),;#H:0123456789ab \t"
    ok = bad = 0
    for it in range(1500):
        h, g = base[it % len(base)]
        s = list(h if it % 2 else g)
        for _ in range(rnd.randint(1, 3)):
            k = rnd.randrange(len(s))
            op = rnd.randrange(3)
            if op == 0:
                del s[k]
            elif op == 1:
                s.insert(k, rnd.choice(alphabet))
            else:
                s[k] = rnd.choice(alphabet)
        m = "".join(s).replace("\n", "")
        text = (m + "\n" + g + "\n") if it % 2 else (h + "\n" + m + "\n")
        host_err = dev_err = None
        q = pt.Packer()
        try:
            q.add_newick_text(text)
        except Exception as ex:
            host_err = ex
        try:
            b = pt.Batch.from_newick_text(text)
        except Exception as ex:
            dev_err = ex
        assert (host_err is None) == (dev_err is None), (text, host_err, dev_err)
        if host_err is None:
            d = b.download()
            a = q.arrays(compact=d["index_bytes"] == 2)
            for key in KEYS:
                assert (d[key] == a[key]).all(), (key, text)
            b.free()
            ok += 1
        else:
            assert type(host_err).__name__ == type(dev_err).__name__, (text, host_err, dev_err)
            bad += 1
    assert ok > 100 and bad > 100


def test_device_reader_short_swapped_pairs_and_repeated_labels(pt):
    """the guest line first (so the host is the SECOND line), lines shorter than a warp, and labels repeated inside one line (MUL-trees):
    label ids still come out in the host reader's first-appearance order"""
    import random
    rnd = random.Random(5)
    lines = []
    for _ in range(400):
        k = rnd.randint(2, 40)
        names = [rnd.choice("abcdefgh") + str(rnd.randint(0, 6)) for _ in range(k)]
        def tree(xs):
            if len(xs) == 1:
                return xs[0]
            c = rnd.randint(1, len(xs) - 1)
            return "(" + tree(xs[:c]) + "," + tree(xs[c:]) + ")"
        t = tree(names) + ";"
        n = "((" + tree(names) + ")#H1,(#H1," + rnd.choice(names) + "));"
        lines += [t, n] if rnd.random() < 0.5 else [n, t]
    b, _ = _same_arrays(pt, "\n".join(lines) + "\n")
    b.free()


@pytest.mark.parametrize("seed", [77] + EXTRA_SEEDS)
def test_device_reader_well_formed_lines_with_rich_formatting(pt, seed):
    """lines the warp-parallel path of nw_parse_kernel takes (well-formed), but with everything the sequential rules look at:
    blanks around names and after '(', branch lengths in several spellings, names made of digits / '.' / 'e' / '+' (which look like a
    length until the ':' is missing), unnamed inner nodes and leaves, hybrids given two or three times with and without names and
    lengths, high degrees, deep chains -- arrays bit-identical to the host reader's"""
    import random
    rnd = random.Random(seed)
    alphabet = ["a", "b", "x1", "12", "3e", "e5", "t.1", "L+", "n-2", "Q", "7", "ab_c", "E", "0.5"]

    def blanks():
        return rnd.choice(["", "", "", " ", "  ", "\t"])

    def length():
        return rnd.choice(["", "", ":1", ":0.5", ":1e-3", ":2.5E+2", ":", ":-1", ":.5"])

    lines = []
    for it in range(600):
        k = rnd.randint(2, 60)
        names = ["%s%d" % (rnd.choice(alphabet), i) if rnd.random() < 0.8 else "%d" % (1000 + i) for i in range(k)]   # distinct

        def build(xs, depth):
            if len(xs) == 1:
                return xs[0]
            if depth > 40 or rnd.random() < 0.15:          # high degree
                return [build([x], depth + 1) for x in xs]
            c = rnd.randint(1, len(xs) - 1) if rnd.random() < 0.7 else 1   # c = 1: deep chains
            return [build(xs[:c], depth + 1), build(xs[c:], depth + 1)]
        tree = build(names, 0)
        inner = []

        def collect(t):
            if isinstance(t, list):
                inner.append(t)
                for c in t:
                    collect(c)
        collect(tree)
        nh = rnd.randint(0, 4)
        hyb = {}
        for h in range(nh):   # a new leaf below a hybrid node that hangs under two or three different inner nodes
            key = rnd.choice([h + 1, 10 * h + 7, 100 + h])
            if key in hyb or len(inner) < 3:
                continue
            homes = rnd.sample(inner, 3 if rnd.random() < 0.3 else 2)
            hname = rnd.choice(["", "h", "r%d" % h])
            hyb[key] = True
            first = True
            for home in reversed(homes):   # the reader meets the LAST one in the line first, but any of them may carry the child
                home.append(("hyb", key, hname, "w%d" % h if first else None))
                first = False

        def render(t, net):
            if isinstance(t, tuple):
                _, key, hname, child = t
                if not net:
                    return None
                tag = "%s#%s%d" % (hname, rnd.choice(["H", "LGT", "R"]), key)
                return ("(" + blanks() + child + length() + ")" if child else "") + tag + length()
            if isinstance(t, str):
                return blanks() + t + length() + blanks() if rnd.random() < 0.5 else t + length()
            kids = [r for r in (render(c, net) for c in t) if r is not None]
            if len(kids) == 1 and not net:
                return kids[0]
            name = rnd.choice(["", "", "i%d" % rnd.randint(0, 9)])
            return "(" + blanks() + ",".join(kids) + ")" + name + length() + blanks()
        net = render(tree, True).strip()
        # the root's token sits right of the last ')' with no ',' or ')' after it: no length is skipped there, keep it plain
        net = net[:net.rindex(")") + 1] + rnd.choice(["", "root"]) + ";"
        gt = render(tree, False).strip()
        gt = gt[:gt.rindex(")") + 1] + ";"
        lines += [net, gt]
    text = "\n".join(lines) + "\n"
    b, _ = _same_arrays(pt, text)
    assert b.num_instances == 600
    b.free()
