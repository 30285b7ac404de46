"""Parity of the CUDA LCA / RMQ path (through the C-ABI) with the oracle and with the reference's own answers."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _shape(rng, n, kind):
    par = np.full(n, -1, dtype=np.int64)
    idx = np.arange(1, n)
    if kind == "random":
        par[1:] = (rng.random(n - 1) * idx).astype(np.int64)
    elif kind == "path":
        par[1:] = idx - 1
    elif kind == "star":
        par[1:] = 0
    elif kind == "cat":
        par[1:] = np.where(idx % 2 == 1, idx - 1, np.maximum(idx - 2, 0))
    elif kind == "bushy":  # ~8 children per node
        par[1:] = (idx - 1) // 8
    perm = rng.permutation(n)
    out = np.full(n, -1, dtype=np.int32)
    out[perm] = np.where(par >= 0, perm[np.maximum(par, 0)], -1)
    return out


def test_reference_answers(pt, gold_lca):
    k = gold_lca["known"]  # rmq/lca.cpp:46-49
    L = pt.Lca(np.array(k["parent"], dtype=np.int32))
    assert L.query(k["u"], k["v"]).tolist() == k["lca"]
    L.free()
    for t in gold_lca["trees"]:
        L = pt.Lca(np.array(t["parent"], dtype=np.int32))
        assert L.query(t["u"], t["v"]).tolist() == t["lca"]
        L.free()


@pytest.mark.parametrize("kind", ["random", "path", "star", "cat", "bushy"])
def test_shapes_and_sizes(pt, oracle, kind):
    rng = np.random.default_rng(11)
    sizes = [1, 2, 3, 31, 32, 33, 64, 65, 1000, 4097, 70000, 1 << 20]   # paths and caterpillars of 10^6 nodes included: the build
    for n in sizes:                                                       # ranks an Euler tour, its cost does not depend on the depth
        par = _shape(rng, n, kind)
        L = pt.Lca(par)
        q = 20000
        u = rng.integers(0, n, q).astype(np.int32)
        v = rng.integers(0, n, q).astype(np.int32)
        # half of the queries are "near" pairs so that the in-block path is exercised
        v[::2] = np.clip(u[::2] + rng.integers(-3, 4, len(u[::2])), 0, n - 1)
        got = L.query(u, v)
        exp = oracle.lca_euler(par, u, v)
        assert (got == exp).all(), (kind, n, int((got != exp).sum()))
        info = L.info()
        assert info["num_nodes"] == n
        d, p = L.node_info()
        assert sorted(p.tolist()) == list(range(n))  # preorder positions are a permutation
        root = int(np.where(par < 0)[0][0])
        assert d[root] == 0 and p[root] == 0
        nz = par >= 0
        assert (d[nz] == d[par[nz]] + 1).all() and (p[nz] > p[par[nz]]).all()
        L.free()


def test_same_block_neighbours(pt, oracle):
    """all pairs inside a few blocks of a small tree: every (l, r) offset combination of the in-block path"""
    rng = np.random.default_rng(5)
    par = _shape(rng, 96, "random")
    L = pt.Lca(par)
    u, v = np.meshgrid(np.arange(96, dtype=np.int32), np.arange(96, dtype=np.int32))
    u, v = u.ravel(), v.ravel()
    assert (L.query(u, v) == oracle.lca_naive(par, u, v)).all()
    L.free()


def test_adjacent_pairs_of_a_list(pt, oracle):
    """pt_lca_query_adjacent: out[i] = lca(nodes[i], nodes[i+1]) -- the reference's induced-subtree step on consecutive
    candidates (utils/induced_tree.hpp:128-138): preorder-sorted lists, random lists, lists with repeats and out-of-range ids,
    every length around the warp / tile boundaries, HOST and DEVICE memory"""
    import torch
    rng = np.random.default_rng(23)
    for kind, n in (("random", 5000), ("bushy", 70000), ("cat", 3000), ("star", 400)):
        par = _shape(rng, n, kind)
        L = pt.Lca(par)
        _, pre = L.node_info()
        by_pre = np.argsort(pre).astype(np.int32)
        for count in (0, 1, 2, 3, 31, 32, 33, 64, 1023, 1024, 1025, 2049, min(n, 50000)):
            for mode in ("sorted", "random", "repeats"):
                if mode == "sorted":
                    nodes = by_pre[np.sort(rng.choice(n, min(count, n), replace=False))] if count else np.zeros(0, np.int32)
                elif mode == "random":
                    nodes = rng.integers(0, n, count).astype(np.int32)
                else:
                    nodes = np.repeat(rng.integers(0, n, count // 2 + 1), 2)[:count].astype(np.int32)
                got = L.query_adjacent(nodes)
                assert len(got) == max(len(nodes) - 1, 0)
                if len(nodes) >= 2:
                    assert (got == oracle.lca_euler(par, nodes[:-1], nodes[1:])).all(), (kind, count, mode)
        # DEVICE lists, with ids outside the tree answering -1
        nodes = rng.integers(0, n, 4000).astype(np.int32)
        nodes[[7, 100, 3999]] = [-1, n, n + 5]
        d = torch.from_numpy(nodes).cuda()
        out = torch.empty(len(nodes) - 1, dtype=torch.int32, device="cuda")
        L.query_adjacent(d, out=out)
        torch.cuda.synchronize()
        got = out.cpu().numpy()
        ok = (nodes[:-1] >= 0) & (nodes[:-1] < n) & (nodes[1:] >= 0) & (nodes[1:] < n)
        assert (got[~ok] == -1).all()
        assert (got[ok] == oracle.lca_euler(par, nodes[:-1][ok], nodes[1:][ok])).all()
        L.free()


def test_device_memory_path(pt, oracle):
    import torch
    par = pt.random_tree(50000, 3)
    dpar = torch.from_numpy(par).cuda()
    L = pt.Lca(dpar)
    rng = np.random.default_rng(1)
    u = rng.integers(0, len(par), 100000).astype(np.int32)
    v = rng.integers(0, len(par), 100000).astype(np.int32)
    du, dv = torch.from_numpy(u).cuda(), torch.from_numpy(v).cuda()
    out = torch.empty_like(du)
    L.query(du, dv, out=out, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert (out.cpu().numpy() == oracle.lca_euler(par, u, v)).all()
    # out-of-range node ids answer -1 instead of reading out of bounds
    bad = torch.tensor([-1, len(par), 5], dtype=torch.int32).cuda()
    ok = torch.tensor([0, 0, 5], dtype=torch.int32).cuda()
    o = torch.empty_like(bad)
    L.query(bad, ok, out=o)
    torch.cuda.synchronize()
    assert o.cpu().tolist() == [-1, -1, 5]
    L.free()


def test_invalid_trees(pt):
    with pytest.raises(pt.PtError):
        pt.Lca(np.array([-1, -1, 0], dtype=np.int32))        # two roots (storage_adj_common.hpp:103-109)
    with pytest.raises(pt.PtError):
        pt.Lca(np.array([-1, 2, 1], dtype=np.int32))         # cycle next to the root
    with pytest.raises(pt.PtError):
        pt.Lca(np.array([1, 0, 0], dtype=np.int32))          # no root
    with pytest.raises(pt.PtError):
        pt.Lca(np.array([-1, 7], dtype=np.int32))            # parent out of range


def test_full_size_properties(pt, oracle):
    """BASELINE config 3 size: 10^7 leaves.  The oracle is too slow for all queries at this size, so: a sample against the
    parent-walk oracle plus size-independent properties (idempotence, symmetry, ancestor closure)."""
    import torch
    leaves = 10_000_000
    par = pt.random_tree(leaves, 0xB200)
    n = len(par)
    dpar = torch.from_numpy(par).cuda()
    L = pt.Lca(dpar)
    g = torch.Generator(device="cuda").manual_seed(7)
    q = 4_000_000
    u = torch.randint(0, n, (q,), generator=g, device="cuda", dtype=torch.int32)
    v = torch.randint(0, n, (q,), generator=g, device="cuda", dtype=torch.int32)
    a = torch.empty_like(u)
    L.query(u, v, out=a)
    b = torch.empty_like(u)
    L.query(v, u, out=b)
    torch.cuda.synchronize()
    assert torch.equal(a, b)                                   # symmetry
    c = torch.empty_like(u)
    L.query(a, u, out=c)
    torch.cuda.synchronize()
    assert torch.equal(c, a)                                   # lca(lca(u,v), u) == lca(u,v)
    L.query(u, u, out=c)
    torch.cuda.synchronize()
    assert torch.equal(c, u)                                   # lca(u,u) == u
    pu = dpar[u.long()]
    has = pu >= 0
    L.query(u[has], pu[has], out=c[: int(has.sum())])
    torch.cuda.synchronize()
    assert torch.equal(c[: int(has.sum())], pu[has])           # lca(u, parent(u)) == parent(u)
    k = 200_000
    uu, vv = u[:k].cpu().numpy(), v[:k].cpu().numpy()
    assert (a[:k].cpu().numpy() == oracle.lca_naive(par, uu, vv)).all()
    assert L.info()["num_levels"] < 200
    L.free()


def test_rmq_reference_answers(pt, oracle, gold_lca):
    for g in gold_lca["rmq"]:
        R = pt.Rmq(g["values"])
        got = R.query(g["l"], g["r"])
        vals = np.asarray(g["values"])
        assert (vals[got] == np.asarray(g["value"])).all()
        if g["mode"] == "rmq":
            assert (got == np.asarray(g["index"])).all()     # rightmost minimum, sparse_rmq.hpp:63,89
        R.free()


def test_rmq_known_answers_and_random(pt, oracle):
    from test_oracle import PM_RMQ_TEST_HPP, RMQ_TEST_HPP
    for arr, checks in RMQ_TEST_HPP + PM_RMQ_TEST_HPP:
        R = pt.Rmq(arr)
        got = R.query([c[0] for c in checks], [c[1] for c in checks])
        for (u, v, exp), idx in zip(checks, got):
            assert arr[idx] == arr[exp]
        R.free()
    rng = np.random.default_rng(9)
    for n in (1, 2, 33, 1000, 200000):
        a = rng.integers(-50, 50, n).astype(np.int32)
        R = pt.Rmq(a)
        l = rng.integers(0, n, 5000)
        r = np.minimum(n, l + 1 + rng.integers(0, max(1, n), 5000))
        got = R.query(l, r)
        assert (got == oracle.sparse_rmq(a, l, r)).all()
        R.free()


def test_who_displays_matches_reference_and_oracle(pt, oracle):
    """pt_tree_who_displays (wavefront over guest levels + device LCA) against the reference's who_displays lists
    (tests/golden/who_ref.json) and, where the reference crashed, against the oracle"""
    from conftest import load_golden
    g = load_golden("who_ref.json")
    for it in g["items"]:
        arrays = oracle.tree_arrays_from_newick(it["host"], it["guest"])
        got = pt.who_displays(*arrays).tolist()
        if isinstance(it["reference"], list):
            assert got == [(x[0] if x else -1) for x in it["reference"]], (it["host"], it["guest"])
        assert got == oracle.who_displays(*arrays).tolist(), (it["host"], it["guest"])


def test_who_displays_larger_trees(pt, oracle):
    rng = np.random.default_rng(8)
    for leaves, kind in [(40, 0), (150, 0), (150, 1), (400, 2)]:
        par = pt.random_tree(leaves, 100 + leaves)
        n = len(par)
        hl = np.full(n, -1, dtype=np.int32)
        hl[:leaves] = rng.permutation(leaves)
        if kind == 0:        # guest = the same tree with node ids permuted: every subtree is displayed by its own image
            perm = rng.permutation(n).astype(np.int32)
            gp = np.full(n, -1, dtype=np.int32)
            gp[perm] = np.where(par >= 0, perm[np.maximum(par, 0)], -1)
            gl = np.full(n, -1, dtype=np.int32)
            gl[perm] = hl
        else:                # another random tree on (a subset of) the labels
            gleaves = leaves if kind == 1 else leaves // 2
            gp = pt.random_tree(gleaves, 7 + leaves)
            gl = np.full(len(gp), -1, dtype=np.int32)
            gl[:gleaves] = rng.permutation(leaves)[:gleaves]
        got = pt.who_displays(par, hl, gp, gl)
        exp = oracle.who_displays(par, hl, gp, gl)
        assert (got == exp).all(), (leaves, kind, np.where(got != exp)[0][:5])
        if kind == 0:
            assert (got >= 0).all()
    with pytest.raises(pt.PtError):
        pt.who_displays([-1, 0, 0], [-1, 5, 5], [-1, 0, 0], [-1, 5, 6])   # multi-labelled host



def test_highest_displayed_ancestor(pt, oracle):
    """pt_tree_highest_displayed (pointer jumping over the who_displays table) against the restated loop of
    TreeInComponent::highest_displayed_ancestor (containment.hpp:328-344): golden pairs, a hand-checked case, deep and random tables"""
    from conftest import load_golden
    hp, hl, gp, gl = oracle.tree_arrays_from_newick("((a,b),(c,d));", "((a,c),(b,d));")
    who = pt.who_displays(hp, hl, gp, gl)
    assert who.tolist() == [-1, 0, 2, 5, 0, 3, 6]
    assert pt.highest_displayed(gp, who, 0).tolist() == [0, 1, 1, 1, 4, 4, 4]
    for it in load_golden("who_ref.json")["items"][:60]:
        hp, hl, gp, gl = oracle.tree_arrays_from_newick(it["host"], it["guest"])
        who = pt.who_displays(hp, hl, gp, gl)
        root = int(np.where(np.asarray(hp) < 0)[0][0])
        assert pt.highest_displayed(gp, who, root).tolist() == oracle.highest_displayed(gp, who, root).tolist(), (it["host"], it["guest"])
    rng = np.random.default_rng(21)
    for n, shape in [(5000, "path"), (20000, "random"), (300000, "random"), (100000, "path")]:
        if shape == "path":      # node i hangs below i - 1: the climb is as long as it can be
            gp = np.arange(-1, n - 1, dtype=np.int32)
        else:
            gp = np.concatenate([[-1], (rng.random(n - 1) * np.arange(1, n)).astype(np.int64)]).astype(np.int32)
        for holes in (0.0, 0.001, 0.2):
            who = rng.integers(0, 50, size=n).astype(np.int32)          # 0 plays the host root
            who[rng.random(n) < holes] = -1
            if holes == 0.0:
                who[:] = 7
            got = pt.highest_displayed(gp, who, 0)
            assert (got == oracle.highest_displayed(gp, who, 0)).all(), (n, shape, holes)
    with pytest.raises(pt.PtError):
        pt.highest_displayed(np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.int32), 0)
    # arrays already on the device (PT_MEM_DEVICE): nothing is copied, the answer stays there
    import torch
    n = 50000
    gp = np.concatenate([[-1], (rng.random(n - 1) * np.arange(1, n)).astype(np.int64)]).astype(np.int32)
    who = rng.integers(0, 9, size=n).astype(np.int32); who[rng.random(n) < 0.05] = -1
    d_gp, d_who, d_out = torch.from_numpy(gp).cuda(), torch.from_numpy(who).cuda(), torch.empty(n, dtype=torch.int32, device="cuda")
    assert pt.lib().pt_tree_highest_displayed(d_gp.data_ptr(), d_who.data_ptr(), n, 0, d_out.data_ptr(), 1, None) == 0
    assert (d_out.cpu().numpy() == oracle.highest_displayed(gp, who, 0)).all()


def _csr(n, edges):
    edges = sorted(edges)
    off = np.zeros(n + 1, dtype=np.int32)
    for a, _ in edges:
        off[a + 1] += 1
    return np.cumsum(off).astype(np.int32), np.array([b for _, b in edges], dtype=np.int32).reshape(-1), edges


def test_bridges_match_reference_and_oracle(pt, oracle):
    """pt_net_bridges (spanning tree + batched LCA + difference counting) against the reference BridgeFinder's bridge
    sets (tests/golden/bridges_ref.json) and the remove-and-test oracle; components = union of non-bridge arcs"""
    from conftest import load_golden
    from oracle.newick import parse_newick
    g = load_golden("bridges_ref.json")
    for it in g["items"]:
        n, e, _ = parse_newick(it["newick"])
        off, adj, es = _csr(n, e)
        br, comp = pt.net_bridges(off, adj)
        assert sorted([list(x) for x, b in zip(es, br) if b]) == it["reference"], it["newick"][:60]
        # components: two nodes share an id iff they are joined by non-bridge arcs
        uf = list(range(n))
        def find(x):
            while uf[x] != x:
                uf[x] = uf[uf[x]]; x = uf[x]
            return x
        for (a, b), isb in zip(es, br):
            if not isb:
                uf[find(a)] = find(b)
        groups = {}
        for v in range(n):
            groups.setdefault(find(v), set()).add(int(comp[v]))
        assert all(len(s) == 1 for s in groups.values()) and len({next(iter(s)) for s in groups.values()}) == len(groups)
    # a larger network against the oracle
    p = pt.Packer()
    p.add_generated(3, 400, 60, 5, 0)
    a = p.arrays()
    for i in range(3):
        hb, he = int(a["host_node_off"][i]), int(a["host_node_off"][i + 1])
        n = he - hb
        off = a["host_succ_off"][hb + i: hb + i + n + 1]
        adj = a["host_succ_adj"][int(a["host_arc_off"][i]): int(a["host_arc_off"][i + 1])]
        br, comp = pt.net_bridges(off, adj)
        edges = [(int(t), int(h)) for t, h in zip(np.repeat(np.arange(n), np.diff(off)), adj)]
        assert (br == oracle.bridges(n, edges)).all()
