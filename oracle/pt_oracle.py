"""oracle/pt_oracle.py -- TEST INFRASTRUCTURE ONLY: ctypes access to oracle/libptoracle.so (pt_oracle.c)
and to the reference binaries in oracle/_ref/.  Imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never by phylo_tools_b200/."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(HERE, "libptoracle.so")
    src = os.path.join(HERE, "pt_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", src, "-o", so])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        L = _LIB
        i32p, u64p = C.POINTER(C.c_int32), C.POINTER(C.c_uint64)
        L.orc_tc_bruteforce.restype = C.c_int
        L.orc_tc_bruteforce.argtypes = [C.c_int, C.c_int, i32p, i32p, i32p, C.c_int, C.c_int, i32p, i32p, i32p, C.c_int, C.c_uint64, u64p, u64p]
        L.orc_lca_naive_batch.restype = None
        L.orc_lca_naive_batch.argtypes = [i32p, C.c_int32, i32p, i32p, i32p, C.c_int64]
        L.orc_lca_euler_build.restype = C.c_void_p
        L.orc_lca_euler_build.argtypes = [i32p, C.c_int32]
        L.orc_lca_euler_batch.restype = None
        L.orc_lca_euler_batch.argtypes = [C.c_void_p, i32p, i32p, i32p, C.c_int64]
        L.orc_lca_euler_free.argtypes = [C.c_void_p]
        L.orc_sparse_rmq_build.restype = C.c_void_p
        L.orc_sparse_rmq_build.argtypes = [i32p, C.c_int64]
        L.orc_sparse_rmq_query.restype = C.c_int64
        L.orc_sparse_rmq_query.argtypes = [C.c_void_p, C.c_int64, C.c_int64]
        L.orc_sparse_rmq_free.argtypes = [C.c_void_p]
        L.orc_bridges.restype = C.c_int
        L.orc_bridges.argtypes = [C.c_int, C.c_int, i32p, i32p, C.POINTER(C.c_uint8)]
        L.orc_who_displays.restype = C.c_int
        L.orc_who_displays.argtypes = [C.c_int, i32p, i32p, C.c_int, i32p, i32p, i32p]
        L.orc_highest_displayed.restype = C.c_int
        L.orc_highest_displayed.argtypes = [C.c_int, i32p, i32p, C.c_int32, i32p]
        L.orc_mul_who_displays.restype = C.c_int
        L.orc_mul_who_displays.argtypes = [C.c_int, i32p, i32p, C.c_int, i32p, i32p, i32p, i32p, C.c_int]
        L.orc_highest_displayed_lists.restype = C.c_int
        L.orc_highest_displayed_lists.argtypes = [C.c_int, i32p, i32p, i32p, C.c_int32, i32p]
        L.orc_unzip.restype = C.c_int
        L.orc_unzip.argtypes = [C.c_int, i32p, i32p, i32p, C.c_int, i32p, i32p, i32p, C.c_int]
        L.orc_iso.restype = C.c_int
        L.orc_iso.argtypes = [C.c_int, C.c_int, i32p, i32p, i32p, C.c_int, C.c_int, i32p, i32p, i32p, C.c_int]
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _arr(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.int32))


def tc_bruteforce(h_edges, h_n, h_labels, t_edges, t_n, t_labels, stop_at_first=True, max_switchings=1 << 26):
    """h_edges/t_edges: lists of (tail, head); *_labels: int array (label id or -1) per node.
    returns (verdict, n_displaying, first_index); verdict 1/0 or negative error."""
    he = _arr(h_edges).reshape(-1, 2)
    te = _arr(t_edges).reshape(-1, 2)
    ht, hh = _arr(he[:, 0]), _arr(he[:, 1])
    tt, th = _arr(te[:, 0]), _arr(te[:, 1])
    hl, tl = _arr(h_labels), _arr(t_labels)
    nd, fi = C.c_uint64(0), C.c_uint64(0)
    rc = lib().orc_tc_bruteforce(int(h_n), len(ht), _p(ht), _p(hh), _p(hl), int(t_n), len(tt), _p(tt), _p(th), _p(tl),
                                 1 if stop_at_first else 0, int(max_switchings), C.byref(nd), C.byref(fi))
    return rc, nd.value, fi.value


def tc_bruteforce_newick(host_nw, guest_nw, **kw):
    """tc semantics of examples/tc.cpp:110-114 on two Newick strings, through oracle/newick.py"""
    from .newick import parse_newick
    n0, e0, l0 = parse_newick(host_nw)
    n1, e1, l1 = parse_newick(guest_nw)
    if len(e0) == n0 - 1 and len(e1) != n1 - 1:
        n0, e0, l0, n1, e1, l1 = n1, e1, l1, n0, e0, l0
    names = {}
    def enc(n, lab):
        out = np.full(n, -1, dtype=np.int32)
        for v, s in lab.items():
            out[v] = names.setdefault(s, len(names))
        return out
    return tc_bruteforce(e0, n0, enc(n0, l0), e1, n1, enc(n1, l1), **kw)


def lca_naive(parent, u, v):
    parent, u, v = _arr(parent), _arr(u), _arr(v)
    out = np.empty(len(u), dtype=np.int32)
    lib().orc_lca_naive_batch(_p(parent), len(parent), _p(u), _p(v), _p(out), len(u))
    return out


def lca_euler(parent, u, v):
    parent, u, v = _arr(parent), _arr(u), _arr(v)
    h = lib().orc_lca_euler_build(_p(parent), len(parent))
    out = np.empty(len(u), dtype=np.int32)
    lib().orc_lca_euler_batch(h, _p(u), _p(v), _p(out), len(u))
    lib().orc_lca_euler_free(h)
    return out


def sparse_rmq(values, ls, rs):
    values = _arr(values)
    h = lib().orc_sparse_rmq_build(_p(values), len(values))
    out = np.array([lib().orc_sparse_rmq_query(h, int(l), int(r)) for l, r in zip(ls, rs)], dtype=np.int64)
    lib().orc_sparse_rmq_free(h)
    return out


def bridges(n, edges):
    """bool per arc: is the arc a bridge of the underlying undirected graph (utils/bridges.hpp)"""
    e = _arr(edges).reshape(-1, 2)
    t, h = _arr(e[:, 0]), _arr(e[:, 1])
    out = np.zeros(max(len(t), 1), dtype=np.uint8)
    lib().orc_bridges(int(n), len(t), _p(t), _p(h), out.ctypes.data_as(C.POINTER(C.c_uint8)))
    return out[:len(t)].astype(bool)


def iso(n1, edges1, labels1, n2, edges2, labels2, flags=1):
    """network isomorphism by definition (backtracking); labels*: dense ids shared by both networks, -1 = none;
    flags: 1 leaf labels count, 2 internal tree nodes, 4 internal reticulations (utils/isomorphism.hpp:10-13)"""
    e1, e2 = _arr(edges1).reshape(-1, 2), _arr(edges2).reshape(-1, 2)
    t1, h1, t2, h2 = _arr(e1[:, 0]), _arr(e1[:, 1]), _arr(e2[:, 0]), _arr(e2[:, 1])
    l1, l2 = _arr(labels1), _arr(labels2)
    return int(lib().orc_iso(int(n1), len(t1), _p(t1), _p(h1), _p(l1), int(n2), len(t2), _p(t2), _p(h2), _p(l2), int(flags)))


def iso_newick(nw1, nw2, flags=1):
    """the same for two extended-Newick strings (label strings -> shared dense ids)"""
    from .newick import parse_newick
    n1, e1, lab1 = parse_newick(nw1)
    n2, e2, lab2 = parse_newick(nw2)
    ids = {}
    def enc(n, lab):
        out = np.full(n, -1, dtype=np.int32)
        for v, name in lab.items():
            if name != "":
                out[int(v)] = ids.setdefault(name, len(ids))
        return out
    return iso(n1, e1, enc(n1, lab1), n2, e2, enc(n2, lab2), flags)


def who_displays(host_parent, host_label, guest_parent, guest_label):
    hp, hl, gp, gl = _arr(host_parent), _arr(host_label), _arr(guest_parent), _arr(guest_label)
    out = np.empty(len(gp), dtype=np.int32)
    rc = lib().orc_who_displays(len(hp), _p(hp), _p(hl), len(gp), _p(gp), _p(gl), _p(out))
    if rc != 0:
        raise ValueError("orc_who_displays failed: %d" % rc)
    return out


def mul_who_displays(host_parent, host_label, guest_parent, guest_label):
    """orc_mul_who_displays: (off[nt+1], adj) = per guest node the host nodes that minimally display it, ascending node id;
    the host may be multi-labelled (utils/containment.hpp:72-230)"""
    hp, hl, gp, gl = _arr(host_parent), _arr(host_label), _arr(guest_parent), _arr(guest_label)
    cap = max(16, 4 * (len(hp) + len(gp)) * 4)
    while True:
        off = np.zeros(len(gp) + 1, dtype=np.int32); adj = np.zeros(cap, dtype=np.int32)
        rc = lib().orc_mul_who_displays(len(hp), _p(hp), _p(hl), len(gp), _p(gp), _p(gl), _p(off), _p(adj), cap)
        if rc == -1:
            cap *= 4
            continue
        if rc < 0:
            raise ValueError("orc_mul_who_displays failed: %d" % rc)
        return off, adj[:rc].copy()


def highest_displayed_lists(guest_parent, off, adj, host_root):
    gp, o, a = _arr(guest_parent), _arr(off), _arr(adj if len(adj) else [0])
    out = np.empty(len(gp), dtype=np.int32)
    lib().orc_highest_displayed_lists(len(gp), _p(gp), _p(o), _p(a), int(host_root), _p(out))
    return out


def unzip(succ_off, succ_adj, label, u):
    """orc_unzip (TreeInComponent::unzip_retis, utils/containment.hpp:262-312) -> (parent, label, origin) of the MUL-tree below u"""
    so, sa, lab = _arr(succ_off), _arr(succ_adj if len(succ_adj) else [0]), _arr(label)
    cap = 1 << 12
    while True:
        par, ol, org = (np.empty(cap, dtype=np.int32) for _ in range(3))
        k = lib().orc_unzip(len(so) - 1, _p(so), _p(sa), _p(lab), int(u), _p(par), _p(ol), _p(org), cap)
        if k >= 0:
            return par[:k].copy(), ol[:k].copy(), org[:k].copy()
        cap *= 8
        if cap > 1 << 27:
            raise ValueError("orc_unzip: the MUL-tree is too large")


def highest_displayed(guest_parent, who, host_root):
    """orc_highest_displayed (utils/containment.hpp:328-344) for every guest node"""
    gp = np.ascontiguousarray(guest_parent, dtype=np.int32); wd = np.ascontiguousarray(who, dtype=np.int32)
    out = np.empty(len(gp), dtype=np.int32)
    lib().orc_highest_displayed(len(gp), _p(gp), _p(wd), int(host_root), _p(out))
    return out


def component_dag(succ_off, succ_adj):
    """TreeComponentInfos for the component roots (utils/tree_components.hpp:68-131), restated: comp_root of every internal tree node,
    and per non-trivial root v the multiset of component roots reached by climbing through the reticulations above v -- one entry per
    PATH, which is what VisibleComponentRule::is_half_eligible counts (utils/containment.hpp:664-676).
    -> (comp_root dict, arcs dict {(root above, v): number of paths}, network root)"""
    so, sa = list(map(int, succ_off)), list(map(int, succ_adj))
    n = len(so) - 1
    parents = [[] for _ in range(n)]
    for u in range(n):
        for k in range(so[u], so[u + 1]):
            parents[sa[k]].append(u)
    root = [v for v in range(n) if not parents[v]][0]
    comp = {root: root}

    def comp_of(u):          # internal tree nodes only (:78-86)
        chain = []
        while u not in comp:
            chain.append(u)
            p = parents[u][0]
            if len(parents[p]) > 1:
                comp[u] = u
                break
            u = p
        for x in reversed(chain):
            if x not in comp:
                comp[x] = comp[parents[x][0]]
        return comp[chain[0]] if chain else comp[u]
    arcs = {}
    for v in range(n):
        if len(parents[v]) == 1 and so[v + 1] > so[v] and len(parents[parents[v][0]]) > 1:   # non-trivial root
            comp[v] = v
            stack = [parents[v][0]]
            while stack:
                r = stack.pop()
                for p in parents[r]:
                    if len(parents[p]) > 1:
                        stack.append(p)
                    else:
                        c = comp_of(p)
                        arcs[(c, v)] = arcs.get((c, v), 0) + 1
    return comp, arcs, root


def vcr_half_eligible(succ_off, succ_adj, order=None):
    """VisibleComponentRule::is_half_eligible (utils/containment.hpp:652-687) for the roots of the component DAG.
    order = None: every path count taken once per component-DAG arc (the semantics the library implements) -> {root: bool}.
    order = [[u, child, child, ...], ...] (the reference's own visiting order, printed by `ref_tc vcr`): the loop exactly as written,
    including its shared counters -- walking the reticulations above a root v adds to num_paths[c][v] for EVERY component c above v, and
    the walk is repeated for every component above v, so later components see doubled counts."""
    comp, arcs, root = component_dag(succ_off, succ_adj)
    kids = {}
    for (c, v), k in arcs.items():
        kids.setdefault(c, []).append(v)
    if order is None:
        half, paths = {}, {}

        def visit(u):
            if u in half:
                return
            for v in kids.get(u, []):
                visit(v)
            ok, acc = True, {}
            for v in kids.get(u, []):
                if not half[v] or arcs[(u, v)] > 1:
                    ok = False
                    continue
                acc[v] = acc.get(v, 0) + 1
                for x, k in paths[v].items():
                    acc[x] = acc.get(x, 0) + k
            ok = ok and all(k <= 1 for k in acc.values())
            half[u], paths[u] = ok, (acc if ok else {})
        for u in [root] + sorted(kids) + sorted(v for vs in kids.values() for v in vs):
            visit(u)
        return half
    num_paths, half = {}, {}
    above = {}
    for (c, v), k in arcs.items():
        above.setdefault(v, []).append((c, k))
    for row in order:
        u, children = row[0], row[1:]
        up = num_paths.setdefault(u, {})
        ok = True
        for v in children:
            if not half.get(v, False):
                ok = False
                break
            for c, k in above.get(v, []):           # the walk above v registers a path for every tree node it ends in (:664-674)
                d = num_paths.setdefault(c, {})
                d[v] = d.get(v, 0) + k
            if up.get(v, 0) > 1:
                ok = False
                break
            bad = False
            for x, k in list(num_paths.setdefault(v, {}).items()):
                up[x] = up.get(x, 0) + k
                if up[x] > 1:
                    bad = True
                    break
            if bad:
                ok = False
                break
        half[u] = ok
    return half


def tree_arrays_from_newick(host_nw, guest_nw):
    """two TREE Newick strings -> (host_parent, host_label, guest_parent, guest_label) with a shared dense label dictionary,
    node ids as the reference reader assigns them (oracle/newick.py)"""
    from .newick import parse_newick
    names = {}
    res = []
    for s in (host_nw, guest_nw):
        n, e, lab = parse_newick(s)
        par = np.full(n, -1, dtype=np.int32)
        for a, b in e:
            par[b] = a
        l = np.full(n, -1, dtype=np.int32)
        for v, name in lab.items():
            l[v] = names.setdefault(name, len(names))
        res += [par, l]
    return tuple(res)


def ref_binary(name):
    """path of a reference binary built by oracle/Makefile (None if absent: /root/reference is not on the GPU box,
    but the prebuilt oracle/_ref/ travels with the snapshot)"""
    p = os.path.join(HERE, "_ref", name)
    return p if os.path.exists(p) else None
