import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import phylo_tools_b200 as pt
pt.set_device(0)
B = 20000
p = pt.Packer(); p.add_generated(B + 1, 100, 20, 0xB200, 0)
a = p.arrays()
n = int(a["host_node_off"][1]); m = int(a["host_arc_off"][1])
assert (np.diff(a["host_node_off"]) == n).all() and (np.diff(a["host_arc_off"]) == m).all()
so = a["host_succ_off"].reshape(B + 1, n + 1); sa = a["host_succ_adj"].reshape(B + 1, m)
po = a["host_pred_off"].reshape(B + 1, n + 1); pa = a["host_pred_adj"].reshape(B + 1, m); lab = a["host_label"].reshape(B + 1, n)
def arrays(idx):
    G = len(idx)
    return dict(num_pairs=G // 2, node_off=np.arange(G + 1, dtype=np.int64) * n, arc_off=np.arange(G + 1, dtype=np.int64) * m,
                succ_off=np.ascontiguousarray(so[idx]).ravel(), succ_adj=np.ascontiguousarray(sa[idx]).ravel(), pred_off=np.ascontiguousarray(po[idx]).ravel(),
                pred_adj=np.ascontiguousarray(pa[idx]).ravel(), label=np.ascontiguousarray(lab[idx]).ravel(), max_nodes=n)
same = arrays(np.repeat(np.arange(B), 2))
diff = arrays(np.stack([np.arange(B), np.arange(B) + 1], axis=1).ravel())
def pin(arr):
    return {k: (torch.from_numpy(v).pin_memory().numpy() if isinstance(v, np.ndarray) else v) for k, v in arr.items()}
for name, arr, want in (("identical copies", same, B), ("different networks", diff, 0), ("identical copies, pinned arrays", pin(same), B), ("different networks, pinned arrays", pin(diff), 0)):
    pt.iso_batch(arr)
    t0 = time.perf_counter(); v, s = pt.iso_batch(arr); dt = time.perf_counter() - t0
    print(f"{name}: {B / dt / 1e6:.3f} M pairs/s ({dt * 1e3:.1f} ms incl. upload), isomorphic {int(v.sum())} (expected {want}), status ok {bool((s == 0).all())}")
# symmetric worst case: complete binary trees, no label counts
for k in (3, 4, 5, 6):
    def tree(lo, hi): return "l%d" % lo if hi - lo == 1 else "(" + tree(lo, (lo + hi) // 2) + "," + tree((lo + hi) // 2, hi) + ")"
    s1 = tree(0, 1 << k) + ";"
    t0 = time.perf_counter(); v, s = pt.iso_newick([(s1, s1)], flags=0); dt = time.perf_counter() - t0
    print(f"complete binary tree, {1 << k} leaves, flags=0: isomorphic {bool(v[0])} status {int(s[0])} in {dt * 1e3:.1f} ms")
