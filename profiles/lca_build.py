"""profiles/lca_build.py -- pt_lca_build time (CUDA events inside the library) by tree shape: random binary tree of 10^7 leaves (BASELINE
config 3), path and caterpillar of 10^6 nodes, star; set PT_LCA_SWEEPS=1 for the level-synchronous build it replaced."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import phylo_tools_b200 as pt
pt.set_device(0)
def shape(kind, n):
    idx = np.arange(1, n)
    par = np.full(n, -1, dtype=np.int32)
    if kind == "path": par[1:] = idx - 1
    elif kind == "cat": par[1:] = np.where(idx % 2 == 1, idx - 1, np.maximum(idx - 2, 0))
    elif kind == "star": par[1:] = 0
    return par
cases = [("random 10^7 leaves", pt.random_tree(10_000_000, 0xB200))]
if not os.environ.get("PT_LCA_SWEEPS"):
    cases += [("path 10^6", shape("path", 1_000_000)), ("caterpillar 10^6", shape("cat", 1_000_000))]
else:
    cases += [("path 2*10^4", shape("path", 20_000))]
cases += [("star 10^6", shape("star", 1_000_000))]
for name, par in cases:
    d = torch.from_numpy(par).cuda()
    best = 1e9
    for _ in range(4):
        L = pt.Lca(d); i = L.info(); L.free(); best = min(best, i["build_ms"])
    print(f"{name:22s} n = {len(par):9d}  levels {i['num_levels']:8d}  build {best:8.3f} ms", flush=True)
