"""profiles/who_timing.py -- wall time (upload and download included) of pt_tree_who_displays and of the general DP (pt_who_build + lists) on
host = guest = random binary tree of 10^6 leaves, and of the general DP on a 3x multi-labelled host (6*10^5 leaves, 2*10^5 labels)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import phylo_tools_b200 as pt

pt.set_device(0)
SEED = 0xB200
par = pt.random_tree(1_000_000, SEED + 7); n = len(par)
deg = np.bincount(par[par >= 0], minlength=n)
lab = np.full(n, -1, dtype=np.int32); lab[deg == 0] = np.arange(int((deg == 0).sum()), dtype=np.int32)
def best(f, reps=4):
    f(); ts = []
    for _ in range(reps):
        t = time.perf_counter(); f(); ts.append(time.perf_counter() - t)
    return min(ts) * 1e3
print("pt_tree_who_displays, 2e6 guest nodes: %.2f ms" % best(lambda: pt.who_displays(par, lab, par, lab)))
def gen():
    W = pt.Who(par, lab, par, lab); W.free()
print("pt_who_build, same pair: %.2f ms" % best(gen))
rng = np.random.default_rng(3)
leaves = 200_000
hp = pt.random_tree(3 * leaves, SEED + 8); hn = len(hp)
hdeg = np.bincount(hp[hp >= 0], minlength=hn); hl = np.full(hn, -1, dtype=np.int32); hl[hdeg == 0] = rng.permutation(np.repeat(np.arange(leaves), 3)).astype(np.int32)
gp = pt.random_tree(leaves, SEED + 9); gn = len(gp)
gdeg = np.bincount(gp[gp >= 0], minlength=gn); gl = np.full(gn, -1, dtype=np.int32); gl[gdeg == 0] = np.arange(leaves, dtype=np.int32)
def mul():
    W = pt.Who(hp, hl, gp, gl); W.free()
print("pt_who_build, multi-labelled host (3 leaves per label, 6e5 leaves), guest 2e5 leaves: %.2f ms" % best(mul))
