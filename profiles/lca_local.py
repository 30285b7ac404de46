"""profiles/lca_local.py -- the LCA workload the containment path really issues (utils/induced_tree.hpp:128-138): the candidates of a
guest node are sorted by host preorder and the k-1 queries are CONSECUTIVE pairs of that list.  Host tree numbered the way the
reference numbers it (node ids = a DFS preorder, io/newick.hpp:164 / containment.hpp:292), candidates = a random subset of the
nodes at density p, sorted by preorder.  Usage: python profiles/lca_local.py [leaves]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import phylo_tools_b200 as pt

leaves = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
pt.set_device(0)
par0 = torch.from_numpy(pt.random_tree(leaves, 0xB200)).cuda()
n = par0.numel()
l0 = pt.Lca(par0)
pre = torch.empty(n, dtype=torch.int32, device="cuda")
pt._check(pt.lib().pt_lca_node_info(l0._h, None, pre.data_ptr(), pt.PT_MEM_DEVICE, None), "node_info")
l0.free()
# renumber: new id = preorder position
newpar = torch.empty_like(par0)
pp = torch.where(par0 >= 0, pre[par0.clamp(min=0).long()], torch.full_like(par0, -1))
newpar[pre.long()] = pp
del par0, pp
lca = pt.Lca(newpar)
pos = torch.empty(n, dtype=torch.int32, device="cuda")
pt._check(pt.lib().pt_lca_node_info(lca._h, None, pos.data_ptr(), pt.PT_MEM_DEVICE, None), "node_info")
stream = torch.cuda.current_stream().cuda_stream
g = torch.Generator(device="cuda").manual_seed(7)
for dens in (1.0, 0.5, 0.25, 0.1):
    keep = torch.rand(n, generator=g, device="cuda") < dens
    cand = torch.nonzero(keep).squeeze(1).to(torch.int32)
    order = torch.argsort(pos[cand.long()])
    cand = cand[order].contiguous()
    u, v = cand[:-1].contiguous(), cand[1:].contiguous()
    q = u.numel()
    out = torch.empty_like(u)
    for _ in range(3):
        lca.query(u, v, out=out, stream=stream)
    torch.cuda.synchronize()
    ms = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); lca.query(u, v, out=out, stream=stream); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    t = float(np.mean(ms))
    ms2 = []
    out2 = torch.empty_like(u)
    for it in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); lca.query_adjacent(cand, out=out2, stream=stream); e1.record(); torch.cuda.synchronize()
        if it >= 3:
            ms2.append(e0.elapsed_time(e1))
    assert bool((out2 == out).all())
    t2 = float(np.mean(ms2))
    print(f"density {dens}: adjacent-list call {t2:.3f} ms = {q / t2 / 1e6:.1f} G queries/s = {40 * q / t2 / 1e6 / 6536.7:.3f} of the measured peak on the 40 B model")
    print(f"density {dens}: {q} queries in {t:.3f} ms = {q / t / 1e6:.1f} G queries/s = {40 * q / t / 1e6:.0f} GB/s on the 40 B model = {40 * q / t / 1e6 / 6536.7:.3f} of the measured peak", flush=True)
# uniform pairs for comparison
u = torch.randint(0, n, (100_000_000,), generator=g, device="cuda", dtype=torch.int32)
v = torch.randint(0, n, (100_000_000,), generator=g, device="cuda", dtype=torch.int32)
out = torch.empty_like(u)
for _ in range(2):
    lca.query(u, v, out=out, stream=stream)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); lca.query(u, v, out=out, stream=stream); e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1)
print(f"uniform pairs: 1e8 queries in {t:.3f} ms = {1e8 / t / 1e6:.1f} G queries/s")
