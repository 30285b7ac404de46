"""profiles/r02_driver.py -- one launch of every kernel family at its bench.py size, for `ncu --set full` captures (round 2):
    python profiles/r02_driver.py tc | tc_shard | lca | who | enum | iso | bridges | text | config4
Each mode warms the library's allocator up first (the captured launches are the steady-state ones: use ncu's -s / -k)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import phylo_tools_b200 as pt

what = sys.argv[1] if len(sys.argv) > 1 else "tc"
pt.set_device(0)
SEED = 0xB200
if what == "tc":
    B = 100_000
    p = pt.Packer(); p.add_generated(B, 100, 20, SEED, 0)
    dev = {k: (torch.from_numpy(v).cuda() if isinstance(v, np.ndarray) else v) for k, v in p.arrays(compact=1).items()}
    bits = torch.zeros((B + 31) // 32, dtype=torch.int32, device="cuda"); stat = torch.zeros(B, dtype=torch.uint8, device="cuda")
    b = pt.Batch(dev)
    for _ in range(4):
        b.contain(result_bits=bits, status=stat, want_counters=False)
    torch.cuda.synchronize(); print("tc", b.round_ms(0))
elif what == "tc_shard":      # one rank's shard of config 2 at 8 GPUs: the small-core variant of the persistent kernel
    B = 12_512
    p = pt.Packer(); p.add_generated(B, 100, 20, SEED, 0)
    dev = {k: (torch.from_numpy(v).cuda() if isinstance(v, np.ndarray) else v) for k, v in p.arrays(compact=1).items()}
    bits = torch.zeros((B + 31) // 32, dtype=torch.int32, device="cuda"); stat = torch.zeros(B, dtype=torch.uint8, device="cuda")
    b = pt.Batch(dev)
    for _ in range(4):
        b.contain(result_bits=bits, status=stat, want_counters=False)
    torch.cuda.synchronize(); print("tc_shard", b.round_ms(0))
elif what == "lca":
    par0 = torch.from_numpy(pt.random_tree(10_000_000, SEED)).cuda(); n = par0.numel()
    L = pt.Lca(par0)
    g = torch.Generator(device="cuda").manual_seed(1)
    u = torch.randint(0, n, (100_000_000,), generator=g, device="cuda", dtype=torch.int32); v = torch.randint(0, n, (100_000_000,), generator=g, device="cuda", dtype=torch.int32)
    out = torch.empty_like(u)
    for _ in range(2):
        L.query(u, v, out=out)
    pre = torch.empty(n, dtype=torch.int32, device="cuda")
    pt._check(pt.lib().pt_lca_node_info(L._h, None, pre.data_ptr(), pt.PT_MEM_DEVICE, None), "node_info")
    L.free(); del u, v, out
    newpar = torch.empty_like(par0)
    newpar[pre.long()] = torch.where(par0 >= 0, pre[par0.clamp(min=0).long()], torch.full_like(par0, -1))
    L = pt.Lca(newpar)
    pos = torch.empty(n, dtype=torch.int32, device="cuda")
    pt._check(pt.lib().pt_lca_node_info(L._h, None, pos.data_ptr(), pt.PT_MEM_DEVICE, None), "node_info")
    cand = torch.argsort(pos).to(torch.int32).contiguous()
    o2 = torch.empty(n - 1, dtype=torch.int32, device="cuda")
    for _ in range(2):
        L.query_adjacent(cand, out=o2)
    torch.cuda.synchronize(); print("lca ok")
elif what == "who":
    par = pt.random_tree(1_000_000, SEED + 7); n = len(par)
    deg = np.bincount(par[par >= 0], minlength=n)
    lab = np.full(n, -1, dtype=np.int32); lab[deg == 0] = np.arange(int((deg == 0).sum()), dtype=np.int32)
    for _ in range(2):
        W = pt.Who(par, lab, par, lab); W.free()
    # a multi-labelled host: every label three times
    rng = np.random.default_rng(3)
    leaves = 200_000
    hp = pt.random_tree(3 * leaves, SEED + 8); hn = len(hp)
    hdeg = np.bincount(hp[hp >= 0], minlength=hn); hl = np.full(hn, -1, dtype=np.int32); hl[hdeg == 0] = rng.permutation(np.repeat(np.arange(leaves), 3)).astype(np.int32)
    gp = pt.random_tree(leaves, SEED + 9); gn = len(gp)
    gdeg = np.bincount(gp[gp >= 0], minlength=gn); gl = np.full(gn, -1, dtype=np.int32); gl[gdeg == 0] = np.arange(leaves, dtype=np.int32)
    for _ in range(2):
        W = pt.Who(hp, hl, gp, gl); off, adj = W.lists(); W.free()
    print("who ok", int(off[-1]), "entries on the multi-labelled host")
elif what == "enum":
    p = pt.Packer(); p.add_generated(1, 30, 24, SEED + 5, 2); a = p.arrays()
    E = pt.Enum(a["host_succ_off"], a["host_succ_adj"], a["host_label"], a["guest_parent"], a["guest_label"])
    for _ in range(2):
        print(E.run(chunk=1 << 24))
    E.free()
elif what == "iso":
    pairs = 20000
    p = pt.Packer(); p.add_generated(pairs + 1, 100, 20, SEED + 6, 0)
    lines = [p.newick(i)[0] for i in range(pairs + 1)]
    text = ("\n".join(lines[i] + "\n" + lines[i + 1] for i in range(pairs)) + "\n")
    for _ in range(2):
        v, s = pt.iso_newick_text(text)
    print("iso ok", int(v.sum()))
elif what == "bridges":
    p = pt.Packer(); p.add_generated(1, 499_001, 1000, SEED + 8, 0); a = p.arrays()
    n, m = int(a["host_node_off"][1]), int(a["host_arc_off"][1])
    for _ in range(2):
        br, comp = pt.net_bridges(a["host_succ_off"][:n + 1], a["host_succ_adj"][:m])
    print("bridges ok", int(br.sum()))
elif what == "text":
    B = 100_000
    p = pt.Packer(); p.add_generated(B, 100, 20, SEED, 0)
    text = ("\n".join(h + "\n" + g for h, g in (p.newick(i) for i in range(B))) + "\n").encode()
    t = torch.frombuffer(bytearray(text), dtype=torch.uint8).cuda()
    for _ in range(2):
        b = pt.Batch.from_newick_text(t); b.free()
    torch.cuda.synchronize(); print("text ok")
