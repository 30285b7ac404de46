"""profiles/prof_target.py tc|lca|enum -- a short run of ONE hot kernel for `ncu --set full` captures (B200_PROFILING.md).
Same kernels, same sizes as bench.py's legs, without the legs that are not being profiled."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import phylo_tools_b200 as pt

what = sys.argv[1] if len(sys.argv) > 1 else "tc"
pt.set_device(0)
if what == "tc":
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
    p = pt.Packer(); p.add_generated(B, 100, 20, 0xB200, 0)
    a = p.arrays(compact=True)   # the layout bench.py uses (int16 data arrays)
    dev = {k: (torch.from_numpy(v).cuda() if isinstance(v, np.ndarray) else v) for k, v in a.items()}
    bits = torch.zeros((B + 31) // 32, dtype=torch.int32, device="cuda"); stat = torch.zeros(B, dtype=torch.uint8, device="cuda")
    b = pt.Batch(dev)
    for _ in range(2):
        b.contain(result_bits=bits, status=stat, want_counters=False)
    torch.cuda.synchronize()
    print("tc ok", b.round_ms(0), "ms round 0")
elif what == "text":
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
    p = pt.Packer(); p.add_generated(B, 100, 20, 0xB200, 0)
    text = ("\n".join(h + "\n" + g for h, g in (p.newick(i) for i in range(B))) + "\n").encode()
    t = torch.frombuffer(bytearray(text), dtype=torch.uint8).cuda()
    for _ in range(2):
        b = pt.Batch.from_newick_text(t)
        b.free()
    torch.cuda.synchronize()
    print("text ok", len(text), "bytes")
elif what == "lca":
    leaves = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
    q = int(sys.argv[3]) if len(sys.argv) > 3 else 100_000_000
    par = pt.random_tree(leaves, 0xB200); n = len(par)
    L = pt.Lca(torch.from_numpy(par).cuda())
    g = torch.Generator(device="cuda").manual_seed(1)
    u = torch.randint(0, n, (q,), generator=g, device="cuda", dtype=torch.int32); v = torch.randint(0, n, (q,), generator=g, device="cuda", dtype=torch.int32)
    out = torch.empty_like(u)
    for _ in range(3):
        L.query(u, v, out=out)
    torch.cuda.synchronize()
    print("lca ok", L.info())
