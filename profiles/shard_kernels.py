"""profiles/shard_kernels.py -- the kernel time of every rank's shard of BASELINE config 2 at N ranks, one after the other on ONE GPU
(the sharded step follows the slowest rank: a straggler in any shard is everybody's step time).  Usage: python profiles/shard_kernels.py [ranks]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import phylo_tools_b200 as pt

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
total = 100_000
pt.set_device(0)
rows = []
for rank in range(world):
    cuts = [min(total, (total * r // world + 31) // 32 * 32) for r in range(world)] + [total]     # bench.py make_shard
    lo, hi = cuts[rank], cuts[rank + 1]
    p = pt.Packer()
    p.add_generated(hi - lo, 100, 20, 0xB200 + lo, 0)
    a = p.arrays(compact=1)
    dev = {k: (torch.from_numpy(v).cuda() if isinstance(v, np.ndarray) else v) for k, v in a.items()}
    b = pt.Batch(dev)
    bits = torch.zeros((hi - lo + 31) // 32, dtype=torch.int32, device="cuda"); stat = torch.zeros(hi - lo, dtype=torch.uint8, device="cuda")
    ts = []
    for it in range(6):
        b.contain(result_bits=bits, status=stat, want_counters=False); torch.cuda.synchronize()
        if it >= 2:
            ts.append(b.round_ms(0))
    _, _, c = b.contain(result_bits=bits, status=stat)
    rows.append((rank, lo, hi, min(ts), max(ts), c["work_items"], c["branched"]))
    b.free()
for r in rows:
    print("rank %d [%d, %d): kernel %.4f .. %.4f ms, work items %d, more than the rules: %d" % r)
print("slowest shard %.4f ms, mean %.4f ms" % (max(r[3] for r in rows), sum(r[3] for r in rows) / len(rows)))
