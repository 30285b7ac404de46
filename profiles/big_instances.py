"""profiles/big_instances.py [leaves] [retis] [count] -- BASELINE config 4 shape (large networks, small batch) through the
global-memory executor; prints generation and solve times."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import phylo_tools_b200 as pt
l = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
r = int(sys.argv[2]) if len(sys.argv) > 2 else 100
cnt = int(sys.argv[3]) if len(sys.argv) > 3 else 8
kind = int(sys.argv[4]) if len(sys.argv) > 4 else 0
pt.set_device(0)
t0 = time.time(); p = pt.Packer(); p.add_generated(cnt, l, r, 4, kind); a = p.arrays(); t1 = time.time()
dev = {k: (torch.from_numpy(v).cuda() if isinstance(v, np.ndarray) else v) for k, v in a.items()}
bits = torch.zeros((cnt + 31) // 32, dtype=torch.int32, device="cuda"); stat = torch.zeros(cnt, dtype=torch.uint8, device="cuda")
b = pt.Batch(dev)
for rep in range(2):
    torch.cuda.synchronize(); t2 = time.time()
    _, _, c = b.contain(result_bits=bits, status=stat, want_counters=True)
    torch.cuda.synchronize(); t3 = time.time()
    print(dict(leaves=l, retis=r, nodes=2 * r + 2 * l - 1, count=cnt, kind=kind, gen_s=round(t1 - t0, 2), solve_s=round(t3 - t2, 4), inst_per_s=round(cnt / (t3 - t2), 2),
               displayed=c["displayed"], errors=c["errors"], work_items=c["work_items"], rounds=c["rounds"], rule_passes=c["rule_passes"]), flush=True)
