"""profiles/sweep_warps.py -- round-0 kernel time of the containment batch (BASELINE config 2) against the number of resident
warps per SM (PT_TC_MAX_WARPS, a tuning aid read by pt_batch_contain)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import phylo_tools_b200 as pt

B = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
pt.set_device(0)
p = pt.Packer(); p.add_generated(B, 100, 20, 0xB200, 0)
a = p.arrays()
dev = {k: (torch.from_numpy(v).cuda() if isinstance(v, np.ndarray) else v) for k, v in a.items()}
bits = torch.zeros((B + 31) // 32, dtype=torch.int32, device="cuda"); stat = torch.zeros(B, dtype=torch.uint8, device="cuda")
b = pt.Batch(dev)
for w in [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else "16,20,24,28,32".split(","))]:
    os.environ["PT_TC_MAX_WARPS"] = str(w)
    ms = []
    for _ in range(4):
        b.contain(result_bits=bits, status=stat, want_counters=False); torch.cuda.synchronize()
        ms.append(b.round_ms(0))
    print(f"max_warps {w:2d}: round 0 {min(ms[1:]):.3f} ms  (all {['%.3f' % x for x in ms]}) displayed {int(sum(bin(int(x) & 0xffffffff).count('1') for x in bits.cpu().numpy()))}")
