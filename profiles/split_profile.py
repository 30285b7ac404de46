"""profiles/split_profile.py -- one pt_batch_contain on BASELINE config 4's shape (8 instances, n = 10^6, r = 10^3) through the bridge-splitting
path (pt_options.path = 5) and through the cluster executor (path = 3); used under ncu for the per-kernel launch list, and with
PT_SPLIT_TRACE=1 for the wall time of every phase of the split path.  Usage: python profiles/split_profile.py [instances] [path]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import phylo_tools_b200 as pt

count = int(sys.argv[1]) if len(sys.argv) > 1 else 8
paths = [int(sys.argv[2])] if len(sys.argv) > 2 else [5, 3]
pt.set_device(0)
p = pt.Packer(); p.add_generated(count, 499_001, 1000, 0xB200 + 4, 0)
a = p.arrays()
dev = {k: (torch.from_numpy(v).cuda() if isinstance(v, np.ndarray) else v) for k, v in a.items()}
bits = torch.zeros((count + 31) // 32, dtype=torch.int32, device="cuda"); stat = torch.zeros(count, dtype=torch.uint8, device="cuda")
b = pt.Batch(dev)
for path in paths:
    b.contain(result_bits=bits, status=stat, want_counters=False, path=path); torch.cuda.synchronize()
    t0 = time.perf_counter()
    _, _, c = b.contain(result_bits=bits, status=stat, want_counters=True, path=path); torch.cuda.synchronize()
    print(f"path {path}: {count} instances in {(time.perf_counter() - t0) * 1e3:.2f} ms, displayed {c['displayed']}, work items {c['work_items']}", flush=True)
b.free()
