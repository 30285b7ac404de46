"""profiles/negatives.py -- throughput of the containment batch on NEGATIVE instances of BASELINE config 2's shape (the headline
is measured on displayed trees, as the north star asks): kind 1 = displayed tree with two leaf labels swapped, kind 2 = random
tree on the same labels."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import phylo_tools_b200 as pt

B = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
pt.set_device(0)
for kind, name in ((0, "displayed by construction"), (1, "two leaf labels swapped"), (2, "random guest tree")):
    p = pt.Packer(); p.add_generated(B, 100, 20, 0xB200 + 7 * kind, kind)
    a = p.arrays(compact=True)
    dev = {k: (torch.from_numpy(v).cuda() if isinstance(v, np.ndarray) else v) for k, v in a.items()}
    bits = torch.zeros((B + 31) // 32, dtype=torch.int32, device="cuda"); stat = torch.zeros(B, dtype=torch.uint8, device="cuda")
    b = pt.Batch(dev)
    b.contain(result_bits=bits, status=stat, want_counters=False); torch.cuda.synchronize()
    t0 = time.perf_counter()
    _, _, c = b.contain(result_bits=bits, status=stat, want_counters=True); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"kind {kind} ({name}): {B / dt / 1e6:.2f} M instances/s  ({dt * 1e3:.2f} ms)  displayed {c['displayed']}  branched {c['branched']}  work items {c['work_items']}  rounds {c['rounds']}  errors {c['errors']}")
    b.free()
