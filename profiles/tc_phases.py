"""profiles/tc_phases.py -- where a pt_batch_contain step spends its time on the persistent kernel: main phase (until the batch
ticket runs out) and tail (until the last warp leaves), from the %globaltimer stamps the kernel leaves in its control block
(pt_debug_last_phases), for the three instance kinds and the three input layouts.  Usage: python profiles/tc_phases.py [B]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import phylo_tools_b200 as pt

B = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
pt.set_device(0)
L = pt.lib()
for kind, name in ((0, "displayed"), (1, "labels swapped"), (2, "random guest")):
    p = pt.Packer(); p.add_generated(B, 100, 20, 0xB200 + 7 * kind, kind)
    for layout in (1, 2, 0):
        a = p.arrays(compact=layout)
        dev = {k: (torch.from_numpy(v).cuda() if isinstance(v, np.ndarray) else v) for k, v in a.items()}
        bits = torch.zeros((B + 31) // 32, dtype=torch.int32, device="cuda"); stat = torch.zeros(B, dtype=torch.uint8, device="cuda")
        b = pt.Batch(dev)
        for _ in range(3):
            b.contain(result_bits=bits, status=stat, want_counters=False)
        ms = []
        ph = (C.c_double * 4)()
        tails, mains = [], []
        for _ in range(5):
            b.contain(result_bits=bits, status=stat, want_counters=False)
            ms.append(b.round_ms(0))
            L.pt_debug_last_phases(b._h, ph)
            mains.append(ph[0]); tails.append(ph[1])
        _, _, c = b.contain(result_bits=bits, status=stat, want_counters=True)
        print(f"kind {kind} ({name:15s}) layout {layout or 4}: kernel {np.mean(ms):.3f} ms = main {np.mean(mains)/1e3:.3f} + tail {np.mean(tails)/1e3:.3f} ms; "
              f"pool slots {int(ph[2])}, items {c['work_items']}, branched {c['branched']}, displayed {c['displayed']}, errors {c['errors']}", flush=True)
        b.free()
