"""profiles/multi_one_process.py [instances] -- pt_batch_contain_multi: one process, one host thread per GPU (pinned HOST arrays,
int16 layout), against the same batch on one device."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import phylo_tools_b200 as pt
B = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
p = pt.Packer(); p.add_generated(B, 100, 20, 0xB200, 0)
a = p.arrays(compact=True)
a = {k: (torch.from_numpy(v).pin_memory().numpy() if isinstance(v, np.ndarray) else v) for k, v in a.items()}
for g in [1] + ([pt.device_count()] if pt.device_count() > 1 else []):
    pt.contain_multi(a, num_devices=g)
    t0 = time.perf_counter(); v, s, c = pt.contain_multi(a, num_devices=g); dt = time.perf_counter() - t0
    print(f"{g} device(s): {B / dt / 1e6:.2f} M instances/s end to end ({dt * 1e3:.1f} ms), displayed {int(v.sum())} of {B}, errors {int((s != 0).sum())}")
