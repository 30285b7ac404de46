"""profiles/e2e_timing.py -- host wall time of the three calls of the end-to-end path (pinned HOST arrays, int16 layout, 10^5
instances of BASELINE config 2).  Measured: pt_batch_create 0.40 ms (33 async copies + streams/events), pt_batch_contain 8.33 ms
(resident step: 7.62 ms), pt_batch_free 0.12 ms."""
import os, sys, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import phylo_tools_b200 as pt
B = 100_000
pt.set_device(0)
p = pt.Packer(); p.add_generated(B, 100, 20, 0xB200, 0)
a = p.arrays(compact=True)
pinned = {k: (torch.from_numpy(v).pin_memory().numpy() if isinstance(v, np.ndarray) else v) for k, v in a.items()}
words = (B + 31) // 32
h_bits = torch.zeros(words, dtype=torch.int32).pin_memory(); h_stat = torch.zeros(B, dtype=torch.uint8).pin_memory()
opt = pt.Options(0, 0)
T = {"create": [], "contain": [], "free": []}
for it in range(8):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); b = pt.Batch(pinned); t1 = time.perf_counter()
    rc = pt.lib().pt_batch_contain(b._h, ctypes.byref(opt), h_bits.data_ptr(), h_stat.data_ptr(), pt.PT_MEM_HOST, None, None); t2 = time.perf_counter()
    b.free(); t3 = time.perf_counter()
    if it >= 3:
        T["create"].append(t1 - t0); T["contain"].append(t2 - t1); T["free"].append(t3 - t2)
for k, v in T.items(): print(k, "%.3f ms" % (1e3 * sum(v) / len(v)))
