"""profiles/e2e_timing.py -- host wall time of the calls of the end-to-end path (pinned HOST arrays, uint8 layout, BASELINE config 2):
pt_batch_create (offset tables + chunked async copies), pt_batch_contain (DEVICE outputs: memsets + ONE launch, no wait), the D2H of
the result words + stream sync, pt_batch_free.  Usage: python profiles/e2e_timing.py [instances]"""
import os, sys, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import phylo_tools_b200 as pt
B = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
pt.set_device(0)
p = pt.Packer(); p.add_generated(B, 100, 20, 0xB200, 0)
a = p.arrays(compact=1)
pinned = {k: (torch.from_numpy(v).pin_memory().numpy() if isinstance(v, np.ndarray) else v) for k, v in a.items()}
words = (B + 31) // 32
d_bits = torch.zeros(words, dtype=torch.int32, device="cuda"); h_bits = torch.zeros(words, dtype=torch.int32).pin_memory()
opt = pt.Options(0, 0)
stream = torch.cuda.current_stream().cuda_stream
T = {"create": [], "contain": [], "d2h+sync": [], "free": [], "total": []}
for it in range(40):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); b = pt.Batch(pinned, stream=stream); t1 = time.perf_counter()
    rc = pt.lib().pt_batch_contain(b._h, ctypes.byref(opt), d_bits.data_ptr(), None, pt.PT_MEM_DEVICE, None, stream); t2 = time.perf_counter()
    h_bits.copy_(d_bits, non_blocking=True); torch.cuda.current_stream().synchronize(); t3 = time.perf_counter()
    b.free(); t4 = time.perf_counter()
    assert rc == 0
    if it >= 10:
        for k, v in zip(T, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t4 - t0)): T[k].append(v)
print("instances", B, " ".join("%s %.1f us" % (k, 1e6 * sorted(v)[len(v) // 2]) for k, v in T.items()))
