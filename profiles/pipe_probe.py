import os, sys, time, ctypes
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import phylo_tools_b200 as pt
B = int(sys.argv[1]) if len(sys.argv) > 1 else 12500
pt.set_device(0)
p = pt.Packer(); p.add_generated(B, 100, 20, 0xB200, 0)
a = p.arrays(compact=1)
pin = [{k: (torch.from_numpy(v.copy()).pin_memory().numpy() if isinstance(v, np.ndarray) else v) for k, v in a.items()} for _ in range(2)]
words = (B + 31) // 32
d_bits = torch.zeros(words, dtype=torch.int32, device="cuda")
opt = pt.Options(0, 0)
stream = torch.cuda.current_stream().cuda_stream
L = pt.lib()
def contain(b):
    return L.pt_batch_contain(b._h, ctypes.byref(opt), d_bits.data_ptr(), None, pt.PT_MEM_DEVICE, None, stream)
for it in range(6):
    torch.cuda.synchronize()
    b0 = pt.Batch(pin[0], stream=stream); contain(b0)
    t0 = time.perf_counter(); b1 = pt.Batch(pin[1], stream=stream); t1 = time.perf_counter(); contain(b1); t2 = time.perf_counter()
    torch.cuda.synchronize(); t3 = time.perf_counter()
    b0.free(); b1.free()
    print("while a kernel runs: create %.1f us, contain %.1f us, then drain %.1f us" % ((t1 - t0) * 1e6, (t2 - t1) * 1e6, (t3 - t2) * 1e6))
