"""profiles/phase_profile.py [B] -- cycles per solver pass of tc_reduce_kernel, from the PT_PROFILE build
(make -C phylo_tools_b200/csrc prof).  Profiling aid only: the probes cost time, so the split matters, not the total."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["PT_GPU_LIB"] = os.path.join(ROOT, "phylo_tools_b200", "libptgpu_prof.so")
sys.path.insert(0, ROOT)
import numpy as np, torch
import phylo_tools_b200 as pt
from phylo_tools_b200 import _capi

NAMES = {0: "idle/ticket", 1: "load: arrays", 15: "load: guest checks", 16: "load: label tables", 2: "clean_guest", 3: "csr: compact arcs", 31: "csr: renumber nodes", 32: "csr: zero+count+offsets", 33: "csr: scan", 34: "csr: fill in-lists",
         71: "true cherry: collapse", 72: "true cherry: parents", 7: "true cherry: queue/overhead", 5: "check_dag", 40: "clean: detect", 41: "clean: dead ends", 43: "clean: suppress",
         44: "clean: merge chains", 46: "clean: parallel arcs", 14: "count_labels", 8: "guest cherries", 9: "ret cherry",
         6: "wavefront", 17: "small core: setup", 18: "small core: switchings", 10: "sibling", 11: "pick branch", 12: "emit child", 13: "epilogue"}
B = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
pt.set_device(0)
p = pt.Packer(); p.add_generated(B, 100, 20, 0xB200, 0)
a = p.arrays()
dev = {k: (torch.from_numpy(v).cuda() if isinstance(v, np.ndarray) else v) for k, v in a.items()}
bits = torch.zeros((B + 31) // 32, dtype=torch.int32, device="cuda"); stat = torch.zeros(B, dtype=torch.uint8, device="cuda")
b = pt.Batch(dev)
L = pt._L
buf = (C.c_ulonglong * 128)()
b.contain(result_bits=bits, status=stat, want_counters=False); torch.cuda.synchronize()
L.pt_debug_profile(buf)  # drop the warm-up
b.contain(result_bits=bits, status=stat, want_counters=False); torch.cuda.synchronize()
L.pt_debug_profile(buf)
cyc = list(buf)[:64]; ent = list(buf)[64:]
tot = sum(cyc)
print(f"B={B} round0 {b.round_ms(0):.3f} ms (with probes); warp-cycles total {tot:.3e}")
print(f"{'pass':24s} {'share':>7s} {'entries/inst':>13s} {'cycles/entry':>13s}")
for k in sorted(range(64), key=lambda k: -cyc[k]):
    if cyc[k] == 0: continue
    print(f"{NAMES.get(k, str(k)):24s} {cyc[k]/tot:7.3f} {ent[k]/B:13.2f} {cyc[k]/max(ent[k],1):13.0f}")
