"""The N > 1 path on real GPUs: one process per GPU, the library's own NCCL collective (pt_comm_*, pt_batch_contain_sharded).
Needs at least two devices; on a one-GPU box the world-size-1 form of the same call is checked instead (no NCCL)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_call_with_one_rank(pt):
    import torch
    p = pt.Packer()
    for kind in (0, 1, 2):
        p.add_generated(333, 14, 6, 9100 + kind, kind)
    total = len(p)
    ref = pt.Batch(p.arrays())
    v, s, c = ref.contain()
    ref.free()
    comm = pt.Comm(1, 0)
    b = pt.Batch(p.arrays(compact=1))
    bits = torch.full(((total + 31) // 32,), -1, dtype=torch.int32, device="cuda")
    stat = torch.full(((total + 3) // 4 * 4,), 9, dtype=torch.uint8, device="cuda")
    cnt = comm.contain_sharded(b, 0, total, bits, stat, want_counters=True)
    got = np.unpackbits(bits.cpu().numpy().view(np.uint8), bitorder="little")[:total].astype(bool)
    assert (got == v).all() and (stat[:total].cpu().numpy() == s).all()
    assert cnt["displayed"] == c["displayed"] and cnt["instances"] == total
    with pytest.raises(pt.PtError):
        comm.contain_sharded(b, 16, total + 16, bits, stat)   # a shard must start on a word boundary
    b.free()
    comm.free()


@pytest.mark.parametrize("exchange", ["peer", "nccl"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_containment_over_nccl(pt, world, exchange):
    """both forms of the result exchange: through peer memory (CUDA IPC mappings over NVLink, the default) and through NCCL (PT_COMM_P2P=0)"""
    if pt.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, PT_COMM_P2P="1" if exchange == "peer" else "0")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tests", "multi_worker.py")], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "MULTI_OK world=%d" % world in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
