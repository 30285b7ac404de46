"""worker of tests/test_gpu_multi.py (one process per GPU under torch.distributed.run): the sharded containment step through the
library's own collective (pt_batch_contain_sharded) must leave, on EVERY rank, the verdict words / status bytes / counters a
single-GPU solve of the whole batch gives."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import phylo_tools_b200 as pt  # noqa: E402


def lo_of(index, tot, world):
    r = min(index * world // tot, world - 1)
    return tot * r // world


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    pt.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def bootstrap(ident):
        box = [ident]
        dist.broadcast_object_list(box, src=0)
        return box[0]
    comm = pt.Comm(world, rank, bootstrap)
    want_p2p = os.environ.get("PT_COMM_P2P", "1") != "0"
    assert comm.peer_exchange() or not want_p2p or os.environ.get("PT_ALLOW_NO_PEER_ACCESS"), "the GPUs of this box should map each other's memory"
    assert want_p2p or not comm.peer_exchange()
    total = 5003   # not a multiple of 32: the last rank owns a ragged tail
    kinds = [(0, 1700), (1, 1700), (2, 1603)]
    # the whole batch (every rank makes it, so that each can check the merged result against a local solve of everything)
    whole = pt.Packer()
    for kind, cnt in kinds:
        whole.add_generated(cnt, 14, 6, 9000 + kind, kind)
    ref_b = pt.Batch(whole.arrays())
    v_ref, s_ref, c_ref = ref_b.contain()
    ref_b.free()
    cuts = [min(total, (total * r // world + 31) // 32 * 32) for r in range(world)] + [total]
    first, count = cuts[rank], cuts[rank + 1] - cuts[rank]
    a = whole.arrays(compact=1)
    # cut this rank's range out of the flat arrays
    hn, ha, gn = a["host_node_off"], a["host_arc_off"], a["guest_node_off"]
    sl = dict(num_instances=count, index_bytes=1, host_node_off=hn[first:first + count + 1] - hn[first], host_arc_off=ha[first:first + count + 1] - ha[first],
              guest_node_off=gn[first:first + count + 1] - gn[first], host_succ_off=a["host_succ_off"][hn[first]:hn[first + count]],
              host_succ_adj=a["host_succ_adj"][ha[first]:ha[first + count]], host_label=a["host_label"][hn[first]:hn[first + count]],
              guest_parent=a["guest_parent"][gn[first]:gn[first + count]], guest_label=a["guest_label"][gn[first]:gn[first + count]])
    sl = {k: (np.ascontiguousarray(v) if isinstance(v, np.ndarray) else v) for k, v in sl.items()}
    b = pt.Batch(sl)
    full_bits = torch.full(((total + 31) // 32,), -1, dtype=torch.int32, device="cuda")     # poisoned: the call must overwrite all of it
    full_stat = torch.full(((total + 3) // 4 * 4,), 77, dtype=torch.uint8, device="cuda")
    for it in range(3):
        cnt = comm.contain_sharded(b, first, total, full_bits, full_stat, want_counters=True)
        torch.cuda.synchronize()
        got = np.unpackbits(full_bits.cpu().numpy().view(np.uint8), bitorder="little")[:total].astype(bool)
        assert (got == v_ref).all(), (rank, it, np.where(got != v_ref)[0][:10])
        assert (full_stat[:total].cpu().numpy() == s_ref).all()
        for k in ("instances", "displayed", "not_displayed", "errors", "branched", "resolved_by_rules"):
            assert cnt[k] == c_ref[k], (k, cnt[k], c_ref[k])
    b.free()
    # the generic all-reduce (the enumerator's found word): max over ranks
    w = torch.tensor([rank + 1, 0], dtype=torch.int64, device="cuda")
    comm.allreduce(w, op="max")
    torch.cuda.synchronize()
    assert w.cpu().tolist() == [world, 0]
    # the enumerator's cross-GPU early exit: every rank enumerates its own slice of the index space; the rank
    # whose slice holds the first witness finds it, the others learn it from the all-reduced found word and stop early
    pe = pt.Packer()
    pe.add_generated(1, 14, 16, 777, 0)       # r = 16: 65536 switchings, displayed by construction
    ea = pe.arrays()
    args = (ea["host_succ_off"], ea["host_succ_adj"], ea["host_label"], ea["guest_parent"], ea["guest_label"])
    tot, nd_all, first_all = pt.enum_switchings(*args)
    assert first_all is not None
    E = pt.Enum(*args)
    lo, hi = tot * rank // world, tot * (rank + 1) // world
    nd, first, early = E.run(begin=lo, end=hi, stop_at_first=True, comm=comm, chunk=256)
    owner = min(first_all * world // tot, world - 1)
    if rank == owner:
        assert first == first_all, (first, first_all)
    t = torch.tensor([1 if early else 0, 1 if first is not None else 0], dtype=torch.int64, device="cuda")
    dist.all_reduce(t)
    assert int(t[1].item()) >= 1                       # somebody has a witness
    if tot // world > 4 * 256 and first_all - lo_of(first_all, tot, world) + 256 < tot // world:
        assert int(t[0].item()) >= world - 1, t        # everybody else stopped before its slice was exhausted
    # without a witness anywhere (random guest) every rank runs to the end and nobody reports an early stop
    pn = pt.Packer()
    pn.add_generated(1, 14, 12, 778, 2)
    na = pn.arrays()
    nargs = (na["host_succ_off"], na["host_succ_adj"], na["host_label"], na["guest_parent"], na["guest_label"])
    tot2, nd2, first2 = pt.enum_switchings(*nargs)
    E2 = pt.Enum(*nargs)
    r2 = E2.run(begin=tot2 * rank // world, end=tot2 * (rank + 1) // world, stop_at_first=True, comm=comm, chunk=512)
    t2 = torch.tensor([r2[0]], dtype=torch.int64, device="cuda")
    dist.all_reduce(t2)
    assert int(t2.item()) == nd2 and (nd2 > 0 or not r2[2])
    E.free(); E2.free()
    comm.free()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_OK world=%d displayed=%d" % (world, int(v_ref.sum())))


if __name__ == "__main__":
    main()
