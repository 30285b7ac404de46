"""Multi-GPU sharding of the batched containment path (host logic only; one process per GPU).

Instances are independent, so a batch is cut into contiguous ranges, one per rank, with NO data-path collective
(SURVEY 8(e)).  Range boundaries are multiples of 32 instances so that every rank owns whole words of the packed result
bit vector; the single collective of the path is then an all-reduce(SUM) of those words (disjoint words: SUM == OR) plus
the counters.  Used by bench.py under torch.distributed (NCCL on the GPUs, gloo in the CPU tests)."""
import numpy as np


def shard_ranges(weights, world_size, align=32):
    """contiguous [begin, end) per rank, balanced by the prefix sum of `weights` (e.g. n_i + m_i), boundaries aligned"""
    w = np.asarray(weights, dtype=np.int64)
    B = len(w)
    cum = np.concatenate([[0], np.cumsum(w)])
    total = int(cum[-1])
    cuts = [0]
    for r in range(1, world_size):
        target = total * r // world_size
        i = int(np.searchsorted(cum, target, side="left"))
        # interior cuts stay aligned even when rounding runs past B: only the LAST rank may own a ragged tail
        i = min(B // align * align, max(cuts[-1], (i + align // 2) // align * align))
        cuts.append(i)
    cuts.append(B)
    return [(cuts[r], cuts[r + 1]) for r in range(world_size)]


def place_words(local_bits_words, begin, total_instances):
    """a rank's packed result words placed into a zeroed full-length word vector (begin must be a multiple of 32 unless the
    rank's range is empty)"""
    full = np.zeros((total_instances + 31) // 32, dtype=np.int64)
    if len(local_bits_words) == 0:
        return full
    assert begin % 32 == 0
    w0 = begin // 32
    full[w0:w0 + len(local_bits_words)] = np.asarray(local_bits_words, dtype=np.uint32).astype(np.int64)
    return full


def merge_results(dist, local_words, begin, total_instances, counters=None, device=None):
    """the path's only collective: all-reduce(SUM) of the packed result words and of the counters.
    returns (uint32 words of the whole batch, summed counters dict)"""
    import torch
    full = torch.from_numpy(place_words(local_words, begin, total_instances))
    keys = sorted(counters) if counters else []
    cnt = torch.tensor([int(counters[k]) for k in keys], dtype=torch.int64)
    if device is not None:
        full, cnt = full.to(device), cnt.to(device)
    dist.all_reduce(full, op=dist.ReduceOp.SUM)
    if keys:
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    return full.cpu().numpy().astype(np.uint32), {k: int(c) for k, c in zip(keys, cnt.cpu().tolist())}
