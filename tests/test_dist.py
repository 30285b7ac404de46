"""The N>1 path on CPU: two gloo ranks shard a batch, each 'solves' its range, and the single all-reduce of packed
result words + counters reproduces the single-rank result.  (The per-instance verdict here comes from a stand-in
function of the instance index -- the GPU kernels are covered by the -m gpu tests; this test is about the host logic.)"""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from phylo_tools_b200.shard import merge_results, shard_ranges  # noqa: E402


def _fake_verdict(i):
    return (i * 2654435761 >> 7) & 1


def _worker(rank, world, port, B, weights, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = shard_ranges(weights, world)[rank]
    verdicts = np.array([_fake_verdict(i) for i in range(b, e)], dtype=np.uint8)
    words = np.packbits(np.pad(verdicts, (0, (-len(verdicts)) % 32)), bitorder="little").view(np.uint32)
    full, cnt = merge_results(dist, words, b, B, counters=dict(instances=e - b, displayed=int(verdicts.sum())))
    if rank == 0:
        np.save(out, full)
        assert cnt["instances"] == B
        assert cnt["displayed"] == sum(_fake_verdict(i) for i in range(B))
    dist.destroy_process_group()


def test_shard_ranges_are_aligned_and_cover():
    rng = np.random.default_rng(0)
    for B in (0, 1, 31, 32, 33, 1000, 12345):
        w = rng.integers(1, 500, B)
        for world in (1, 2, 4, 8):
            r = shard_ranges(w, world)
            assert r[0][0] == 0 and r[-1][1] == B
            for (b0, e0), (b1, _) in zip(r, r[1:]):
                assert e0 == b1 and b0 <= e0 and b1 % 32 == 0 or b1 == B
            if B >= 2000:
                loads = [int(w[b:e].sum()) for b, e in r]
                assert max(loads) < 1.3 * (sum(loads) / world) + 32 * 500


def test_only_the_last_rank_owns_a_ragged_tail():
    """B not a multiple of 32 and more ranks than 32-instance words: every non-empty range starts on a word boundary, so the
    all-reduce SUM of the placed words is an OR (no two ranks share a word)"""
    from phylo_tools_b200.shard import place_words
    for B, world in ((49, 8), (33, 4), (95, 8), (31, 2)):
        r = shard_ranges(np.ones(B), world)
        assert r[0][0] == 0 and r[-1][1] == B
        owners = np.zeros((B + 31) // 32, dtype=int)
        for b, e in r:
            if e > b:
                assert b % 32 == 0, (B, world, r)
                owners[b // 32:(e + 31) // 32] += 1
            place_words(np.zeros((e - b + 31) // 32, dtype=np.uint32), b, B)   # empty ranges are accepted at any begin
        assert (owners == 1).all(), (B, world, r)


def test_two_rank_allreduce_merge(tmp_path):
    B = 1000
    weights = np.random.default_rng(1).integers(100, 600, B)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "full.npy")
    mp.spawn(_worker, args=(2, port, B, weights, out), nprocs=2, join=True)
    full = np.load(out)
    bits = np.unpackbits(full.view(np.uint8), bitorder="little")[:B]
    assert bits.tolist() == [_fake_verdict(i) for i in range(B)]
