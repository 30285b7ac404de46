"""tests/lca_model.py -- TEST INFRASTRUCTURE: a numpy model of the data structure built by
phylo_tools_b200/csrc/lca_kernels.cu (preorder positions, 32-wide blocks, per-node records with stack masks,
sparse table over packed block minima).  It mirrors the kernels step by step so that the index arithmetic can be
checked against the oracle on a machine without a GPU.  Not used by the product."""
import numpy as np

INF = np.uint64(0xFFFFFFFFFFFFFFFF)


def build(parent):
    parent = np.asarray(parent, dtype=np.int64)
    n = len(parent)
    root = int(np.where(parent < 0)[0][0])
    kids = [[] for _ in range(n)]
    for v in range(n):
        if parent[v] >= 0:
            kids[parent[v]].append(v)
    # BFS order + depth (lca_sweeps_kernel, sweep 1)
    order, depth = [root], np.zeros(n, dtype=np.int64)
    i = 0
    while i < len(order):
        v = order[i]; i += 1
        for c in kids[v]:
            depth[c] = depth[v] + 1; order.append(c)
    size = np.ones(n, dtype=np.int64)
    for v in reversed(order):          # sweep 2
        for c in kids[v]:
            size[v] += size[c]
    pre = np.zeros(n, dtype=np.int64)
    for v in order:                    # sweep 3
        run = pre[v] + 1
        for c in kids[v]:
            pre[c] = run; run += size[c]
    keypos = np.zeros(n, dtype=np.uint64)
    keypos[pre] = (depth.astype(np.uint64) << np.uint64(32)) | (parent.astype(np.int64) & 0xFFFFFFFF).astype(np.uint64)
    nb = (n + 31) // 32
    pad = np.full(nb * 32, INF, dtype=np.uint64); pad[:n] = keypos
    blk = pad.reshape(nb, 32)
    prefix = np.minimum.accumulate(blk, axis=1)
    sufi = np.minimum.accumulate(blk[:, ::-1], axis=1)[:, ::-1]
    suf = np.full_like(blk, INF); suf[:, :31] = sufi[:, 1:]
    mask = np.zeros((nb, 32), dtype=np.uint64)
    for r in range(32):                # lca_blocks_kernel: bit j of mask[r] iff key[j] <= min(key[j+1..r])
        run = np.full(nb, INF, dtype=np.uint64)
        for j in range(r, -1, -1):
            hit = blk[:, j] <= run
            mask[:, r] |= np.where(hit, np.uint64(1) << np.uint64(j), np.uint64(0))
            run = np.where(hit, blk[:, j], run)
    table = [prefix[:, 31].copy()]
    k = 1
    while (1 << k) <= nb:
        prev = table[-1]; half = 1 << (k - 1)
        cnt = nb - (1 << k) + 1
        table.append(np.minimum(prev[:cnt], prev[half:half + cnt])); k += 1
    return dict(n=n, pre=pre, keypos=keypos, prefix=prefix.reshape(-1), suf=suf.reshape(-1), mask=mask.reshape(-1), table=table)


def query(S, u, v):
    if u == v:
        return u
    lp, rp = int(S["pre"][u]), int(S["pre"][v])
    if lp > rp:
        lp, rp = rp, lp
    bl, br = lp >> 5, rp >> 5
    if bl == br:
        m = int(S["mask"][rp]) >> ((lp & 31) + 1)
        j = (lp & 31) + 1 + ((m & -m).bit_length() - 1)
        best = int(S["keypos"][(bl << 5) + j])
    else:
        best = min(int(S["suf"][lp]), int(S["prefix"][rp]))
        if br - bl >= 2:
            k = (br - bl - 1).bit_length() - 1
            best = min(best, int(S["table"][k][bl + 1]), int(S["table"][k][br - (1 << k)]))
    return best & 0xFFFFFFFF
