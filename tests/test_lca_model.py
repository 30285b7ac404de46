"""The index arithmetic of csrc/lca_kernels.cu, mirrored in numpy (tests/lca_model.py), against the oracle."""
import numpy as np

import lca_model as M


def _tree(rng, n, kind):
    par = np.full(n, -1, dtype=np.int32)
    for v in range(1, n):
        par[v] = {"random": rng.integers(0, v), "path": v - 1, "star": 0, "cat": v - 1 if v % 2 else max(v - 2, 0)}[kind]
    perm = rng.permutation(n)
    out = np.full(n, -1, dtype=np.int32)
    for v in range(n):
        out[perm[v]] = perm[par[v]] if par[v] >= 0 else -1
    return out


def test_model_matches_oracle(oracle):
    rng = np.random.default_rng(3)
    for kind in ("random", "path", "star", "cat"):
        for n in (1, 2, 31, 32, 33, 65, 500, 2100):
            par = _tree(rng, n, kind)
            S = M.build(par)
            u, v = rng.integers(0, n, 400).astype(np.int32), rng.integers(0, n, 400).astype(np.int32)
            exp = oracle.lca_naive(par, u, v)
            got = np.array([M.query(S, int(a), int(b)) for a, b in zip(u, v)])
            assert (got == exp).all(), (kind, n)
