"""oracle/make_golden_r02.py -- regenerates the config-2-sized containment fixtures IN THE BUILD CONTAINER (needs the reference
through oracle/_ref/ref_tc; the GPU box has neither the reference nor this binary, the committed JSON travels).

  tests/golden/tc_ref_l100.json   REFERENCE verdicts (oracle/_ref/ref_tc verdicts = TreeInNetContainment::displayed() of the
        unmodified headers, utils/containment.hpp:1202) on seeded generator batches at the headline size: 300 instances x kinds
        0/1/2 at l=100, r=20 and 50 x 3 at l=300, r=60.  Only (l, r, seed, kind, count) and the verdict string are stored: the
        instances are re-made by the seeded packer (pt_packer_add_generated), which is deterministic.
        Contract: the reference is never wrong on "not displayed", so '0' must be matched; '1' must be matched or
        be settled by exhaustive switching; every such disagreement found at generation time (by the CPU simulator of the
        kernel logic, tests/host_sim) is listed under "settled" with its exhaustive truth.
  tests/golden/tc_exh_l100.json   exhaustive-switching truth (oracle/pt_oracle.c, all 2^20 switchings unless one displays) of a
        seeded UNIFORM sample of kinds 1 (two labels swapped) and 2 (random guest) at l=100, r=20 -- not a solver-chosen subset.
"""
import json
import os
import subprocess
import sys
import tempfile
from concurrent.futures import ProcessPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import phylo_tools_b200 as pt  # noqa: E402  (host-side generator only; no GPU needed)
from oracle import pt_oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
REF_TC = os.path.join(ROOT, "oracle", "_ref", "ref_tc")
SIM_SRC = os.path.join(ROOT, "tests", "host_sim", "sim_tc.cpp")
SIM_BIN = "/tmp/sim_tc_fast"

REF_CONFIGS = [(100, 20, 777000 + 1000 * k, k, 300) for k in (0, 1, 2)] + [(300, 60, 778000 + 1000 * k, k, 50) for k in (0, 1, 2)]
EXH_CONFIGS = [(100, 20, 779001, 1, 80), (100, 20, 779002, 2, 80)]


def pairs_of(l, r, seed, kind, count):
    p = pt.Packer()
    p.add_generated(count, l, r, seed, kind)
    return [p.newick(i) for i in range(count)]


def ref_verdicts(pairs):
    with tempfile.NamedTemporaryFile("w", suffix=".nw", delete=False) as f:
        for h, g in pairs:
            f.write(h + "\n" + g + "\n")
    out = subprocess.run([REF_TC, "verdicts", f.name], capture_output=True, text=True).stdout.strip()
    os.unlink(f.name)
    assert len(out) == len(pairs), (len(out), len(pairs))
    return out


def brute(pair):
    return int(O.tc_bruteforce_newick(pair[0], pair[1], stop_at_first=True, max_switchings=1 << 62)[0])


def one_ref_config(cfg):
    l, r, seed, kind, count = cfg
    pairs = pairs_of(l, r, seed, kind, count)
    ref = ref_verdicts(pairs)
    sim = subprocess.run([SIM_BIN, "gen", str(count), str(l), str(r), str(seed), str(kind), "0"], capture_output=True, text=True).stdout.strip()
    return cfg, pairs, ref, sim


if __name__ == "__main__":
    subprocess.check_call(["g++", "-std=c++17", "-O2", SIM_SRC, "-o", SIM_BIN])
    with ProcessPoolExecutor(6) as ex:
        res = list(ex.map(one_ref_config, REF_CONFIGS))
        configs, todo = [], []
        for (l, r, seed, kind, count), pairs, ref, sim in res:
            dis = [i for i in range(count) if ref[i] in "01" and ref[i] != sim[i]]
            hard = [i for i in dis if ref[i] == "0"]
            assert not hard, ("the kernel logic says displayed where the reference says not displayed", l, r, seed, kind, hard)
            configs.append(dict(leaves=l, reticulations=r, seed=seed, kind=kind, count=count, reference=ref, settled={}))
            todo += [(len(configs) - 1, i, pairs[i]) for i in dis if r <= 24]
            print("l=%d r=%d kind=%d: reference 1:%d 0:%d crashed:%d; kernel logic differs on %s" % (l, r, kind, ref.count("1"), ref.count("0"), count - ref.count("1") - ref.count("0"), dis), flush=True)
        for (c, i, _), t in zip(todo, ex.map(brute, [p for _, _, p in todo])):
            configs[c]["settled"][str(i)] = t
        json.dump(dict(source="oracle/_ref/ref_tc verdicts (reference utils/containment.hpp:1202) on pt_packer_add_generated(count, leaves, reticulations, seed, kind); "
                              "settled = exhaustive switching (oracle/pt_oracle.c) where the reference says displayed and the kernel logic does not",
                       configs=configs), open(os.path.join(GOLD, "tc_ref_l100.json"), "w"), indent=0)
        exh = []
        for (l, r, seed, kind, count) in EXH_CONFIGS:
            pairs = pairs_of(l, r, seed, kind, count)
            truth = list(ex.map(brute, pairs))
            exh.append(dict(leaves=l, reticulations=r, seed=seed, kind=kind, count=count, truth="".join(map(str, truth))))
            print("exhaustive l=%d r=%d kind=%d: %d displayed of %d" % (l, r, kind, sum(truth), count), flush=True)
        json.dump(dict(source="oracle/pt_oracle.c exhaustive switching on a uniform seeded sample: pt_packer_add_generated(count, leaves, reticulations, seed, kind)",
                       configs=exh), open(os.path.join(GOLD, "tc_exh_l100.json"), "w"), indent=0)
