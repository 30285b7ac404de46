#!/usr/bin/env python
"""bench.py -- the headline benchmark of BASELINE.json on B200.

metric   : tree-containment instances/s  (BASELINE.json "metric"; config = configs[1]: a batch of 10^5 `gen` networks
           l=100, r=20, each with a displayed tree)
step     : one pass of the hot path (pt_batch_contain) over one batch of 10^5 synthetic instances per GPU
value    : whole-job instances/s with the batch RESIDENT in HBM when the timed region starts (CUDA events, max over ranks)
e2e      : the same through the public C-ABI call with HOST (pinned) buffers: pt_batch_create (H2D) + pt_batch_contain
           (kernels + D2H of the result bits) + pt_batch_free, every step
roofline : dominant kernel = tc_reduce_kernel over the original instances (round 0), timed live with CUDA events inside
           the library; algorithmic bytes = 8124 B/instance; peak = MEASURED_PEAKS.json hbm_gbs
cpu_baseline / --impl reference : the reference's own TreeInNetContainment::displayed() (oracle/_ref/ref_tc, unmodified
           reference headers, std::cout silenced) on the box's host cores, one process per core over disjoint slices
lca      : secondary leg (BASELINE config 3): 10^8 LCA queries on a 10^7-leaf tree, outside the timed region

N > 1: one process per GPU (torchrun), instances sharded with no data-path collective; the only collective is the
all-reduce of the packed result words and counters (phylo_tools_b200/shard.py), inside the timed region. Weak scaling:
10^5 instances per GPU.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALG_BYTES_PER_INSTANCE = 8124  # int32 CSR host 4940 B + guest 3184 B at l=100, r=20
LCA_BYTES_PER_QUERY = 40
L, R = 100, 20
SEED = 0xB200


def bind_to_gpu_numa_node(local_rank):
    """one process per GPU: run on the CPU cores next to that GPU, so that the pinned staging buffers of the end-to-end path are
    first-touched on its NUMA node (8 ranks uploading through one node's memory halve the end-to-end rate)"""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * w + b for w in range(words) for b in range(64) if (mask[w] >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def measured_traffic(kernel, units):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/traffic_r01.json), only when the
    launch here processes the same number of units as the captured one; else None"""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic_r01.json")))[kernel]
        return t["dram_bytes_per_launch"] if t["units_per_launch"] == units else None
    except Exception:
        return None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock / throttle reasons during the timed region (B200_PROFILING.md recipe): NVML polled in-process, nvidia-smi as the fallback"""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.stop, self.t = index, [], False, None

    def _run_nvml(self):
        """NVML in-process: a sample every 2 ms (the timed region of the default run is ~80 ms, too short for nvidia-smi's ~100 ms per call)"""
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
        get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}   # nvmlClocksEventReason*
        while not self.stop:
            sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            r = get_reasons(h)
            self.rows.append([str(sm), str(mx)] + [("Active" if r & bits[n] else "Not Active") for n in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")])
            time.sleep(0.002)

    def _run(self):
        try:
            self._run_nvml()
            return
        except Exception:
            pass
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def __enter__(self):
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[2:6]) if v.strip().lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(self.rows)}


def _ref_json(stderr_text):
    for line in reversed((stderr_text or "").strip().split("\n")):
        line = line.strip()
        if line.startswith("{"):
            try:
                return json.loads(line)
            except Exception:
                pass
    return {}


def reference_arm(args, rank, world):
    """the reference's CPU implementation of the path on the host cores (rank 0 only)"""
    if rank != 0:
        return
    import phylo_tools_b200 as pt
    from oracle import pt_oracle as O
    ref = O.ref_binary("ref_tc")
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 64))
    per = args.ref_instances_per_core
    res = {"metric": "tree-containment instances/s", "unit": "instances/s", "impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
           "config": {"workload": "batch of 10^5 gen networks l=100 r=20 with a displayed tree each (BASELINE configs[1]); bounded sample per step",
                      "leaves": L, "reticulations": R}}
    if ref is None:
        res["unavailable"] = "oracle/_ref/ref_tc was not built (needs /root/reference at build time)"
        print(json.dumps(res))
        return
    p = pt.Packer()
    p.add_generated(procs * per, L, R, SEED, 0)
    files = []
    for k in range(procs):
        f = tempfile.NamedTemporaryFile("w", suffix=".nw", delete=False)
        for i in range(k * per, (k + 1) * per):
            h, g = p.newick(i)
            f.write(h + "\n" + g + "\n")
        f.close()
        files.append(f.name)

    def one_step():
        t0 = time.perf_counter()
        ps = [subprocess.Popen([ref, "bench", fn, str(per)], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True) for fn in files]
        outs = [q.communicate()[1] for q in ps]
        dt = time.perf_counter() - t0
        js = [_ref_json(o) for o in outs]
        return dt, sum(j.get("displayed", 0) for j in js), sum(j.get("displayed", 0) + j.get("not_displayed", 0) for j in js)
    for _ in range(args.warmup):
        one_step()
    times, disp, done = [], 0, 0
    for _ in range(args.steps):
        dt, disp, done = one_step()
        times.append(dt)
    for fn in files:
        os.unlink(fn)
    total = procs * per
    ms = 1000.0 * sum(times) / len(times)
    val = done / (ms / 1000.0)  # instances on which the reference terminated with a verdict
    res.update({"value": val, "ms_per_step": ms, "gpu_launches": 0,
                "cpu_baseline": {"value": val, "unit": "instances/s", "cores": procs, "kind": "reference",
                                 "sample": "%d instances per step (%d per core); reference gave a verdict on %d (%d displayed), crashed/threw on %d" % (total, per, done, disp, total - done)},
                "e2e": {"value": val, "unit": "instances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(res))


def cpu_baseline_leg(pt, sample_seconds=15.0):
    """bounded sample of the same workload through the reference on all host cores (reported, not the target)"""
    from oracle import pt_oracle as O
    ref = O.ref_binary("ref_tc")
    if ref is None:
        return {"value": None, "unit": "instances/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref/ref_tc not built"}
    cores = max(1, min(os.cpu_count() or 1, 64))
    per = max(8, int(sample_seconds * 45))  # ~50 instances/s/core survey-time
    per = min(per, 400)
    p = pt.Packer()
    p.add_generated(cores * per, L, R, SEED, 0)
    files = []
    for k in range(cores):
        f = tempfile.NamedTemporaryFile("w", suffix=".nw", delete=False)
        for i in range(k * per, (k + 1) * per):
            h, g = p.newick(i)
            f.write(h + "\n" + g + "\n")
        f.close()
        files.append(f.name)
    t0 = time.perf_counter()
    ps = [subprocess.Popen([ref, "bench", fn, str(per)], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True) for fn in files]
    outs = [q.communicate()[1] for q in ps]
    dt = time.perf_counter() - t0
    for fn in files:
        os.unlink(fn)
    js = [_ref_json(o) for o in outs]
    disp = sum(j.get("displayed", 0) for j in js)
    done = sum(j.get("displayed", 0) + j.get("not_displayed", 0) for j in js)
    crashed = cores * per - done
    single = [j.get("inst_per_s", 0.0) for j in js]
    return {"value": done / dt, "reference_crashes_or_exceptions": crashed, "unit": "instances/s", "cores": cores, "kind": "reference",
            "sample": "first %d instances of the batch, %d per core, reference TreeInNetContainment::displayed() with std::cout silenced; %d/%d displayed; per-core rate %.1f inst/s"
                      % (cores * per, per, disp, cores * per, sum(single) / len(single))}


def lca_leg(pt, torch, leaves, queries, peak, rank=0, world=1, dist=None):
    """BASELINE configs[2]; with N ranks the table is replicated (built on every GPU) and every rank answers its own `queries`
    (weak scaling, no data-path collective): aggregate = N * queries / max over ranks of the batch time"""
    par = pt.random_tree(leaves, SEED)
    n = len(par)
    dpar = torch.from_numpy(par).cuda()
    lca = pt.Lca(dpar)
    info = lca.info()
    g = torch.Generator(device="cuda").manual_seed(1 + rank)
    u = torch.randint(0, n, (queries,), generator=g, device="cuda", dtype=torch.int32)
    v = torch.randint(0, n, (queries,), generator=g, device="cuda", dtype=torch.int32)
    out = torch.empty_like(u)
    stream = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        lca.query(u, v, out=out, stream=stream)
    torch.cuda.synchronize()
    reps, best, tot = 5, 1e30, 0.0
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lca.query(u, v, out=out, stream=stream)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best, tot = min(best, ms), tot + ms
    avg = tot / reps
    checksum = int(out.long().sum().item())
    lca.free()
    if world > 1:
        t = torch.tensor([avg, best], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        avg, best = float(t[0].item()), float(t[1].item())
    qps = world * queries / (avg / 1000.0)
    ach = LCA_BYTES_PER_QUERY * qps / 1e9
    return {"workload": "batched LCA: %d queries per GPU on a %d-leaf random binary tree (BASELINE configs[2])" % (queries, leaves),
            "n_gpus": world, "scaling": "weak (table replicated, queries sharded, no collective)", "queries_per_s": qps, "ms_per_batch": avg, "best_ms": best, "build_ms": info["build_ms"], "levels": info["num_levels"],
            "resident_bytes": info["device_bytes"], "checksum": checksum,
            "roofline": {"bound": "hbm", "achieved": ach / world, "peak": peak, "unit": "GB/s", "frac": ach / world / peak, "traffic": measured_traffic("lca_query_kernel", queries),
                         "model": "40 B/query (SURVEY 8(d)); this layout needs 2 random records per query and B200 moves a 128-byte line per random read (profiles/micro_gather.cu): 254 B/query measured, ceiling ~25 G queries/s"}}


def config4_leg(pt, torch, count):
    """BASELINE configs[3]: large networks n = 10^6, r = 10^3 (global-memory executor, one CTA per instance)"""
    import numpy as np
    leaves, retis = 499_001, 1000
    t0 = time.perf_counter()
    p = pt.Packer()
    p.add_generated(count, leaves, retis, SEED + 4, 0)
    a = p.arrays()
    gen_s = time.perf_counter() - t0
    dev = {k: (torch.from_numpy(v).cuda() if isinstance(v, np.ndarray) else v) for k, v in a.items()}
    bits = torch.zeros((count + 31) // 32, dtype=torch.int32, device="cuda")
    stat = torch.zeros(count, dtype=torch.uint8, device="cuda")
    b = pt.Batch(dev)
    b.contain(result_bits=bits, status=stat, want_counters=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _, _, c = b.contain(result_bits=bits, status=stat, want_counters=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    b.free()
    nbytes = sum(int(dev[k].numel()) * dev[k].element_size() for k in ("host_succ_off", "host_succ_adj", "host_label", "guest_parent", "guest_label"))
    return {"workload": "%d gen networks n=%d r=%d with a displayed tree each (BASELINE configs[3], one GPU's share)" % (count, 2 * retis + 2 * leaves - 1, retis),
            "instances_per_s": count / (ms / 1000.0), "ms": ms, "displayed": c["displayed"], "errors": c["errors"], "work_items": c["work_items"],
            "rule_passes": c["rule_passes"], "input_bytes": nbytes, "host_generation_s": gen_s,
            "note": "one thread-block cluster per instance, workspace in global memory; the reference cannot run this size (extrapolated 3e5 s per instance, BASELINE.md)"}


def config5_leg(pt, retis=24, leaves=30, rank=0, world=1, dist=None, torch=None):
    """BASELINE configs[4]: exhaustive 2^r switching enumeration of one general network (guest = random tree: not displayed,
    so the whole index space is visited)"""
    p = pt.Packer()
    p.add_generated(1, leaves, retis, SEED + 5, 2)
    a = p.arrays()
    args = (a["host_succ_off"], a["host_succ_adj"], a["host_label"], a["guest_parent"], a["guest_label"])
    pt.enum_switchings(*args, begin=0, end=1 << 16)  # warm-up
    total = 1 << retis  # binary network: in-degree 2 at every reticulation
    lo, hi = total * rank // world, total * (rank + 1) // world  # contiguous sub-ranges of [0, 2^r), one reduction
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    tot, nd, first = pt.enum_switchings(*args, begin=lo, end=hi) if world > 1 else pt.enum_switchings(*args)
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([float(tot), float(nd)], dtype=torch.float64, device="cuda"); m = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM); dist.all_reduce(m, op=dist.ReduceOp.MAX)
        tot, nd, dt = int(tot), int(t[1].item()), float(m[0].item())  # n_switchings is the size of the whole index space on every rank
    return {"workload": "exhaustive switching enumeration, gen network l=%d r=%d, random guest tree (BASELINE configs[4])" % (leaves, retis),
            "n_gpus": world, "switchings": tot, "displaying": nd, "seconds": dt, "switchings_per_s": tot / dt, "host_nodes": int(a["host_node_off"][1]),
            "note": "wall time of pt_enum_switchings incl. upload; bound by integer ALU + L1-resident accumulators, not HBM"}


def config1_leg(pt):
    """BASELINE configs[0]: the reference's own data/test.nw (one instance; ground truth "not displayed", the reference crashes on it):
    latency of one pt_batch_create + pt_batch_contain + pt_batch_free from host arrays"""
    p = pt.Packer()
    p.add_newick("((a,(((b)#H1,(c)#H2),(#H1,#H2))),d);", "((a,b),(c,d));")
    a = p.arrays()
    ts = []
    for _ in range(20):
        t0 = time.perf_counter()
        b = pt.Batch(a)
        v, s, _ = b.contain(want_counters=False)
        b.free()
        ts.append(time.perf_counter() - t0)
    ts.sort()
    return {"workload": "data/test.nw of the reference, one instance", "displayed": bool(v[0]), "status": int(s[0]), "latency_us_median": ts[len(ts) // 2] * 1e6,
            "latency_us_min": ts[0] * 1e6, "note": "launch-latency bound: memsets, one kernel, one branching round, result copy"}


def who_leg(pt, torch, leaves=1_000_000):
    """TreeInTreeContainment::who_displays for every guest node, host = random binary tree with `leaves`
    leaves, guest = the same tree (every node is displayed, by itself); wall time of the call incl. upload and download"""
    import numpy as np
    par = pt.random_tree(leaves, SEED + 7)
    n = len(par)
    deg = np.bincount(par[par >= 0], minlength=n)
    lab = np.full(n, -1, dtype=np.int32)
    lab[deg == 0] = np.arange(int((deg == 0).sum()), dtype=np.int32)
    par, lab = torch.from_numpy(par).pin_memory().numpy(), torch.from_numpy(lab).pin_memory().numpy()
    pt.who_displays(par, lab, par, lab)  # warm-up
    t0 = time.perf_counter()
    out = pt.who_displays(par, lab, par, lab)
    dt = time.perf_counter() - t0
    # highest_displayed_ancestor of every guest node from that table (containment.hpp:328-344): everything is displayed here, so every
    # climb ends at the guest root -- the longest possible chains for the pointer jumping
    root = int(np.where(par < 0)[0][0])
    who = torch.from_numpy(out).pin_memory().numpy()
    pt.highest_displayed(par, who, root)
    t1 = time.perf_counter()
    hda = pt.highest_displayed(par, who, root)
    dt1 = time.perf_counter() - t1
    return {"workload": "who_displays of all %d guest nodes, host = guest = random binary tree with %d leaves" % (n, leaves),
            "guest_nodes_per_s": n / dt, "ms": dt * 1e3, "verdicts_ok": bool((out == np.arange(n)).all()),
            "highest_displayed_ancestor": {"guest_nodes_per_s": n / dt1, "ms": dt1 * 1e3, "ok": bool((hda == root).all())}}


def bridges_leg(pt, torch, leaves=499_001, retis=1000):
    """bridges and bridge-free components of one network of BASELINE config 4's size (n = 10^6); wall time of
    pt_net_bridges incl. upload and download"""
    p = pt.Packer()
    p.add_generated(1, leaves, retis, SEED + 8, 0)
    a = p.arrays()
    n, m = int(a["host_node_off"][1]), int(a["host_arc_off"][1])
    so, sa = torch.from_numpy(a["host_succ_off"][:n + 1].copy()).pin_memory().numpy(), torch.from_numpy(a["host_succ_adj"][:m].copy()).pin_memory().numpy()
    pt.net_bridges(so, sa)  # warm-up
    t0 = time.perf_counter()
    br, comp = pt.net_bridges(so, sa)
    dt = time.perf_counter() - t0
    return {"workload": "bridges + components of a gen network with n = %d, m = %d" % (n, m), "arcs_per_s": m / dt, "ms": dt * 1e3,
            "bridges": int(br.sum()), "components": int(len(set(comp.tolist())))}


def iso_leg(pt, torch, pairs=20000):
    """batched network isomorphism (pt_iso_batch) on pairs of config-2-sized networks: each network against an
    identical copy (isomorphic) and against the next one (not isomorphic); wall time of the call incl. the upload from pinned memory"""
    import numpy as np
    p = pt.Packer()
    p.add_generated(pairs + 1, L, R, SEED + 6, 0)
    a = p.arrays()
    n, m = int(a["host_node_off"][1]), int(a["host_arc_off"][1])
    so, sa = a["host_succ_off"].reshape(pairs + 1, n + 1), a["host_succ_adj"].reshape(pairs + 1, m)
    po, pa, lab = a["host_pred_off"].reshape(pairs + 1, n + 1), a["host_pred_adj"].reshape(pairs + 1, m), a["host_label"].reshape(pairs + 1, n)

    def arrays(idx):
        g = len(idx)
        c = lambda x: torch.from_numpy(np.ascontiguousarray(x[idx]).ravel()).pin_memory().numpy()   # pinned, like the e2e leg
        return dict(num_pairs=g // 2, node_off=np.arange(g + 1, dtype=np.int64) * n, arc_off=np.arange(g + 1, dtype=np.int64) * m,
                    succ_off=c(so), succ_adj=c(sa), pred_off=c(po), pred_adj=c(pa), label=c(lab), max_nodes=n)
    out = {"workload": "%d pairs of gen networks l=%d r=%d (n = %d), leaf labels have to match" % (pairs, L, R, n)}
    for name, arr, want in (("isomorphic", arrays(np.repeat(np.arange(pairs), 2)), pairs),
                            ("not_isomorphic", arrays(np.stack([np.arange(pairs), np.arange(pairs) + 1], axis=1).ravel()), 0)):
        pt.iso_batch(arr)  # warm-up
        t0 = time.perf_counter()
        v, st = pt.iso_batch(arr)
        dt = time.perf_counter() - t0
        out[name] = {"pairs_per_s": pairs / dt, "ms": dt * 1e3, "verdicts_ok": bool(int(v.sum()) == want and (st == 0).all())}
    # the same question asked from extended-Newick TEXT (pt_iso_from_newick: parsed and compared on the device)
    lines = [p.newick(i)[0] for i in range(pairs + 1)]
    text = ("\n".join(lines[i] + "\n" + lines[i + 1] for i in range(pairs)) + "\n").encode()
    pinned = torch.frombuffer(bytearray(text), dtype=torch.uint8).pin_memory()
    cap = len(text) // 4 + 2
    bits = torch.zeros((cap + 31) // 32 + 1, dtype=torch.int32).pin_memory()
    stat = torch.zeros(cap + 1, dtype=torch.uint8).pin_memory()
    found = ctypes.c_int64(0)
    best = 1e30
    for it in range(3):
        t0 = time.perf_counter()
        rc = pt.lib().pt_iso_from_newick(pinned.data_ptr(), len(text), 0, 1, bits.data_ptr(), stat.data_ptr(), 0, ctypes.byref(found), None)
        best = min(best, time.perf_counter() - t0)
        if rc != 0:
            raise RuntimeError("pt_iso_from_newick failed: %d" % rc)
    nz = int((bits.numpy().view("uint32") != 0).sum())
    out["not_isomorphic_from_text"] = {"pairs_per_s": pairs / best, "ms": best * 1e3, "text_mb": len(text) / 1e6,
                                       "verdicts_ok": bool(found.value == pairs and nz == 0 and int(stat[:pairs].sum()) == 0)}
    return out


def text_leg(pt, torch, pairs=100_000):
    """SURVEY 8(f) rank 1 on the device: extended-Newick TEXT of the whole workload (two lines per instance) -> verdicts.
    pt_batch_create_from_newick parses on the GPU (text from pinned host memory, upload included), pt_batch_contain solves;
    compare newick_ingest (the same text through the host reader)"""
    import numpy as np
    p = pt.Packer()
    p.add_generated(pairs, L, R, SEED, 0)
    text = ("\n".join(h + "\n" + g for h, g in (p.newick(i) for i in range(pairs))) + "\n").encode()
    pinned = torch.frombuffer(bytearray(text), dtype=torch.uint8).pin_memory()
    lib = pt.lib()
    words = (pairs + 31) // 32
    h_bits = torch.zeros(words + 1, dtype=torch.int32).pin_memory()
    h_stat = torch.zeros(pairs + 1, dtype=torch.uint8).pin_memory()
    opt = pt.Options(0, 0)
    best_parse = best_total = 1e30
    for it in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        h = ctypes.c_void_p()
        n = ctypes.c_int64(0)
        rc = lib.pt_batch_create_from_newick(ctypes.byref(h), pinned.data_ptr(), len(text), pt.PT_MEM_HOST, ctypes.byref(n), None)
        assert rc == 0, lib.pt_last_error()
        t1 = time.perf_counter()
        rc = lib.pt_batch_contain(h, ctypes.byref(opt), h_bits.data_ptr(), h_stat.data_ptr(), pt.PT_MEM_HOST, None, None)
        assert rc == 0, lib.pt_last_error()
        t2 = time.perf_counter()
        lib.pt_batch_free(h)
        if it:
            best_parse, best_total = min(best_parse, t1 - t0), min(best_total, t2 - t0)
    disp = int(np.unpackbits(h_bits.numpy().view(np.uint8), bitorder="little")[:pairs].sum())
    return {"workload": "extended-Newick text of %d instances (l=%d, r=%d), %d bytes, pinned host memory -> verdicts" % (pairs, L, R, len(text)),
            "parse_pairs_per_s": pairs / best_parse, "parse_ms": best_parse * 1e3, "parse_GB_per_s": len(text) / best_parse / 1e9,
            "text_to_verdict_pairs_per_s": pairs / best_total, "text_to_verdict_ms": best_total * 1e3, "displayed": disp}


def ingest_leg(pt, pairs=20000):
    """text -> flat arrays: the parallel batch reader on the first `pairs` instances of the workload"""
    p = pt.Packer()
    p.add_generated(pairs, L, R, SEED, 0)
    text = ("\n".join(h + "\n" + g for h, g in (p.newick(i) for i in range(pairs))) + "\n").encode()
    q = pt.Packer()
    t0 = time.perf_counter()
    k = q.add_newick_text(text)
    dt = time.perf_counter() - t0
    r = pt.Packer()
    t1 = time.perf_counter()
    r.add_newick_text(text, threads=1)
    dt1 = time.perf_counter() - t1
    return {"workload": "extended-Newick text of %d instances (l=%d, r=%d) -> flat arrays" % (k, L, R), "pairs_per_s": k / dt, "MB_per_s": len(text) / dt / 1e6,
            "threads": os.cpu_count(), "pairs_per_s_one_thread": k / dt1, "bytes": len(text)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--instances", type=int, default=100_000, help="instances per GPU")
    ap.add_argument("--lca-leaves", type=int, default=10_000_000)
    ap.add_argument("--lca-queries", type=int, default=100_000_000)
    ap.add_argument("--int32-input", action="store_true", help="use the int32 batch layout instead of the compact one")
    ap.add_argument("--layout", type=int, default=1, choices=[1, 2, 4], help="element size of the batch arrays (pt_batch_desc.index_bytes): 1 = uint8 with out-degrees (default), 2 = int16, 4 = int32")
    ap.add_argument("--no-lca", action="store_true")
    ap.add_argument("--no-config4", action="store_true")
    ap.add_argument("--no-config5", action="store_true")
    ap.add_argument("--config4-instances", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-instances-per-core", type=int, default=100)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3
    import numpy as np
    import torch
    import torch.distributed as dist

    import phylo_tools_b200 as pt
    from phylo_tools_b200.shard import place_words

    if pt.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU path")
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None  # before any pinned allocation (first touch)
    torch.cuda.set_device(local_rank)
    pt.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG", "WARN")  # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    peak, peak_src = measured_peak()
    B = args.instances
    total_B = B * world
    # ---- synthetic batch of this rank: instances [rank*B, (rank+1)*B) of the global batch (seed = SEED + global index) ----
    t_gen = time.perf_counter()
    packer = pt.Packer()
    packer.add_generated(B, L, R, SEED + rank * B, 0)
    layout = 4 if args.int32_input else args.layout
    host = packer.arrays(compact=layout if layout != 4 else 0)  # uint8 data arrays (pt_batch_desc.index_bytes = 1): a quarter of the int32 upload
    t_gen = time.perf_counter() - t_gen
    pinned = {k: (torch.from_numpy(v).pin_memory() if isinstance(v, np.ndarray) else v) for k, v in host.items()}
    dev = {k: (v.cuda(non_blocking=True) if hasattr(v, "cuda") else v) for k, v in pinned.items()}
    torch.cuda.synchronize()
    words = (B + 31) // 32
    d_bits = torch.zeros(words, dtype=torch.int32, device="cuda")
    d_stat = torch.zeros(B, dtype=torch.uint8, device="cuda")
    full_words = torch.zeros((total_B + 31) // 32, dtype=torch.int64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    resident = pt.Batch(dev, stream=stream)
    in_bytes_actual = sum(int(dev[k].numel()) * dev[k].element_size() for k in ("host_node_off", "host_arc_off", "guest_node_off", "host_succ_off", "host_succ_adj", "host_label", "guest_parent", "guest_label"))

    def step_resident():
        resident.contain(result_bits=d_bits, status=d_stat, want_counters=False, stream=stream)
        if world > 1:  # the path's only collective: packed result words (disjoint per rank -> SUM == OR)
            full_words.zero_()
            w0 = rank * B // 32
            full_words[w0:w0 + words] = d_bits.to(torch.int64) & 0xFFFFFFFF
            dist.all_reduce(full_words, op=dist.ReduceOp.SUM)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_resident()
    barrier()
    launches = 0
    k0_ms = []
    with ClockSampler(local_rank) as clocks:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(args.steps):
            step_resident()
            launches += resident.last_launches()
            k0_ms.append(resident.round_ms(0))
        e1.record()
        barrier()
        elapsed_ms = e0.elapsed_time(e1)
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / args.steps
    value = total_B / (ms_per_step / 1000.0)
    # verdict sanity (outside the timed region): every instance is a positive by construction
    n_disp = int(np.unpackbits(d_bits.cpu().numpy().view(np.uint8), bitorder="little")[:B].sum())
    n_err = int((d_stat != 0).sum().item())
    _, _, counters = resident.contain(result_bits=d_bits, status=d_stat, want_counters=True, stream=stream)
    k0 = sum(k0_ms) / len(k0_ms)
    ach = ALG_BYTES_PER_INSTANCE * B / (k0 / 1000.0) / 1e9
    roofline = {"bound": "hbm", "kernel": "tc_persistent_kernel (the one launch of a step, %d instances)" % B, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": measured_traffic("tc_reduce_kernel", B), "kernel_ms": k0, "kernel_share_of_step": k0 / ms_per_step, "peak_source": peak_src,
                "algorithmic_bytes_per_instance": ALG_BYTES_PER_INSTANCE, "bytes_actually_read_per_instance": in_bytes_actual / B,
                "note": "instance-resident shared-memory solver: bound by dependent shared-memory latency and barriers, not HBM (SURVEY 8(d)); see profiles/"}
    resident.free()

    # ---- end to end: HOST (pinned) buffers -> pt_batch_create (H2D) -> pt_batch_contain (HOST result bits, D2H) -> free ----
    h_bits = torch.zeros(max(words, 1), dtype=torch.int32).pin_memory()
    h_stat = torch.zeros(max(B, 1), dtype=torch.uint8).pin_memory()
    pinned_np = {k: (v.numpy() if hasattr(v, "numpy") else v) for k, v in pinned.items()}
    h2d = 0

    def step_e2e():
        nonlocal h2d
        b = pt.Batch(pinned_np, stream=stream)
        h2d = b.h2d_bytes()
        opt = pt.Options(0, 0)
        rc = pt.lib().pt_batch_contain(b._h, ctypes.byref(opt), h_bits.data_ptr(), h_stat.data_ptr(), pt.PT_MEM_HOST, None, stream)
        assert rc == 0, pt.lib().pt_last_error()
        b.free()
        if world > 1:
            fw = torch.from_numpy(place_words(h_bits.numpy().view(np.uint32)[:words], rank * B // 32 * 32, total_B)).cuda()
            dist.all_reduce(fw, op=dist.ReduceOp.SUM)

    for _ in range(args.warmup):
        step_e2e()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        step_e2e()
    e1.record()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1000.0
    e2e_ms = max(e0.elapsed_time(e1), wall_ms)  # host syncs inside the call: take the larger of device and wall time
    t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = total_B / (float(t.item()) / args.steps / 1000.0)
    e2e_disp = int(np.unpackbits(h_bits.numpy().view(np.uint8), bitorder="little")[:B].sum())

    out = {"metric": "tree-containment instances/s", "value": value, "unit": "instances/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16", "data": "synthetic",
           "config": {"workload": "batch of 10^5 gen networks l=100 r=20 with a displayed tree each, per GPU (BASELINE configs[1])",
                      "instances_per_gpu": B, "leaves": L, "reticulations": R, "host_nodes": 2 * R + 2 * L - 1, "generator_seed": SEED,
                      "input_layout": {4: "int32 arrays", 2: "int16 arrays (pt_batch_desc.index_bytes = 2)", 1: "uint8 arrays, out-degrees instead of offsets (pt_batch_desc.index_bytes = 1)"}[layout],
                      "l2": "inputs (%.0f MB per step) exceed the 126 MB L2; no explicit flush" % (in_bytes_actual / 1e6),
                      "parallelism": "instances sharded %d-way, one all-reduce of result words" % world if world > 1 else "single GPU",
                      "cpu_affinity": ("%d cores next to the GPU" % numa) if numa else "default"},
           "e2e": {"value": e2e_val, "unit": "instances/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": words * 4 + B,
                   "api": "pt_batch_create(HOST pinned) + pt_batch_contain(HOST outputs) + pt_batch_free", "displayed": e2e_disp},
           "gpu_launches": launches, "clocks": clocks.summary(), "roofline": roofline,
           "verdicts": {"displayed": n_disp, "errors": n_err, "expected_displayed": B}, "counters": counters,
           "host_generation_s": t_gen}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline_leg(pt)
    if not args.no_lca:  # every rank: queries are sharded
        try:
            out["lca"] = lca_leg(pt, torch, args.lca_leaves, args.lca_queries, peak, rank, world, dist if world > 1 else None)
        except Exception as ex:  # the secondary leg must not take the headline line down
            if world > 1:
                raise
            out["lca"] = {"error": repr(ex)}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            out["newick_ingest"] = ingest_leg(pt)
        except Exception as ex:
            out["newick_ingest"] = {"error": repr(ex)}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            out["iso"] = iso_leg(pt, torch)
        except Exception as ex:
            out["iso"] = {"error": repr(ex)}
        try:
            out["newick_on_gpu"] = text_leg(pt, torch)
        except Exception as ex:
            out["newick_on_gpu"] = {"error": repr(ex)}
        try:
            out["config1"] = config1_leg(pt)
        except Exception as ex:
            out["config1"] = {"error": repr(ex)}
        for key, leg in (("who_displays", who_leg), ("bridges", bridges_leg)):
            try:
                out[key] = leg(pt, torch)
            except Exception as ex:
                out[key] = {"error": repr(ex)}
    if rank == 0 and not args.no_config4:
        try:
            out["config4"] = config4_leg(pt, torch, args.config4_instances)
        except Exception as ex:
            out["config4"] = {"error": repr(ex)}
    if not args.no_config5:  # every rank: the index range is sharded
        try:
            out["config5"] = config5_leg(pt, rank=rank, world=world, dist=dist if world > 1 else None, torch=torch)
        except Exception as ex:
            if world > 1:
                raise
            out["config5"] = {"error": repr(ex)}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))


if __name__ == "__main__":
    main()
