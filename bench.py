#!/usr/bin/env python
"""bench.py -- the headline benchmark of BASELINE.json on B200.

metric   : tree-containment instances/s  (BASELINE.json "metric"; config = configs[1]: ONE batch of 10^5 `gen` networks
           l=100, r=20, each with a displayed tree)
step     : one pass of the hot path over the whole batch.  With N GPUs the batch is cut into N contiguous ranges (STRONG scaling:
           12 500 instances per GPU at N = 8, as SURVEY 8(e) and the north star name it); every rank solves its range with one
           persistent kernel launch and the library's own collective (pt_batch_contain_sharded: one NCCL all-reduce of the packed
           result words) leaves the whole verdict vector on every rank.
value    : whole-job instances/s with the batch RESIDENT in HBM when the timed region starts (CUDA events, max over ranks)
e2e      : the same through the public C-ABI with HOST (pinned) buffers: pt_batch_create (H2D, uint8 layout) +
           pt_batch_contain_sharded (kernel + collective) + D2H of the result words + host sync + pt_batch_free, every step;
           "value" = double-buffered (step k+1's batch is created while step k's kernel runs), "serial_value" = one step after the other
roofline : dominant kernel = tc_persistent_kernel (the one launch of a step), timed live with CUDA events inside the library;
           algorithmic bytes = 8124 B/instance; peak = MEASURED_PEAKS.json hbm_gbs
cpu_baseline / --impl reference : the reference's own TreeInNetContainment::displayed() (oracle/_ref/ref_tc, unmodified
           reference headers, std::cout silenced) on the box's host cores, one process per core over disjoint slices
extras   : weak scaling (10^5 instances per GPU), the LCA legs (BASELINE config 3: uniform pairs, and the consecutive
           preorder-sorted candidates the path really issues), config 4 (64 large instances, 8 per rank at N = 8), config 5
           (switching enumeration with the cross-GPU found flag), text / isomorphism / who_displays / bridges legs at N = 1
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
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/traffic_r02.json), only when the
    launch here processes the same number of units as the captured one; else None"""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic_r02.json")))[kernel]
        return t["dram_bytes_per_launch"] if t["units_per_launch"] == units else None
    except Exception:
        return None


def measured_ncu(kernel):
    """what the committed ncu --set full capture of `kernel` says beside the bytes: L2 hit rate, warp execution efficiency (active threads
    per instruction), achieved occupancy, issue-slot utilisation, top stall reasons (profiles/traffic_r02.json; static, from the capture)"""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic_r02.json")))[kernel].get("ncu")
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
        self.index, self.rows, self.stop, self.t, self.sample = index, [], False, None, None

    def _nvml_open(self):
        """NVML in-process: a sample every 1 ms (the timed region is 13 ms at N = 8, 80 ms at N = 1: too short for nvidia-smi's ~100 ms per call).
        Opened BEFORE the timed region starts, so that the first sample falls inside it."""
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
        get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}   # nvmlClocksEventReason*

        def sample():
            sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            r = get_reasons(h)
            self.rows.append([str(sm), str(mx)] + [("Active" if r & bits[n] else "Not Active") for n in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")])
        sample()          # (fails here, not in the thread, if NVML cannot be used)
        self.rows.clear()
        return sample

    def _run_nvml(self):
        while not self.stop:
            self.sample()
            time.sleep(0.001)

    def _run(self):
        if self.sample is not None:
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
        try:
            self.sample = self._nvml_open()
        except Exception:
            self.sample = None
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()
        return self

    def __exit__(self, *a):
        if self.sample is not None:
            try:
                self.sample()     # one more while the last timed step is still in flight / just over
            except Exception:
                pass
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


def _gen_pair_files(count_per_file, nfiles, seed):
    """`nfiles` files of `count_per_file` (network, tree) pairs each, made by the stand-alone generator phylo_tools_b200/bin/gen
    (host code only; it does not load libptgpu.so) -- the same instances pt_packer_add_generated(count, L, R, seed, 0) makes"""
    gen = os.path.join(ROOT, "phylo_tools_b200", "bin", "gen")
    files = []
    for k in range(nfiles):
        f = tempfile.NamedTemporaryFile("w", suffix=".nw", delete=False)
        f.close()
        subprocess.check_call([gen, "--pairs", str(count_per_file), "-l", str(L), "-r", str(R), "--seed", str(seed + k * count_per_file), f.name],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        files.append(f.name)
    return files


def reference_arm(args, rank, world):
    """the reference's CPU implementation of the path on the host cores (rank 0 only); the product library is not imported"""
    if rank != 0:
        return
    from oracle import pt_oracle as O
    ref = O.ref_binary("ref_tc")
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 64))
    per = args.ref_instances_per_core
    res = {"metric": "tree-containment instances/s", "unit": "instances/s", "impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
           "config": {"workload": "batch of 10^5 gen networks l=100 r=20 with a displayed tree each (BASELINE configs[1]); bounded sample per step",
                      "leaves": L, "reticulations": R}}
    if ref is None:
        res["unavailable"] = "oracle/_ref/ref_tc was not built (needs /root/reference at build time)"
        print(json.dumps(res))
        return
    files = _gen_pair_files(per, procs, SEED)

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
                                 "sample": "%d instances per step (the first %d of the batch, %d per core); reference gave a verdict on %d (%d displayed), crashed/threw on %d" % (total, total, per, done, disp, total - done)},
                "e2e": {"value": val, "unit": "instances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(res))


def cpu_baseline_leg(sample_seconds=15.0):
    """bounded sample of the same workload through the reference on all host cores (reported, not the target), plus the one-core
    rate of the same driver"""
    from oracle import pt_oracle as O
    ref = O.ref_binary("ref_tc")
    if ref is None:
        return {"value": None, "unit": "instances/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref/ref_tc not built"}
    cores = max(1, min(os.cpu_count() or 1, 64))
    per = max(8, int(sample_seconds * 45))  # ~50 instances/s/core survey-time
    per = min(per, 400)
    files = _gen_pair_files(per, cores, SEED)
    t0 = time.perf_counter()
    ps = [subprocess.Popen([ref, "bench", fn, str(per)], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True) for fn in files]
    outs = [q.communicate()[1] for q in ps]
    dt = time.perf_counter() - t0
    js = [_ref_json(o) for o in outs]
    disp = sum(j.get("displayed", 0) for j in js)
    done = sum(j.get("displayed", 0) + j.get("not_displayed", 0) for j in js)
    crashed = cores * per - done
    # one core alone (no neighbours competing for the memory system): the first 100 instances
    one = _ref_json(subprocess.run([ref, "bench", files[0], "100"], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True).stderr)
    # the reference's own `tc` tool as it is (examples/tc.cpp with its unconditional printing, one process per instance): SURVEY 8(d)
    stock, stock_rate = O.ref_binary("tc_stock"), None
    if stock is not None:
        lines = open(files[0]).read().split("\n")
        n_stock, verdicts = 40, 0
        tmp = tempfile.NamedTemporaryFile("w", suffix=".nw", delete=False)
        tmp.close()
        t1 = time.perf_counter()
        for i in range(n_stock):
            with open(tmp.name, "w") as f:
                f.write(lines[2 * i] + "\n" + lines[2 * i + 1] + "\n")
            r = subprocess.run([stock, tmp.name], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            tail = r.stdout.rstrip().rsplit("\n", 1)[-1] if r.stdout else ""
            verdicts += tail in ("displayed", "not displayed")
        stock_rate = verdicts / (time.perf_counter() - t1) if verdicts else None
        os.unlink(tmp.name)
    for fn in files:
        os.unlink(fn)
    return {"value": done / dt, "reference_crashes_or_exceptions": crashed, "unit": "instances/s", "cores": cores, "kind": "reference",
            "one_core_instances_per_s": one.get("inst_per_s"), "stock_tc_one_core_instances_per_s": stock_rate,
            "sample": "first %d instances of the batch, %d per core, reference TreeInNetContainment::displayed() with std::cout silenced; %d/%d displayed"
                      % (cores * per, per, disp, cores * per)}


def lca_cpu_baseline(leaves=1_000_000, queries=2_000_000):
    """the reference's own lca (rmq/lca.hpp: Euler tour + pm_rmq) through oracle/_ref/ref_lca on one host core; the reference's
    recursive tree<T> does not reach 10^7 leaves comfortably, so the sample is a 10^6-leaf tree"""
    from oracle import pt_oracle as O
    ref = O.ref_binary("ref_lca")
    if ref is None:
        return {"value": None, "kind": "reference", "sample": "oracle/_ref/ref_lca not built"}
    try:
        out = subprocess.run([ref, "bench_lca", str(leaves), str(queries), str(SEED)], capture_output=True, text=True, timeout=300).stdout.strip()
        j = json.loads(out)
        return {"value": j["queries"] / j["query_s"], "unit": "queries/s", "cores": 1, "kind": "reference", "build_s": j["build_s"],
                "sample": "%d uniform queries on a %d-leaf random binary tree, reference lca<>::query (rmq/lca.hpp:91-106)" % (queries, leaves)}
    except Exception as ex:
        return {"value": None, "kind": "reference", "sample": "ref_lca failed: %r" % (ex,)}


def lca_leg(pt, torch, leaves, queries, peak, rank=0, world=1, dist=None):
    """BASELINE configs[2]; with N ranks the table is replicated (built on every GPU) and every rank answers its own `queries`
    (weak scaling, no data-path collective): aggregate = N * queries / max over ranks of the batch time.  Two workloads:
    uniform pairs (the config as SURVEY 8(d) defines it) and the one the containment path really issues -- consecutive
    candidates of a preorder-sorted list (utils/induced_tree.hpp:128-138) on a tree numbered in preorder like the reference's"""
    import numpy as np
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

    def timed(fn, reps=5, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        best, tot = 1e30, 0.0
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best, tot = min(best, ms), tot + ms
        avg = tot / reps
        if world > 1:
            t = torch.tensor([avg, best], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            avg, best = float(t[0].item()), float(t[1].item())
        return avg, best
    avg, best = timed(lambda: lca.query(u, v, out=out, stream=stream))
    checksum = int(out.long().sum().item())
    qps = world * queries / (avg / 1000.0)
    ach = LCA_BYTES_PER_QUERY * qps / 1e9
    res = {"workload": "batched LCA: %d uniform queries per GPU on a %d-leaf random binary tree (BASELINE configs[2])" % (queries, leaves),
           "n_gpus": world, "scaling": "weak (table replicated, queries sharded, no collective)", "queries_per_s": qps, "ms_per_batch": avg, "best_ms": best, "build_ms": info["build_ms"], "levels": info["num_levels"],
           "resident_bytes": info["device_bytes"], "checksum": checksum,
           "roofline": {"bound": "hbm", "achieved": ach / world, "peak": peak, "unit": "GB/s", "frac": ach / world / peak, "traffic": measured_traffic("lca_query_kernel", queries),
                        "model": "40 B/query (SURVEY 8(d)); uniform pairs need 2 random records per query and B200 moves a 128-byte line per random read (profiles/micro_gather.cu): 254 B/query measured, "
                                 "i.e. the kernel moves 5.3 TB/s of DRAM lines = 0.81 of the measured peak; bound for uniform pairs with this layout = peak / (2 x 128-byte lines + 12 B streamed) = 24.4 G queries/s (DESIGN.md section 3)"}}
    if rank == 0:
        # end to end: queries in pinned host memory in, answers in pinned host memory out (H2D 8 B + D2H 4 B per query inside the call)
        q2 = min(queries, 20_000_000)
        hu, hv = u[:q2].cpu().pin_memory(), v[:q2].cpu().pin_memory()
        ho = torch.empty(q2, dtype=torch.int32).pin_memory()
        lib = pt.lib()
        for _ in range(2):
            lib.pt_lca_query(lca._h, hu.data_ptr(), hv.data_ptr(), ho.data_ptr(), q2, pt.PT_MEM_HOST, stream)
        t0 = time.perf_counter()
        for _ in range(3):
            lib.pt_lca_query(lca._h, hu.data_ptr(), hv.data_ptr(), ho.data_ptr(), q2, pt.PT_MEM_HOST, stream)
        dt = (time.perf_counter() - t0) / 3
        res["e2e"] = {"value": q2 / dt, "unit": "queries/s", "h2d_bytes_per_step": 8 * q2, "d2h_bytes_per_step": 4 * q2, "queries": q2,
                      "answers_ok": bool((ho[:1000].numpy() == out[:1000].cpu().numpy()).all()), "api": "pt_lca_query(HOST pinned u, v -> HOST pinned out)"}
    lca.free()
    # ---- the consumer's workload: tree numbered in preorder, candidates sorted by preorder, consecutive pairs ----
    l0 = pt.Lca(dpar)
    pre = torch.empty(n, dtype=torch.int32, device="cuda")
    pt._check(pt.lib().pt_lca_node_info(l0._h, None, pre.data_ptr(), pt.PT_MEM_DEVICE, None), "pt_lca_node_info")
    l0.free()
    newpar = torch.empty_like(dpar)
    newpar[pre.long()] = torch.where(dpar >= 0, pre[dpar.clamp(min=0).long()], torch.full_like(dpar, -1))
    del dpar
    lca = pt.Lca(newpar)
    pos = torch.empty(n, dtype=torch.int32, device="cuda")
    pt._check(pt.lib().pt_lca_node_info(lca._h, None, pos.data_ptr(), pt.PT_MEM_DEVICE, None), "pt_lca_node_info")
    res["adjacent"] = {"workload": "consecutive pairs of a preorder-sorted candidate list (utils/induced_tree.hpp:128-138), tree ids = a preorder like the reference's readers assign; "
                                   "pt_lca_query_adjacent: every record read once"}
    for dens in (1.0, 0.25):
        keep = torch.rand(n, generator=g, device="cuda") < dens
        cand = torch.nonzero(keep).squeeze(1).to(torch.int32)
        cand = cand[torch.argsort(pos[cand.long()])].contiguous()
        o2 = torch.empty(cand.numel() - 1, dtype=torch.int32, device="cuda")
        a2, b2 = timed(lambda: lca.query_adjacent(cand, out=o2, stream=stream))
        qn = cand.numel() - 1
        qps2 = world * qn / (a2 / 1000.0)
        ok = bool((o2[:2000].cpu().numpy() == lca.query(cand[:2000].cpu().numpy(), cand[1:2001].cpu().numpy())).all())
        res["adjacent"]["density_%g" % dens] = {"queries": qn, "ms": a2, "queries_per_s": qps2, "answers_ok": ok,
                                                "roofline": {"bound": "hbm", "achieved": LCA_BYTES_PER_QUERY * qps2 / 1e9 / world, "peak": peak, "unit": "GB/s",
                                                             "frac": LCA_BYTES_PER_QUERY * qps2 / 1e9 / world / peak, "traffic": measured_traffic("lca_adjacent_kernel", qn),
                                                             "model": "40 B/query: 4 B id in, one 32-byte record, 4 B out"}}
    lca.free()
    if rank == 0 and world == 1:
        res["cpu_baseline"] = lca_cpu_baseline()
    return res


def config4_leg(pt, torch, count, rank=0, world=1, dist=None):
    """BASELINE configs[3]: large networks n = 10^6, r = 10^3, 64 instances sharded over 8 GPUs = 8 per rank; every rank solves
    its own 8 (global-memory executor, one thread-block cluster per instance); aggregate = world * count / max over ranks"""
    import numpy as np
    leaves, retis = 499_001, 1000
    t0 = time.perf_counter()
    p = pt.Packer()
    p.add_generated(count, leaves, retis, SEED + 4 + 1000 * rank, 0)
    a = p.arrays()
    gen_s = time.perf_counter() - t0
    dev = {k: (torch.from_numpy(v).cuda() if isinstance(v, np.ndarray) else v) for k, v in a.items()}
    bits = torch.zeros((count + 31) // 32, dtype=torch.int32, device="cuda")
    stat = torch.zeros(count, dtype=torch.uint8, device="cuda")
    b = pt.Batch(dev)
    b.contain(result_bits=bits, status=stat, want_counters=False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _, _, c = b.contain(result_bits=bits, status=stat, want_counters=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    b.free()
    disp, errs = c["displayed"], c["errors"]
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        t = torch.tensor([disp, errs], dtype=torch.int64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        disp, errs = int(t[0].item()), int(t[1].item())
    nbytes = sum(int(dev[k].numel()) * dev[k].element_size() for k in ("host_succ_off", "host_succ_adj", "host_label", "guest_parent", "guest_label"))
    return {"workload": "%d gen networks n=%d r=%d with a displayed tree each, %d per GPU on %d GPU(s) (BASELINE configs[3]: 64 over 8)" % (count * world, 2 * retis + 2 * leaves - 1, retis, count, world),
            "n_gpus": world, "instances_per_s": world * count / (ms / 1000.0), "ms": ms, "displayed": disp, "errors": errs, "work_items": c["work_items"],
            "rule_passes": c["rule_passes"], "input_bytes_per_gpu": nbytes, "host_generation_s": gen_s,
            "note": "one thread-block cluster per instance, workspace in global memory; the reference cannot run this size (extrapolated 3e5 s per instance, BASELINE.md)"}


def config5_leg(pt, retis=24, leaves=30, rank=0, world=1, dist=None, torch=None, comm=None):
    """BASELINE configs[4]: exhaustive 2^r switching enumeration of one general network, the index space cut into one contiguous
    range per GPU.  Persistent handle (pt_enum_create once; a run = kernel launches only).  Two runs:
    a random guest (not displayed: the whole space is visited), and a displayed guest with stop_at_first, where the found word
    is all-reduced after every chunk and every GPU stops as soon as one has a witness."""
    def make(kind, seed):
        p = pt.Packer()
        p.add_generated(1, leaves, retis, seed, kind)
        a = p.arrays()
        return pt.Enum(a["host_succ_off"], a["host_succ_adj"], a["host_label"], a["guest_parent"], a["guest_label"]), int(a["host_node_off"][1])
    E, hn = make(2, SEED + 5)
    E.run(begin=0, end=1 << 16)  # warm-up
    total = E.total
    lo, hi = total * rank // world, total * (rank + 1) // world

    def timed(fn):
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        r = fn()
        dt = time.perf_counter() - t0
        if world > 1:
            m = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(m, op=dist.ReduceOp.MAX)
            dt = float(m[0].item())
        return r, dt
    (nd, first, _), dt = timed(lambda: E.run(begin=lo, end=hi, comm=comm, chunk=1 << 21))
    if world > 1:
        t = torch.tensor([nd], dtype=torch.int64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        nd = int(t[0].item())
    E.free()
    out = {"workload": "exhaustive switching enumeration, gen network l=%d r=%d, random guest tree (BASELINE configs[4]); index space sharded over %d GPU(s)" % (leaves, retis, world),
           "n_gpus": world, "switchings": total, "displaying": nd, "seconds": dt, "switchings_per_s": total / dt, "host_nodes": hn,
           "note": "wall time of pt_enum_run on a persistent handle (no allocation / upload per run); bound by integer ALU + L1-resident accumulators, not HBM"}
    # early exit: a displayed guest; every rank stops once any rank has a witness
    E2, _ = make(0, SEED + 6)
    (nd2, first2, early), dt2 = timed(lambda: E2.run(begin=lo, end=hi, stop_at_first=True, comm=comm, chunk=1 << 18))
    E2.free()
    flags = [1 if early else 0, 1 if first2 is not None else 0]
    if world > 1:
        t = torch.tensor(flags, dtype=torch.int64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        flags = [int(t[0].item()), int(t[1].item())]
    out["early_exit"] = {"workload": "the same network with a DISPLAYED guest, stop_at_first, found word all-reduced (MAX) across the GPUs every 2^18 switchings",
                         "seconds": dt2, "ranks_with_a_witness": flags[1], "ranks_stopped_early": flags[0], "fraction_of_full_enumeration_time": dt2 / dt}
    return out


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
    # the same door double-buffered, as a caller with a stream of text batches would use it: batch k+1 is created on a second stream
    # (its 121 MB upload runs on the copy engine while batch k's persistent kernel holds the SMs; its parse kernels follow that
    # kernel), its solve and result read are queued, then the host waits for batch k's verdict words.  Two pinned text buffers.
    pinned2 = pinned.clone().pin_memory()
    texts = [pinned, pinned2]
    s_solve, s_parse = torch.cuda.Stream(), torch.cuda.Stream()
    d_bits = [torch.zeros(words + 1, dtype=torch.int32, device="cuda") for _ in range(2)]
    hb = [torch.zeros(words + 1, dtype=torch.int32).pin_memory() for _ in range(2)]

    def issue(k):
        h = ctypes.c_void_p()
        n = ctypes.c_int64(0)
        rc = lib.pt_batch_create_from_newick(ctypes.byref(h), texts[k % 2].data_ptr(), len(text), pt.PT_MEM_HOST, ctypes.byref(n), ctypes.c_void_p(s_parse.cuda_stream))
        assert rc == 0, lib.pt_last_error()
        rc = lib.pt_batch_contain(h, ctypes.byref(opt), d_bits[k % 2].data_ptr(), None, pt.PT_MEM_DEVICE, None, ctypes.c_void_p(s_solve.cuda_stream))
        assert rc == 0, lib.pt_last_error()
        with torch.cuda.stream(s_solve):
            hb[k % 2].copy_(d_bits[k % 2], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(s_solve)
        return h, ev, k

    def retire(x):
        h, ev, k = x
        ev.synchronize()
        lib.pt_batch_free(h)
        return hb[k % 2]
    torch.cuda.synchronize()
    inflight = issue(0)
    for k in range(1, 3):           # warm-up
        nxt = issue(k); retire(inflight); inflight = nxt
    K = 6
    t0 = time.perf_counter()
    for k in range(3, 3 + K):       # K creates (K uploads, K parses), K solves, K result reads inside the timed region
        nxt = issue(k); last = retire(inflight); inflight = nxt
    piped = (time.perf_counter() - t0) / K
    last = retire(inflight)
    torch.cuda.synchronize()
    disp2 = int(np.unpackbits(last.numpy().view(np.uint8), bitorder="little")[:pairs].sum())
    return {"workload": "extended-Newick text of %d instances (l=%d, r=%d), %d bytes, pinned host memory -> verdicts" % (pairs, L, R, len(text)),
            "parse_pairs_per_s": pairs / best_parse, "parse_ms": best_parse * 1e3, "parse_GB_per_s": len(text) / best_parse / 1e9,
            "text_to_verdict_pairs_per_s": pairs / best_total, "text_to_verdict_ms": best_total * 1e3, "displayed": disp,
            "double_buffered": {"text_to_verdict_pairs_per_s": pairs / piped, "ms_per_batch": piped * 1e3, "displayed": disp2,
                                "mode": "batch k+1 created (upload + GPU parse, second stream) while batch k is solved; wall clock over %d batches, every batch uploaded, parsed, solved and its verdict words read back" % K}}


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
    ap.add_argument("--instances", type=int, default=100_000, help="instances of the WHOLE batch (sharded over the GPUs)")
    ap.add_argument("--lca-leaves", type=int, default=10_000_000)
    ap.add_argument("--lca-queries", type=int, default=100_000_000)
    ap.add_argument("--layout", type=int, default=1, choices=[1, 2, 4], help="element size of the batch arrays (pt_batch_desc.index_bytes): 1 = uint8 with out-degrees (default), 2 = int16, 4 = int32")
    ap.add_argument("--no-lca", action="store_true")
    ap.add_argument("--no-weak", action="store_true")
    ap.add_argument("--no-config4", action="store_true")
    ap.add_argument("--no-config5", action="store_true")
    ap.add_argument("--config4-instances", type=int, default=8, help="per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-instances-per-core", type=int, default=100)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return
    # stdout carries exactly ONE line, the JSON: whatever a library prints to file descriptor 1 from here on (NCCL's version banner
    # comes from ncclCommInitRank whatever NCCL_DEBUG_FILE says) is sent to stderr; the JSON line goes to the saved descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if args.warmup < 3:
        args.warmup = 3
    import numpy as np
    import torch
    import torch.distributed as dist

    import phylo_tools_b200 as pt

    if pt.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU path")
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None  # before any pinned allocation (first touch)
    torch.cuda.set_device(local_rank)
    pt.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG", "WARN")              # no version banner: rank 0 prints ONE JSON line on stdout
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # and whatever a caller's NCCL_DEBUG asks for goes to stderr
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def bootstrap(ident):  # the library's NCCL id travels by torch.distributed (the host program's own bootstrap)
        box = [ident]
        dist.broadcast_object_list(box, src=0)
        return box[0]
    comm = pt.Comm(world, rank, bootstrap if world > 1 else None)
    peak, peak_src = measured_peak()
    layout = args.layout
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.int64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return int(t.item())

    def make_shard(total, seed):
        """this rank's contiguous range of a batch of `total` generated instances (boundaries multiples of 32): instance i of
        the batch is pt_packer_add_generated(..., seed + i, kind 0) whichever rank makes it"""
        cuts = [min(total, (total * r // world + 31) // 32 * 32) for r in range(world)] + [total]
        first, count = cuts[rank], cuts[rank + 1] - cuts[rank]
        if count == 0:
            first = first // 32 * 32      # an empty trailing shard of a tiny batch: any word boundary will do
        packer = pt.Packer()
        packer.add_generated(count, L, R, seed + first, 0)
        host = packer.arrays(compact=layout if layout != 4 else 0)
        pinned = {k: (torch.from_numpy(v).pin_memory() if isinstance(v, np.ndarray) else v) for k, v in host.items()}
        return first, count, pinned

    L2_ROTATE_BYTES = 256 << 20   # twice the 126 MB L2: a resident copy is read again only after this many other input bytes went through

    def run_resident(total, first, count, pinned, steps, warmup, clocks_for=None):
        """K timed steps over this rank's RESIDENT shard.  Timing rule "inputs larger than L2": the shard is kept in R device copies
        (R * shard bytes >= 2 x L2) and step i reads copy i mod R, so no step finds its input in L2 from an earlier step."""
        dev = {k: (v.cuda(non_blocking=True) if hasattr(v, "cuda") else v) for k, v in pinned.items()}
        torch.cuda.synchronize()
        in_keys = ("host_node_off", "host_arc_off", "guest_node_off", "host_succ_off", "host_succ_adj", "host_label", "guest_parent", "guest_label")
        in_bytes = sum(int(dev[k].numel()) * dev[k].element_size() for k in in_keys)
        copies = max(2, min(32, -(-L2_ROTATE_BYTES // max(in_bytes, 1))))
        devs = [dev] + [{k: (v.clone() if hasattr(v, "clone") else v) for k, v in dev.items()} for _ in range(copies - 1)]
        torch.cuda.synchronize()
        words = (total + 31) // 32
        full_bits = torch.zeros(words, dtype=torch.int32, device="cuda")
        full_stat = torch.zeros((total + 3) // 4 * 4, dtype=torch.uint8, device="cuda")
        residents = [pt.Batch(d, stream=stream) for d in devs]
        turn = [0]

        def step():
            b = residents[turn[0] % copies]
            turn[0] += 1
            comm.contain_sharded(b, first, total, full_bits, full_stat, stream=stream)
            return b
        for _ in range(max(warmup, copies)):      # every copy's handle has run once (allocator, function attributes) before the clock starts
            step()
        barrier()
        launches = 0
        timed = []                  # (handle, index of its first launch of the step, one past its last)
        ctx = ClockSampler(local_rank) if clocks_for is not None else None
        if ctx:
            ctx.__enter__()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):      # the steps are queued back to back: the host does not wait inside the timed region
            b = step()
            n_l = b.launch_count()
            timed.append((b, n_l - b.last_launches(), n_l))
            launches += b.last_launches() + (1 if comm.peer_exchange() and world > 1 else 0)   # the persistent kernel (+ the library's exchange kernel)
        e1.record()
        barrier()
        k0 = [sum(b.launch_ms(i) for i in range(lo, hi)) for (b, lo, hi) in timed if b.launch_count() - lo <= 64]   # every timed step's kernel, from the handles' event rings
        if ctx:
            ctx.__exit__(None, None, None)
            clocks_for.update(ctx.summary())
        ms = max_over_ranks(e0.elapsed_time(e1)) / steps
        counters = comm.contain_sharded(residents[0], first, total, full_bits, full_stat, want_counters=True, stream=stream)
        n_disp = int(np.unpackbits(full_bits.cpu().numpy().view(np.uint8), bitorder="little")[:total].sum())
        n_err = int((full_stat[:total] != 0).sum().item())
        for b in residents:
            b.free()
        k_avg = sum(k0) / len(k0)
        k_all = [k_avg]
        if world > 1:      # every rank's own kernel time (the step follows the slowest rank)
            k_all = [None] * world
            dist.all_gather_object(k_all, (k_avg, max(k0)))
        return dict(ms=ms, k0=k_avg, k_all=k_all, launches=launches, counters=counters, displayed=n_disp, errors=n_err, in_bytes=in_bytes, copies=copies)

    # ---- headline: ONE batch of `args.instances`, sharded over the ranks (strong scaling) ----
    total_B = args.instances
    t_gen = time.perf_counter()
    first, B, pinned = make_shard(total_B, SEED)
    t_gen = time.perf_counter() - t_gen
    clocks = {}
    r = run_resident(total_B, first, B, pinned, args.steps, args.warmup, clocks_for=clocks)
    ms_per_step = r["ms"]
    value = total_B / (ms_per_step / 1000.0)
    k0 = r["k0"]
    ach = ALG_BYTES_PER_INSTANCE * B / (k0 / 1000.0) / 1e9
    roofline = {"bound": "hbm", "kernel": "tc_persistent_kernel (the one launch of a step; %d instances on this GPU)" % B, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": measured_traffic("tc_persistent_kernel", B), "kernel_ms": k0, "kernel_ms_per_rank_avg_max": r["k_all"], "kernel_share_of_step": k0 / ms_per_step, "peak_source": peak_src,
                "algorithmic_bytes_per_instance": ALG_BYTES_PER_INSTANCE, "bytes_actually_read_per_instance": r["in_bytes"] / max(B, 1), "ncu": measured_ncu("tc_persistent_kernel"),
                "note": "instance-resident shared-memory solver: bound by instruction fetch and dependent shared-memory latency, not HBM (SURVEY 8(d)); see profiles/"}

    # ---- end to end: HOST (pinned) buffers -> pt_batch_create (H2D) -> pt_batch_contain_sharded -> D2H of the result words -> free ----
    words = (total_B + 31) // 32
    full_bits = torch.zeros(words, dtype=torch.int32, device="cuda")
    h_bits = torch.zeros(max(words, 1), dtype=torch.int32).pin_memory()
    pinned_np = {k: (v.numpy() if hasattr(v, "numpy") else v) for k, v in pinned.items()}
    h2d = 0

    def time_e2e(pipelined):
        """K timed steps, each = pt_batch_create from pinned HOST arrays (H2D) + pt_batch_contain_sharded + D2H of the result words + a
        host wait for that read + pt_batch_free.  serial: one step after the other (the latency view).  pipelined: double-buffered --
        the host keeps ONE step queued behind the running one: while the kernel of step k runs, step k+1's batch is created (its
        upload overlaps the solve), its solve and result read are enqueued, and only then the host waits for step k's result and frees
        it (two pinned input buffers, two result buffers, two handles alive).  Step 0 is issued before the clock starts (pipeline
        fill) and the last timed iteration still issues a successor, so the timed region holds exactly K creates (K uploads), K
        launches and K result reads."""
        nonlocal h2d
        bufs = [pinned_np, pinned_np2]
        hb = [h_bits, h_bits2]
        state = {"k": 0, "inflight": None}

        def issue(k):
            nonlocal h2d
            b = pt.Batch(bufs[k % 2], stream=stream)
            h2d = b.h2d_bytes()
            comm.contain_sharded(b, first, total_B, full_bits, None, stream=stream)
            hb[k % 2].copy_(full_bits, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            return (b, ev, k)

        def retire(x):
            b, ev, k = x
            ev.synchronize()
            state["disp"] = hb[k % 2]
            b.free()

        def step():
            k = state["k"]; state["k"] = k + 1
            if pipelined:
                nxt = issue(k + 1)
                retire(state["inflight"])
                state["inflight"] = nxt
            else:
                retire(issue(k))
        if pipelined:
            state["inflight"] = issue(0)
        for _ in range(args.warmup):
            step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        wall_ms = (time.perf_counter() - t0) * 1000.0
        barrier()
        if pipelined:
            retire(state["inflight"])
        torch.cuda.synchronize()
        ms = max_over_ranks(max(e0.elapsed_time(e1), wall_ms))  # the host waits inside every step: take the larger of device and wall time
        disp = int(np.unpackbits(state["disp"].numpy().view(np.uint8), bitorder="little")[:total_B].sum())
        return total_B / (ms / args.steps / 1000.0), disp

    h_bits2 = torch.zeros(max(words, 1), dtype=torch.int32).pin_memory()
    pinned_np2 = {k: (torch.from_numpy(v.copy()).pin_memory().numpy() if isinstance(v, np.ndarray) else v) for k, v in pinned_np.items()}
    e2e_serial, disp_serial = time_e2e(False)
    e2e_val, e2e_disp = time_e2e(True)
    if disp_serial != e2e_disp:
        e2e_disp = -1
    h2d_total = sum_over_ranks(h2d)

    out = {"metric": "tree-containment instances/s", "value": value, "unit": "instances/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int16", "data": "synthetic",
           "config": {"workload": "ONE batch of %d gen networks l=100 r=20 with a displayed tree each (BASELINE configs[1]), sharded over %d GPU(s)" % (total_B, world),
                      "instances": total_B, "instances_on_this_gpu": B, "leaves": L, "reticulations": R, "host_nodes": 2 * R + 2 * L - 1, "generator_seed": SEED,
                      "input_layout": {4: "int32 arrays", 2: "int16 arrays (pt_batch_desc.index_bytes = 2)", 1: "uint8 arrays, out-degrees instead of offsets (pt_batch_desc.index_bytes = 1)"}[layout],
                      "l2": "inputs larger than L2: the resident shard (%.1f MB on this GPU) is kept in %d device copies (%.0f MB >= 2 x the 126 MB L2) and step i reads copy i mod %d, so no step finds its input in L2; the end-to-end steps upload fresh bytes every step; no explicit flush" % (r["in_bytes"] / 1e6, r["copies"], r["in_bytes"] * r["copies"] / 1e6, r["copies"]),
                      "parallelism": ("instances sharded %d-way (contiguous ranges); result words exchanged by the library (pt_batch_contain_sharded) %s" %
                                      (world, "through peer memory: every rank pushes its words into the other ranks' buffers over NVLink, no reduction" if comm.peer_exchange()
                                       else "with one NCCL all-reduce")) if world > 1 else "single GPU",
                      "cpu_affinity": ("%d cores next to the GPU" % numa) if numa else "default"},
           "e2e": {"value": e2e_val, "unit": "instances/s", "h2d_bytes_per_step": h2d_total, "d2h_bytes_per_step": words * 4 * world,
                   "api": "pt_batch_create(HOST pinned) + pt_batch_contain_sharded + D2H of the result words + host sync + pt_batch_free, on every rank, every step",
                   "mode": "double-buffered: while step k's kernel runs the host creates step k+1's batch (its H2D overlaps the solve) and queues its solve + result read, then waits for step k's result; K creates, K launches, K result reads inside the timed region",
                   "serial_value": e2e_serial, "serial_mode": "one step after the other (upload chunks overlap only the step's own kernel)", "displayed": e2e_disp},
           "gpu_launches": r["launches"] * world, "clocks": clocks, "roofline": roofline,
           "verdicts": {"displayed": r["displayed"], "errors": r["errors"], "expected_displayed": total_B}, "counters": r["counters"],
           "host_generation_s": t_gen}
    del pinned, pinned_np, pinned_np2
    # ---- extra: weak scaling, 10^5 instances per GPU ----
    if world > 1 and not args.no_weak:
        wf, wB, wp = make_shard(args.instances * world, SEED + (1 << 30))
        wr = run_resident(args.instances * world, wf, wB, wp, args.steps, args.warmup)
        out["weak_scaling"] = {"instances": args.instances * world, "instances_per_gpu": wB, "value": args.instances * world / (wr["ms"] / 1000.0), "unit": "instances/s",
                               "ms_per_step": wr["ms"], "displayed": wr["displayed"], "errors": wr["errors"]}
        del wp
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline_leg()
    if not args.no_lca:  # every rank: queries are sharded
        try:
            out["lca"] = lca_leg(pt, torch, args.lca_leaves, args.lca_queries, peak, rank, world, dist if world > 1 else None)
        except Exception as ex:  # the secondary leg must not take the headline line down
            if world > 1:
                raise
            out["lca"] = {"error": repr(ex)}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        for key, leg in (("newick_ingest", lambda: ingest_leg(pt)), ("iso", lambda: iso_leg(pt, torch)), ("newick_on_gpu", lambda: text_leg(pt, torch)),
                         ("config1", lambda: config1_leg(pt)), ("who_displays", lambda: who_leg(pt, torch)), ("bridges", lambda: bridges_leg(pt, torch))):
            try:
                out[key] = leg()
            except Exception as ex:
                out[key] = {"error": repr(ex)}
    if not args.no_config4:  # every rank: 8 instances each
        try:
            out["config4"] = config4_leg(pt, torch, args.config4_instances, rank, world, dist if world > 1 else None)
        except Exception as ex:
            if world > 1:
                raise
            out["config4"] = {"error": repr(ex)}
    if not args.no_config5:  # every rank: the index range is sharded
        try:
            out["config5"] = config5_leg(pt, rank=rank, world=world, dist=dist if world > 1 else None, torch=torch, comm=comm)
        except Exception as ex:
            if world > 1:
                raise
            out["config5"] = {"error": repr(ex)}
    comm.free()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        os.write(json_fd, (json.dumps(out) + "\n").encode())


if __name__ == "__main__":
    main()
