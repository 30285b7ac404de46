"""bench.py's contract on a machine without a GPU: the reference arm (the reference's own CPU path on the host cores) prints ONE JSON
line with the agreed keys, the product arm refuses to run without a CUDA device (no CPU fallback), and the roofline helpers read the
committed captures."""
import importlib.util
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def _bench_module():
    spec = importlib.util.spec_from_file_location("bench_under_test", BENCH)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--steps", "1", "--warmup", "1", "--ref-instances-per-core", "8"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [x for x in r.stdout.splitlines() if x.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "tree-containment instances/s" and d["unit"] == "instances/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1 and d["gpu_launches"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_without_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, BENCH, "--steps", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout)
    assert not [x for x in r.stdout.splitlines() if x.startswith("{")]     # no number without the CUDA path


def test_roofline_helpers_read_the_committed_captures():
    b = _bench_module()
    assert b.measured_traffic("tc_persistent_kernel", 100_000) == 127110400        # profiles/traffic_r02.json
    assert b.measured_traffic("tc_persistent_kernel", 12_500) is None              # a different launch size: no claim
    ncu = b.measured_ncu("tc_persistent_kernel")
    assert ncu and 0 < ncu["active_threads_per_instruction"] <= 32 and 0 < ncu["l2_hit_pct"] <= 100
    peak, src = b.measured_peak()
    assert peak > 1000 and src
    assert b.ALG_BYTES_PER_INSTANCE == 8124                                         # 4 940 + 3 184
