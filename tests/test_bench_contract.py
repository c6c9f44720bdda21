"""bench.py contract (CPU part): the reference arm prints ONE JSON line with the keys the driver reads, and the GPU arm refuses to
run without a CUDA device instead of falling back to the CPU."""

import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-batch", "2"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1                                   # exactly one line on stdout
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "cells/s" and j["higher_is_better"] is True and j["vs_baseline"] is None
    assert j["metric"].startswith("generated cells/sec") and j["value"] > 0 and j["steps"] == 1
    from oracle import ref_loader

    # the reference's own modules when they are reachable (/root/reference or the copies staged under oracle/_ref), the oracle port otherwise
    assert j["cpu_baseline"]["kind"] == ("reference" if ref_loader.reference_available() else "port")
    assert j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in j["config"] and "dentate_gyrus" in j["config"]["workload"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_has_no_cpu_fallback():
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "CUDA" in (r.stderr + r.stdout)
