"""2-rank NCCL test of sharded generation through the product API (`scldm_b200.dist.sample_sharded`): the gathered output of two
ranks is bit-identical to the single-process output (cells are independent; Philox streams are keyed by the global cell index).
Needs two GPUs on the box (`gpurun --gpus 2`); skipped otherwise."""

import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_sharded_sample_is_bit_identical_to_one_rank():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29612",
           os.path.join(ROOT, "tests", "dist_sample_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    out = json.loads([l for l in res.stdout.splitlines() if l.startswith("DIST_RESULT ")][-1][len("DIST_RESULT "):])
    print(out)
    for r in out:
        assert r["counts_equal"] and r["z_equal"] and r["second_call_equal"] and r["calls_differ"], r
        assert r["shape"] == [74, 1200] and r["nnz"] > 0
