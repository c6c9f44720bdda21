"""2-rank NCCL test of the data-parallel VAE training step (reference: Lightning DDP around VAE.training_step,
experiments/scripts/train.py).  Needs two GPUs on the box (`gpurun --gpus 2`); skipped otherwise."""

import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_vae_training_matches_single_process():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29613",
           os.path.join(ROOT, "tests", "ddp_vae_train_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("DDP_RESULT ")][-1]
    out = json.loads(line[len("DDP_RESULT "):])
    print(out)
    for r in out:
        assert r["max_abs_diff_across_ranks"] == 0.0, r          # identical replicas after the all-reduce + optimizer
    # 2 x 4 cells with the 1 / global-batch loss scale == 8 cells in one process, up to the fp32 summation order of the atomics
    assert out[0]["grad_rel_l2_vs_single_process"] < 1e-4, out[0]
    assert out[0]["losses"][-1] < out[0]["losses"][0]
