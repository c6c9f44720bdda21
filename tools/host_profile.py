"""Host-side cost of LatentDiffusion.sample (what the GPU waits for at the start of a step): cProfile of a few calls.
Usage (GPU box): python tools/host_profile.py [cells]"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import GUIDANCE, build_models
from scldm_b200 import synthetic

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
dev = torch.device("cuda:0")
ldm, dcfg, vcfg = build_models(dev)
lab = {k: synthetic.randint("hp." + k, v, (B,)).to(dev) for k, v in dcfg.class_vocab_sizes.items()}
gw = {k: GUIDANCE for k in dcfg.class_vocab_sizes}
genes = torch.arange(1, vcfg.n_genes + 1, device=dev).unsqueeze(0).expand(B, -1)
for _ in range(3):
    ldm.sample(lab, gw, B, genes)
torch.cuda.synchronize()
from scldm_b200 import ops

first = []
orig = ops.dit_sample_ode


def probe(*a, **k):
    r = orig(*a, **k)
    first.append(time.perf_counter())
    return r


ops.dit_sample_ode = probe
import scldm_b200.transport.transport as tr

tr.ops.dit_sample_ode = probe
for _ in range(3):
    first.clear()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ldm.sample(lab, gw, B, genes)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"host: first solve queued after {1e3 * (first[0] - t0):.2f} ms; whole sample() enqueued after {1e3 * (t1 - t0):.2f} ms; GPU done after {1e3 * (t2 - t0):.2f} ms")
ops.dit_sample_ode = orig
tr.ops.dit_sample_ode = orig
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    ldm.sample(lab, gw, B, genes)
    torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
