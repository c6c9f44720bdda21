"""Where a training step's time goes: CPU enqueue time vs GPU time of forward / backward / optimizer, eager vs CUDA-graph replay.
    python tools/train_profile.py [--batch 128]
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scldm_b200 import synthetic  # noqa: E402
from scldm_b200.config import DiTConfig  # noqa: E402
from scldm_b200.nnets import DiT  # noqa: E402
from scldm_b200.training import DiTTrainer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--once", action="store_true", help="one eager step only (for ncu)")
args = ap.parse_args()
dev = torch.device("cuda")
cfg = DiTConfig(class_vocab_sizes={"cell_type": 50})
dit = DiT(**cfg.kwargs())
dit.load_state_dict(synthetic.dit_state_dict(cfg, 1234))
dit = dit.to(dev).train()
tr = DiTTrainer(dit, lr=5e-4, max_grad_norm=10.0)
B = args.batch
xt = synthetic.randn("p.x", (B, 16, 16)).to(dev)
ut = synthetic.randn("p.u", (B, 16, 16)).to(dev)
t = torch.rand(B).to(dev)
cls = tr.cls_rows({"cell_type": synthetic.randint("p.l", 50, (B,)).to(dev)}, B)
dv = torch.empty_like(xt)


def step(events=False):
    v = tr.forward(xt, t, cls)
    torch.sub(v, ut, out=dv)
    dv.mul_(2.0 / (B * 256))
    tr.backward(dv, with_events=events)
    tr.optimizer_step()


for _ in range(3):
    step()
torch.cuda.synchronize()
if args.once:
    step()
    torch.cuda.synchronize()
    sys.exit(0)


def gpu_ms(fn, n=20):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    cpu = (time.perf_counter() - t0) / n * 1e3
    b.synchronize()
    return a.elapsed_time(b) / n, cpu


print("eager step            gpu %.3f ms  cpu-enqueue %.3f ms" % gpu_ms(step))
print("eager forward         gpu %.3f ms  cpu-enqueue %.3f ms" % gpu_ms(lambda: tr.forward(xt, t, cls)))
print("eager backward        gpu %.3f ms  cpu-enqueue %.3f ms" % gpu_ms(lambda: tr.backward(dv, with_events=False)))
print("eager optimizer       gpu %.3f ms  cpu-enqueue %.3f ms" % gpu_ms(tr.optimizer_step))
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    step()
    with torch.cuda.graph(g, stream=s):
        step()
torch.cuda.current_stream().wait_stream(s)
print("graph replay step     gpu %.3f ms  cpu-enqueue %.3f ms" % gpu_ms(g.replay))
