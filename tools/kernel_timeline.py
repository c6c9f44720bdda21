"""Per-phase cycle timeline of the DiT GEMM kernels (clock64 stamps written by CTA thread 0 / warp leaders).
Usage (GPU box): python tools/kernel_timeline.py [cells]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import build_models
from scldm_b200 import _lib

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 392
dev = torch.device("cuda:0")
ldm, dcfg, vcfg = build_models(dev)
ldm.cell_chunk = cells
ldm.num_steps = 4
lab = {"clusters": torch.randint(0, 14, (cells,), device=dev)}
genes = torch.arange(1, 101, device=dev).unsqueeze(0).expand(cells, -1)
ldm.vae_model.decode_counts = lambda *a, **k: (None, None, None)  # DiT only
buf = torch.zeros(4 << 17, dtype=torch.int64, device=dev)
for _ in range(2):
    ldm.transport_sampler  # warm
from scldm_b200.transport.transport import FusedCFGModel

fn = ldm.transport_sampler.sample_ode(sampling_method="euler", num_steps=4)
z = torch.randn(cells, 16, 16, device=dev)
cond = {"clusters": torch.cat([lab["clusters"], lab["clusters"]])}
fn(torch.cat([z, z]), FusedCFGModel(ldm.diffusion_model, {"clusters": 2.0}), condition=cond)
torch.cuda.synchronize()
_lib.load().scldm_debug_timeline(buf.data_ptr(), 3)
fn(torch.cat([z, z]), FusedCFGModel(ldm.diffusion_model, {"clusters": 2.0}), condition=cond)
torch.cuda.synchronize()
_lib.load().scldm_debug_timeline(None, -1)
b = buf.cpu().view(4, -1, 32)
n_cta = (3 * cells + 7) // 8
if os.environ.get("SCLDM_MEGA", "1") != "0":
    n_cta = min(n_cta, 148)   # persistent kernel: one CTA per SM, stamps of its second tile
names = ["qkv", "proj", "mlp1", "mlp2"]
spans = {}
for k, name in enumerate(names):
    tt = b[k, :n_cta]
    if (tt[:, 29] > 0).any():
        spans[name] = (int(tt[:, 29].min()), int(tt[:, 30].max()))
print("wall ns (last layer): ", {k: v[1] - v[0] for k, v in spans.items()})
if "proj" in spans:
    print(" qkv end -> proj start (attn + 2 gaps)", spans["proj"][0] - spans["qkv"][1], " proj end -> mlp start", spans["mlp1"][0] - spans["proj"][1])
else:
    print(" attn_block end -> mlp start", spans["mlp1"][0] - spans["qkv"][1])
for k, name in enumerate(names):
    if name not in spans:
        continue
    t = b[k, :n_cta]
    t0 = t[:, 0:1]
    rel = (t - t0).float()
    rel[t == 0] = float("nan")
    med = rel.nanmedian(0).values
    print(name, "n_cta", n_cta, "median cycles since CTA start per stamp:")
    print("   ", [(i, int(v)) for i, v in enumerate(med.tolist()) if v == v and i > 0 and not 28 <= i <= 30])
    if os.environ.get("SCLDM_PAIR") == "1":   # stamps of a pair are relative to each CTA's own phase start: show leader / peer
        for r, nm in ((0, "leader"), (1, "peer  ")):
            m2 = rel[r::2].nanmedian(0).values
            print("    ", nm, [(i, int(v)) for i, v in enumerate(m2.tolist()) if v == v and i > 0 and not 28 <= i <= 30])
    start_spread = (t[:, 0] - t[:, 0].min()).float()
    end = (t[:, 31] - t[:, 0].min()).float()
    g0, g1 = t[:, 29], t[:, 30]
    life = (g1 - g0).float()
    print("    wall ns: kernel span (first CTA start -> last CTA end)", int(g1.max() - g0.min()), " CTA life median", int(life.median()),
          "max", int(life.max()), " start spread: first-wave", int((g0.sort().values[min(147, n_cta - 1)] - g0.min())), "last", int(g0.max() - g0.min()))
    sm = t[:, 28]
    print("    CTAs per SM: max", int(torch.bincount(sm).max()), "SMs used", int((torch.bincount(sm) > 0).sum()))
