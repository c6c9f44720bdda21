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
names = ["qkv", "proj", "mlp1", "mlp2"]
for k, name in enumerate(names):
    t = b[k, :n_cta]
    t0 = t[:, 0:1]
    rel = (t - t0).float()
    rel[t == 0] = float("nan")
    med = rel.nanmedian(0).values
    print(name, "n_cta", n_cta, "median cycles since CTA start per stamp:")
    print("   ", [(i, int(v)) for i, v in enumerate(med.tolist()) if v == v and i > 0])
    start_spread = (t[:, 0] - t[:, 0].min()).float()
    end = (t[:, 31] - t[:, 0].min()).float()
    print("    CTA start spread max", int(start_spread.max()), "kernel span (first start -> last end)", int(end.max()))
