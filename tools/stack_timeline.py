"""Per-phase cycle timeline of dit_stack_kernel (clock64 stamps of worker warp 0, second tile of every CTA, one layer).
Usage (GPU box): python tools/stack_timeline.py [cells] [layer]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import build_models
from scldm_b200 import _lib
from scldm_b200.transport.transport import FusedCFGModel

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
layer = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
ldm, dcfg, vcfg = build_models(dev)
ldm.cell_chunk = cells
fn = ldm.transport_sampler.sample_ode(sampling_method="euler", num_steps=4)
lab = torch.randint(0, 14, (cells,), device=dev)
z = torch.randn(cells, 16, 16, device=dev)
cond = {"clusters": torch.cat([lab, lab])}
model = FusedCFGModel(ldm.diffusion_model, {"clusters": 2.0})
fn(torch.cat([z, z]), model, condition=cond)
torch.cuda.synchronize()
buf = torch.zeros(148 * 64, dtype=torch.int64, device=dev)
_lib.load().scldm_debug_timeline(buf.data_ptr(), layer)
fn(torch.cat([z, z]), model, condition=cond)
torch.cuda.synchronize()
_lib.load().scldm_debug_timeline(None, -1)
b = buf.cpu().view(148, 64)
b = b[b[:, 4] > 0]
rel = (b - b[:, 4:5]).float()
rel[b == 0] = float("nan")
med = rel.nanmedian(0).values.tolist()
names = {4: "attn start (A tile written)", 5: "layer end"}
for hp in range(4):
    names[8 + 3 * hp] = f"  hp{hp} accQ full"
    names[9 + 3 * hp] = f"  hp{hp} q/k/v staged"
    names[10 + 3 * hp] = f"  hp{hp} core + AO done"
for base, nm in ((20, "attn->mlp boundary"), (40, "mlp->attn boundary")):
    names[base] = f"{nm}: accB full"
    names[base + 1] = f"{nm}: pass 1 done"
    names[base + 2] = f"{nm}: stats combined"
    names[base + 3] = f"{nm}: pass 2 done (a_ready)"
    names[base + 4] = f"{nm}: last chunk done on every warp (region free)"
    names[base + 5] = f"{nm}: x_old staged in TMEM"
    names[base + 7] = f"{nm}: drain warps done (x_new stored)"
for j in range(6):
    names[28 + 2 * j] = f"  chunk{j} acc1 full"
    names[29 + 2 * j] = f"  chunk{j} H written"
order = sorted((v, i) for i, v in enumerate(med) if v == v and i in names)
prev = 0.0
print(f"{len(b)} CTAs, layer {layer}: median cycles since the attention half's start (delta to previous stamp)")
for v, i in order:
    print(f"  {int(v):7d} (+{int(v - prev):5d})  {names[i]}")
    prev = v
