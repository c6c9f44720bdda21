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
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
dev = torch.device("cuda:0")
ldm, dcfg, vcfg = build_models(dev)
ldm.cell_chunk = cells
fn = ldm.transport_sampler.sample_ode(sampling_method="euler", num_steps=steps)
lab = torch.randint(0, 14, (cells,), device=dev)
z = torch.randn(cells, 16, 16, device=dev)
cond = {"clusters": torch.cat([lab, lab])}
model = FusedCFGModel(ldm.diffusion_model, {"clusters": 2.0})
fn(torch.cat([z, z]), model, condition=cond)
torch.cuda.synchronize()
buf = torch.zeros(148 * 128, dtype=torch.int64, device=dev)
_lib.load().scldm_debug_timeline(buf.data_ptr(), layer)
fn(torch.cat([z, z]), model, condition=cond)
torch.cuda.synchronize()
_lib.load().scldm_debug_timeline(None, -1)
b = buf.cpu().view(148, 128)
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

tot = (b[:, 61] - b[:, 60]).float()
ns = (b[:, 63] - b[:, 62]).float()
print(f"per CTA: kernel body {tot.median():.0f} cycles median (min {tot.min():.0f}, max {tot.max():.0f}); wall {ns.median():.0f} ns median "
      f"(max {ns.max():.0f}) -> SM clock {1000 * (tot / ns).median():.0f} MHz; grid span {int(b[:, 63].max() - b[:, 62].min())} ns")
ts = (b[:, 52:56] - b[:, 60:61]).float()
print("tile starts (cycles since kernel start, median):", [int(v) for v in ts.median(0).values.tolist()])
print(f"cycles per tile-evaluation: {float(tot.median()) / 3 / (steps - 1):.0f} (whole-solve mode; 3 tiles per CTA, {steps - 1} evaluations)")

if (b[:, 48] > 0).any():
    d = b
    print("whole-solve step tail (median cycles): final boundary done -> final Linear seen", int((d[:, 49] - d[:, 48]).float().median()),
          " -> state updated, z written", int((d[:, 50] - d[:, 49]).float().median()), " | layer 0 LN boundary done -> (8 layers) -> final boundary done",
          int((d[:, 48] - d[:, 51]).float().median()), " | z written -> next evaluation's layer 0 LN boundary done (previous stamp pair, same evaluation index shifted)",
          int((d[:, 51] - d[:, 50]).float().median()))

    period = float(tot.median()) / 3 / (steps - 1)
    rel = lambda k: int((d[:, k] - d[:, 50]).float().median() + period)  # noqa: E731
    print("input-projection boundary, cycles after 'z written' of the previous evaluation: region free", rel(68), " x_old staged", rel(69), " accB full", rel(64),
          " pass 1 done", rel(65), " stats", rel(66), " pass 2 done", rel(67))
    print("  per-warp arrival at the region-free barrier:", [rel(80 + w) for w in range(16)])
