#!/usr/bin/env python
"""Benchmark of the scLDM generation hot path: generated cells/sec (ODE steps + CFG + VAE decode).

    python bench.py --gpus N --steps K --warmup W            # ours (one process per GPU; torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port), rank 0 only

A "step" is one `LatentDiffusion.sample()` call for B cells per GPU on the dentate_gyrus-shaped model
(BASELINE.json configs[1]): device-side size factors + noise -> 49-eval Euler ODE with classifier-free
guidance (3 DiT forwards per cell and eval, batched) -> fused VAE decode -> NB sampling; it returns 2B
generated rows (B unconditional + B guided), which is what `value` counts.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "generated cells/sec (ODE steps + CFG + VAE decode)"
DATASET = "dentate_gyrus"
NUM_STEPS = 50  # grid points => 49 Euler steps
GUIDANCE = 2.0


def measured_peaks() -> dict:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        d["source"] = "measured"
        return d
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name).read().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
            except ValueError:
                continue
            for n, v in zip(names, r[5:9]):
                if v.strip().lower() == "active":
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        load = [s for s, p in zip(sm, power) if p > 0.5 * max(power)] or sm
        return {"sm_mhz": statistics.median(load), "sm_max_mhz": max(mx), "power_w_max": max(power), "samples": len(sm),
                "reasons": sorted(reasons)}


def build_models(device, dataset=DATASET, seed=1234, method="euler"):
    from scldm_b200 import synthetic
    from scldm_b200.config import dataset_configs
    from scldm_b200.models import LatentDiffusion
    from scldm_b200.nnets import DiT
    from scldm_b200.transport import create_transport
    from scldm_b200.vae import TransformerVAE

    dcfg, vcfg = dataset_configs(dataset)
    dit = DiT(**dcfg.kwargs())
    dit.load_state_dict(synthetic.dit_state_dict(dcfg, seed))
    vae = TransformerVAE.from_config(vcfg)
    vae.load_state_dict(synthetic.vae_state_dict(vcfg, seed))
    mu_t, sd_t = synthetic.size_factor_tables(dcfg.class_vocab_sizes, seed)
    ldm = LatentDiffusion(vae.to(device).eval(), dit.to(device).eval(), create_transport("Linear", "velocity", "velocity"),
                          mu_size_factor=mu_t, sd_size_factor=sd_t, sampling_method=method, num_steps=NUM_STEPS, seed=4321)
    return ldm, dcfg, vcfg


def algorithmic_flops_per_row(G: int, evals: int = 49) -> dict:
    """DiT forward 210.8 MFLOP/cell; 49 evals (Euler); CFG => 1.5 forwards per output row; decode 2.2M + G*23104."""
    D, H, M, L, N = 256, 684, 16, 16, 8
    dit = M * (N * (8 * D * D + 6 * D * H + 4 * M * D) + 4 * L * D) + N * 12 * D * D + 4 * D * D + 2 * (256 * D + D * D)
    dec = 2.2e6 + G * 23104
    return {"dit_forward": dit, "decode": dec, "row_cfg": 1.5 * evals * dit + dec}


# per-launch algorithmic FLOPs of the GEMM kernel classes (rows = slots*16 actual, unpadded N/K)
def kernel_flops(name: str, rows: int, mod_rows: int, evals: int = 1) -> float | None:
    D, H = 256, 684
    return {
        "gemm_ares<LN,QKV>": 2.0 * rows * D * 3 * D,
        "gemm_astream<proj>": 2.0 * rows * D * D,
        "gemm_ares<LN,SWIGLU>": 2.0 * rows * D * 2 * H,
        "gemm_astream<mlp2>": 2.0 * rows * H * D,
        "gemm_ares<COND,MOD>": 2.0 * mod_rows * D * (8 * 6 * D + 2 * D),
        "attn16": 4.0 * rows * 16 * D,
        "mlp_fused": 2.0 * rows * D * 2 * H + 2.0 * rows * H * D,
        "attn_block": 2.0 * rows * D * 3 * D + 4.0 * rows * 16 * D + 2.0 * rows * D * D,
        "dit_blocks": 8 * (2.0 * rows * D * 3 * D + 4.0 * rows * 16 * D + 2.0 * rows * D * D + 2.0 * rows * D * 2 * H + 2.0 * rows * H * D),
        "dit_stack": 8 * (2.0 * rows * D * 3 * D + 4.0 * rows * 16 * D + 2.0 * rows * D * D + 2.0 * rows * D * 2 * H + 2.0 * rows * H * D),
        # whole-solve launch: the block stack of every evaluation + input projection (256 x 16) and final Linear (16 x 256) per row
        "dit_solve": evals * (8 * (2.0 * rows * D * 3 * D + 4.0 * rows * 16 * D + 2.0 * rows * D * D + 2.0 * rows * D * 2 * H + 2.0 * rows * H * D)
                              + 2 * 2.0 * rows * 16 * D),
    }.get(name)


def reference_kind() -> str:
    """"reference": the reference's own modules are staged under oracle/_ref (oracle/build_ref.py; they travel to the GPU box);
    "port": only the oracle restatement is available."""
    from oracle import ref_loader

    return "reference" if ref_loader.reference_available() else "port"


def cpu_sample(B: int, dcfg, vcfg, threads: int, device="cpu"):
    """The reference's generation path in eager PyTorch, fp32: its OWN modules (oracle/_ref, `kind: "reference"`) composed as
    `LatentDiffusion.sample` composes them, or - when they are not staged - the oracle restatement (`kind: "port"`)."""
    from oracle import scldm_oracle as O
    from scldm_b200 import synthetic

    if device == "cpu":
        torch.set_num_threads(threads)
    dsd, vsd = synthetic.dit_state_dict(dcfg, 1234), synthetic.vae_state_dict(vcfg, 1234)
    z0 = synthetic.randn("cpu.z0", (B, 16, 16)).to(device)
    lab = {k: synthetic.randint("cpu.lab." + k, v, (B,)).to(device) for k, v in dcfg.class_vocab_sizes.items()}
    w = {k: GUIDANCE for k in dcfg.class_vocab_sizes}
    lsf = (8.0 + 0.3 * synthetic.randn("cpu.lsf", (B,))).to(device)
    genes = torch.arange(1, vcfg.n_genes + 1, device=device).unsqueeze(0).expand(B, -1)
    if reference_kind() == "reference":
        from oracle import ref_loader

        ref_step = ref_loader.reference_sample_fn(dcfg, vcfg, dsd, vsd, device=device, num_steps=NUM_STEPS, method="euler")
        return lambda: ref_step(z0, lab, w, genes, lsf)
    dsd = {k: v.to(device) for k, v in dsd.items()}
    vsd = {k: v.to(device) for k, v in vsd.items()}

    def step():
        with torch.no_grad():
            mu, theta, z = O.latent_diffusion_sample(z0, lab, w, genes, lsf, dsd, dcfg, vsd, vcfg, num_steps=NUM_STEPS, method="euler")
            return O.nb_sample(mu, theta)

    return step


def workload_name(dataset, dcfg, vcfg, method="euler") -> str:
    cls = ", ".join(f"{k}:{v}" for k, v in dcfg.class_vocab_sizes.items())
    tag = " (BASELINE configs[1])" if dataset == "dentate_gyrus" and method == "euler" else ""
    evals = {"euler": f"{NUM_STEPS - 1} evals", "heun2": f"{2 * (NUM_STEPS - 1)} evals", "midpoint": f"{2 * (NUM_STEPS - 1)} evals",
             "dopri5": "adaptive, atol = rtol = 1e-5 (the reference's default solver)"}[method]
    return (f"{dataset}-shaped generation with CFG{tag}: G={vcfg.n_genes}, classes {{{cls}}} {dcfg.condition_strategy}, "
            f"sample_ode('{method}', num_steps={NUM_STEPS}) = {evals}, guidance {GUIDANCE}, decode + NB draw")


def gpu_eager_sample(B: int, dcfg, vcfg, device):
    """The same reference path as `cpu_sample` with every tensor on the GPU: what the reference's eager PyTorch generation costs on
    this device (library kernels, TF32 matmuls as `experiments/scripts/inference.py:26` sets them).  A reported comparator only -
    never part of `value` / `e2e`."""
    return cpu_sample(B, dcfg, vcfg, 0, device=device)


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores - its unmodified modules staged under
    oracle/_ref by oracle/build_ref.py (the oracle port only if they are absent).  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from scldm_b200.config import dataset_configs

    dcfg, vcfg = dataset_configs(args.dataset)
    threads = os.cpu_count() or 1
    B = args.ref_batch
    step = cpu_sample(B, dcfg, vcfg, threads)
    for _ in range(args.warmup if args.warmup < 2 else 1):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = 2 * B * args.steps / dt
    line = {
        "metric": METRIC, "value": val, "unit": "cells/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "impl": "reference",
        "config": {"workload": workload_name(args.dataset, dcfg, vcfg),
                   "cells_per_step": B, "rows_per_step": 2 * B, "note": "bounded sample of the GPU arm's workload; CPU only"},
        "cpu_baseline": {"value": val, "unit": "cells/s", "cores": threads, "kind": reference_kind(),
                         "sample": f"{args.steps} x sample() of {B} cells (2B={2 * B} rows), " + ("the reference's own modules (oracle/_ref)" if reference_kind() == "reference" else "oracle port of the reference modules")},
        "e2e": {"value": val, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ==================================================================================================================
# --mode train: one LDM training step (BASELINE configs[4]): frozen census-vocabulary VAE encode -> flow-matching loss ->
# DiT forward/backward -> DDP gradient all-reduce (NCCL, bucketed, overlapped with the backward) -> clip + AdamW.
# Reference: LatentDiffusion.training_step (src/scldm/models.py:634-666), per-GPU batch 128 (ldm_base.yaml:58).
# ==================================================================================================================
TRAIN_METRIC = "LDM training cells/sec (frozen MCAB encode + FM loss + DiT fwd/bwd + DDP all-reduce + AdamW)"
TRAIN_CLASSES = {"cell_type": 50}
TRAIN_S = 8000   # census genes_seq_len (datamodule/default.yaml:130)


def train_flops_per_cell() -> float:
    return 3.0 * algorithmic_flops_per_row(1)["dit_forward"]   # forward + backward (dgrad + wgrad) of the DiT; the frozen encode adds 52.9 MFLOP


def encoder_inputs(B: int, G: int, S: int, seed: int):
    import numpy as np

    rng = np.random.default_rng(seed)
    gs = np.zeros((B, S), dtype=np.int64)
    cs = np.zeros((B, S), dtype=np.float32)
    for b in range(B):
        n = int(rng.integers(int(0.05 * G), min(S, int(0.3 * G))))
        gs[b, :n] = np.sort(rng.choice(np.arange(1, G + 1), size=n, replace=False))
        cs[b, :n] = 1 + rng.poisson(2.0, size=n)
    return torch.from_numpy(cs), torch.from_numpy(gs)


def oracle_train_step_fn(B: int, dcfg, device, threads: int | None = None, use_reference_modules: bool = False):
    """fwd + bwd + clip + AdamW of the same DiT in eager PyTorch: the oracle port (CPU: the reference's flex_attention has no CPU
    backward) or the staged reference modules themselves (GPU).  Baseline / comparator only."""
    from oracle import scldm_oracle as O
    from scldm_b200 import synthetic

    if threads:
        torch.set_num_threads(threads)
    sd = synthetic.dit_state_dict(dcfg, 1234)
    z = synthetic.randn("trainb.z", (B, 16, 16)).to(device)
    lab = {k: synthetic.randint("trainb.lab." + k, v, (B,)).to(device) for k, v in dcfg.class_vocab_sizes.items()}
    if use_reference_modules:
        from oracle import ref_loader

        ref = ref_loader.load_reference()
        model = ref_loader.build_reference_dit(dcfg, sd).to(device).train()
        params = [p for p in model.parameters() if p.requires_grad]
        transport = ref.transport.create_transport(path_type="Linear", prediction="velocity", loss_weight="velocity", train_eps=1e-5, sample_eps=1e-5)

        def loss_fn():
            return transport.training_losses(model, z, {"condition": lab})["loss"].mean()
    else:
        sdg = {k: v.clone().to(device).requires_grad_(k != "pos_embed") for k, v in sd.items()}
        params = [v for v in sdg.values() if v.requires_grad]

        def loss_fn():
            x0 = torch.randn_like(z)
            t = torch.rand(B).to(device)
            return O.fm_training_losses(z, t, x0, lambda xt, tt: O.dit_forward(xt, tt, lab, sdg, dcfg))["loss"].mean()
    opt = torch.optim.AdamW(params, lr=5e-4, weight_decay=0.0)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = loss_fn()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 10.0)
        opt.step()
        return loss

    return step


def run_train(args):
    from scldm_b200.config import DiTConfig, VAEConfig

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dcfg = DiTConfig(class_vocab_sizes=dict(TRAIN_CLASSES))
    vcfg = VAEConfig(n_genes=36130)
    B = args.train_batch
    workload = (f"census-vocabulary LDM training step (BASELINE configs[4]): frozen TransformerVAE.encode G={vcfg.n_genes} S={TRAIN_S} -> FM loss "
                f"-> DiT (8 layers, D=256, classes {TRAIN_CLASSES}) fwd/bwd -> DDP all-reduce of {9.7:.1f} M fp32 grads -> clip 10 + AdamW; per-GPU batch {B}")
    if args.impl == "reference":
        if rank != 0:
            return
        threads = os.cpu_count() or 1
        Bc = args.ref_batch
        step = oracle_train_step_fn(Bc, dcfg, torch.device("cpu"), threads)
        step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        dt = time.perf_counter() - t0
        val = Bc * args.steps / dt
        print(json.dumps({
            "metric": TRAIN_METRIC, "value": val, "unit": "cells/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "impl": "reference", "config": {"workload": workload, "cells_per_step": Bc,
                                            "note": "DiT fwd/bwd + clip + AdamW of the oracle port with torch autograd on the host cores (the reference's flex_attention has no CPU backward; frozen encode not included)"},
            "cpu_baseline": {"value": val, "unit": "cells/s", "cores": threads, "kind": "port", "sample": f"{args.steps} training steps of {Bc} cells"},
            "e2e": {"value": val, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (scldm_b200 has no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=device)
    from scldm_b200 import ops, synthetic
    from scldm_b200.nnets import DiT
    from scldm_b200.training import DiTTrainer, wsd_schedule
    from scldm_b200.transport import create_transport
    from scldm_b200.vae import TransformerVAE

    dit = DiT(**dcfg.kwargs())
    dit.load_state_dict(synthetic.dit_state_dict(dcfg, 1234))   # identical replicas on every rank
    dit = dit.to(device).train()
    vae = TransformerVAE.from_config(vcfg)
    vae.load_state_dict(synthetic.vae_state_dict(vcfg, 1234))
    vae = vae.to(device).eval()
    trainer = DiTTrainer(dit, lr=5e-4 * world, max_grad_norm=10.0, lr_lambda=wsd_schedule(100_000, num_warmup_steps=1000),
                         ema_decay=0.9999, ema_update_every=10, ema_update_after_step=10_000, n_buckets=args.buckets)
    transport = create_transport("Linear", "velocity", "velocity")
    cs_h, gs_h = encoder_inputs(B, vcfg.n_genes, TRAIN_S, 500 + rank)
    cs_h, gs_h = cs_h.pin_memory(), gs_h.pin_memory()
    lab_h = {k: torch.randint(0, v, (B,), generator=torch.Generator().manual_seed(900 + rank)).pin_memory() for k, v in TRAIN_CLASSES.items()}
    cs_d, gs_d = cs_h.to(device), gs_h.to(device)
    lab_d = {k: v.to(device) for k, v in lab_h.items()}
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)

    def barrier():
        if world > 1:
            import torch.distributed as dist

            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        z = vae.encode(None, None, cs_d, gs_d)
        return trainer.fm_step(z, lab_d, transport)

    def step_e2e():
        cs, gs = cs_h.to(device, non_blocking=True), gs_h.to(device, non_blocking=True)
        lab = {k: v.to(device, non_blocking=True) for k, v in lab_h.items()}
        z = vae.encode(None, None, cs, gs)
        return float(trainer.fm_step(z, lab, transport))     # D2H read of the loss

    def timed(fn, steps):
        total = 0.0
        for _ in range(steps):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            b.synchronize()
            total += a.elapsed_time(b)
        return total

    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = ops.launch_count()
    ms = timed(step_device, args.steps)
    launches = ops.launch_count() - l0
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ar_ms = None
    if world > 1 and getattr(trainer, "_ar_events", None):
        ar_ms = trainer._ar_events[0].elapsed_time(trainer._ar_events[1])
    # exposed all-reduce time: the same steps with the collective switched off (gradients stay rank-local; timing only)
    noar_ms = None
    if world > 1:
        saved = trainer.world
        trainer.world = 1
        barrier()
        noar_ms = timed(step_device, args.steps)
        trainer.world = saved
        barrier()
    step_e2e()
    barrier()
    e2e_ms = timed(step_e2e, args.steps)
    barrier()
    t = torch.tensor([ms, e2e_ms, noar_ms or 0.0, ar_ms or 0.0], device=device, dtype=torch.float64)
    if world > 1:
        import torch.distributed as dist

        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms, noar_ms, ar_ms = (float(v) for v in t)

    breakdown, roofline = None, None
    peaks = measured_peaks()
    if rank == 0 and not args.no_prof:
        saved_world, trainer.world = trainer.world, 1      # rank-0-only step: no collective (the other ranks are not in it)
        ops.prof_enable(True, device)
        step_device()
        prof = ops.prof_summary()
        ops.prof_enable(False, device)
        trainer.world = saved_world
        tot = sum(v[1] for v in prof.values())
        breakdown = {k: {"launches": v[0], "ms": round(v[1], 3), "share": round(v[1] / tot, 4)} for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])}
        gemm_ms = sum(v[1] for k, v in prof.items() if k.startswith("trn_gemm") or k.startswith("trn_dgrad") or k.startswith("trn_wgrad"))
        gemm_n = sum(v[0] for k, v in prof.items() if k.startswith("trn_gemm") or k.startswith("trn_dgrad") or k.startswith("trn_wgrad"))
        rows, D, H, L = B * 16, 256, 684, 8
        blk = 2.0 * rows * (D * 3 * D + D * D + D * 2 * H + H * D)            # the four GEMMs of a block, forward
        modf = 2.0 * B * D * (L * 6 * D + 2 * D)
        fl = 3.0 * (L * blk + modf)                                           # forward + dgrad + wgrad
        ach = fl / (gemm_ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "kernel": "trn::gemm_kernel (all forward / dgrad / wgrad launches of a step)", "achieved": round(ach, 2),
                    "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": round(ach / peaks["bf16_tflops_sustained"], 4), "traffic": None,
                    "peak_source": peaks["source"] + " (sustained cuBLAS bf16)", "avg_launch_us": round(1e3 * gemm_ms / max(gemm_n, 1), 2),
                    "flops_per_step": fl, "launches_per_step": gemm_n,
                    "note": "2048-row GEMMs (batch 128 x 16 tokens): 16 row tiles per launch, latency-bound; the step is bounded by its ~260 dependent launches"}

    cpu_baseline, gpu_eager = None, None
    if rank == 0 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        stepc = oracle_train_step_fn(args.cpu_batch, dcfg, torch.device("cpu"), threads)
        stepc()
        t0 = time.perf_counter()
        stepc()
        dtc = time.perf_counter() - t0
        cpu_baseline = {"value": args.cpu_batch / dtc, "unit": "cells/s", "cores": threads, "kind": "port", "seconds": round(dtc, 2),
                        "sample": f"1 training step (DiT fwd/bwd + clip + AdamW, torch autograd through the oracle port) of {args.cpu_batch} cells"}
    if rank == 0 and not args.no_gpu_eager_train:
        prev = torch.get_float32_matmul_precision()
        try:
            torch.set_float32_matmul_precision("high")   # as experiments/scripts/train_ldm.py:18
            try:
                stepg = oracle_train_step_fn(B, dcfg, device, use_reference_modules=True)
                kind = "reference"
            except Exception:
                stepg = oracle_train_step_fn(B, dcfg, device)
                kind = "port"
            for _ in range(3):
                stepg()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                stepg()
            b.record(); b.synchronize()
            gpu_eager = {"value": 5 * B / (a.elapsed_time(b) / 1e3), "unit": "cells/s", "kind": kind, "ms_per_step": a.elapsed_time(b) / 5,
                         "sample": f"5 eager PyTorch training steps of {B} cells on the same GPU (DiT fwd/bwd + clip + AdamW, TF32 'high'; no encode, no DDP)"}
        except Exception as e:
            gpu_eager = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
        finally:
            torch.set_float32_matmul_precision(prev)

    if rank == 0:
        cells = B * world
        line = {
            "metric": TRAIN_METRIC, "value": cells * args.steps / (ms / 1e3), "unit": "cells/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "steps_per_s": args.steps / (ms / 1e3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "mode": "train",
            "config": {"workload": workload, "cells_per_step_per_gpu": B, "l2": "256 MB flush buffer written between timed steps",
                       "parallelism": f"dp{world}: replicated DiT, flat fp32 gradient buffer all-reduced in {len(trainer.bucket_bounds) - 1} buckets on a side stream",
                       "grad_bytes": trainer.n_params * 4, "algorithmic_gflop_per_cell": round(train_flops_per_cell() / 1e9, 4)},
            "clocks": clocks, "gpu_launches": int(launches),
            "model_tflops": round(cells * args.steps / (ms / 1e3) * train_flops_per_cell() / 1e12, 2),
            "e2e": {"value": cells * args.steps / (e2e_ms / 1e3), "unit": "cells/s", "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": cs_h.numel() * 4 + gs_h.numel() * 8 + sum(v.numel() * 8 for v in lab_h.values()), "d2h_bytes_per_step": 4},
        }
        if world > 1:
            exposed = max(ms - noar_ms, 0.0) / args.steps
            line["allreduce"] = {"ms_per_step": ar_ms, "exposed_ms_per_step": exposed, "overlap_frac": (1.0 - exposed / ar_ms) if ar_ms else None,
                                 "ms_per_step_without_collective": noar_ms / args.steps,
                                 "note": "ms_per_step = first bucket start to last bucket end on the side stream (includes waiting for the backward); exposed = step time minus the same step with the collective switched off"}
        if roofline:
            line["roofline"] = roofline
            line["kernel_breakdown"] = breakdown
        if cpu_baseline:
            line["cpu_baseline"] = cpu_baseline
        if gpu_eager:
            line["gpu_eager_baseline"] = gpu_eager
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


def quiet_stdout():
    """Multi-rank runs: NCCL writes its version banner to fd 1 at the first collective.  Point fd 1 at stderr for the whole run and
    return a file object on the original stdout for the single JSON line."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(saved, "w")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=2368, help="cells per step per GPU (2x rows are generated)")
    ap.add_argument("--chunk", type=int, default=0, help="cells per ODE chunk (0 = library default)")
    ap.add_argument("--strong", action="store_true", help="strong scaling: --batch is the GLOBAL batch, split over the GPUs (default: weak, --batch cells per GPU)")
    ap.add_argument("--ref-batch", type=int, default=64, help="cells per step of the CPU reference arm (BASELINE configs[0]: batch 64)")
    ap.add_argument("--cpu-batch", type=int, default=32)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-prof", action="store_true")
    ap.add_argument("--method", default="euler", choices=["euler", "heun2", "midpoint", "dopri5"],
                    help="ODE solver of sample_ode; the headline line is the fixed-grid Euler of BASELINE configs[1]")
    ap.add_argument("--gpu-eager", action="store_true", help="(default) kept for older command lines")
    ap.add_argument("--no-gpu-eager", action="store_true",
                    help="skip the reference modules run eagerly on the same GPU (PyTorch library kernels, TF32 'high'): the honest library-call comparator")
    ap.add_argument("--eager-batch", type=int, default=128, help="cells per eager step (the reference's default generation batch, generation.yaml:16)")
    ap.add_argument("--dataset", default=DATASET, choices=["dentate_gyrus", "hlca", "tabula_muris", "parse1m", "replogle"],
                    help="gene-vocabulary / class-table shape (BASELINE configs 2-4); the headline line is dentate_gyrus")
    ap.add_argument("--mode", default="generate", choices=["generate", "train"], help="generate: the headline generation step; train: one LDM training step (BASELINE configs[4])")
    ap.add_argument("--train-batch", type=int, default=128, help="cells per training step per GPU (ldm_base.yaml:58)")
    ap.add_argument("--buckets", type=int, default=3, help="gradient all-reduce buckets of the training step")
    ap.add_argument("--no-gpu-eager-train", action="store_true")
    args = ap.parse_args()

    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        out = quiet_stdout()
        _print = print
        import builtins

        builtins.print = lambda *a, **k: _print(*a, **{**k, "file": k.get("file") or out, "flush": True})
    if args.mode == "train":
        run_train(args)
        return
    if args.impl == "reference":
        run_reference(args)
        return

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (scldm_b200 has no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist

        # keep stdout to the single JSON line: NCCL prints its version banner (NCCL_DEBUG >= VERSION) to stdout
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=device)
    from scldm_b200 import ops

    ldm, dcfg, vcfg = build_models(device, args.dataset, method=args.method)
    if args.chunk > 0:
        ldm.cell_chunk = args.chunk
    B, G = (max(8, args.batch // world) if args.strong else args.batch), vcfg.n_genes
    gw = {k: GUIDANCE for k in dcfg.class_vocab_sizes}
    # the GLOBAL batch (B cells per GPU, weak scaling) is described once, identically on every rank; `dist.sample_sharded` - the
    # product's multi-GPU call - makes every rank generate its contiguous slice (Philox streams keyed by the global cell index)
    from scldm_b200 import dist as sdist

    Bg = B * world
    gen = torch.Generator().manual_seed(100)
    labels_gh = {k: torch.randint(0, v, (Bg,), generator=gen) for k, v in dcfg.class_vocab_sizes.items()}
    a0, a1 = sdist.shard_range(Bg, rank, world)
    labels_h = {k: v[a0:a1].clone().pin_memory() for k, v in labels_gh.items()}
    genes_row_h = torch.arange(1, G + 1, dtype=torch.int64).pin_memory()
    labels_g = {k: v.to(device) for k, v in labels_gh.items()}
    labels_d = {k: v[a0:a1] for k, v in labels_g.items()}
    genes_g = genes_row_h.to(device).unsqueeze(0).expand(Bg, -1)   # (B,G) view of one row, as the tokenizer tiles it
    genes_d = genes_g[a0:a1]
    counts_h = torch.empty(2 * B, G, dtype=torch.float32).pin_memory()
    z_h = torch.empty(2 * B, 16, 16, dtype=torch.float32).pin_memory()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)  # > 126 MB L2

    def barrier():
        if world > 1:
            import torch.distributed as dist

            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        return sdist.sample_sharded(ldm, labels_g, gw, Bg, genes_g, gather=False)

    def step_gather():
        return sdist.sample_sharded(ldm, labels_g, gw, Bg, genes_g, gather=True)    # + NCCL all_gather of counts and z to every rank

    def step_e2e():
        lab = {k: v.to(device, non_blocking=True) for k, v in labels_h.items()}
        genes = genes_row_h.to(device, non_blocking=True).unsqueeze(0).expand(B, -1)
        ldm.sample(lab, gw, B, genes, host_out=(counts_h, z_h))   # rows are copied out chunk by chunk, overlapping the next chunk's ODE
        torch.cuda.current_stream().synchronize()
        return counts_h, z_h

    csr_bytes = [0]

    def step_e2e_csr():
        """same call, count matrix sparsified on the device (what the reference's host-side process_generation_output builds)"""
        lab = {k: v.to(device, non_blocking=True) for k, v in labels_h.items()}
        genes = genes_row_h.to(device, non_blocking=True).unsqueeze(0).expand(B, -1)
        (indptr, indices, data), z = ldm.sample_csr(lab, gw, B, genes)
        out = [t.to("cpu", non_blocking=True) for t in (indptr, indices, data)]
        z_h.copy_(z, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        csr_bytes[0] = sum(t.numel() * t.element_size() for t in out) + z_h.numel() * 4
        return out

    def timed(fn, steps):
        """device time of `steps` calls (CUDA events per step on the launching stream, L2 flushed between steps)."""
        total = 0.0
        for _ in range(steps):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            b.synchronize()
            total += a.elapsed_time(b)
        return total  # ms

    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ops.launch_count()
    ms = timed(step_device, args.steps)
    launches = ops.launch_count() - launches0
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    e2e_ms = None
    if not args.no_e2e:
        step_e2e()
        barrier()
        e2e_ms = timed(step_e2e, args.steps)
        barrier()
    e2e_csr_ms = None
    if not args.no_e2e:
        step_e2e_csr()
        barrier()
        e2e_csr_ms = timed(step_e2e_csr, args.steps)
        barrier()
    gather_ms = None
    if world > 1:
        step_gather()
        barrier()
        gather_ms = timed(step_gather, args.steps)
        barrier()
    t = torch.tensor([ms, e2e_ms or 0.0, gather_ms or 0.0], device=device, dtype=torch.float64)
    if world > 1:
        import torch.distributed as dist

        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms_max, gather_ms = float(t[0]), float(t[1]), float(t[2])
    rows_per_step = 2 * B * world
    value = rows_per_step * args.steps / (ms / 1e3)

    # ---- live per-kernel timing (CUDA events around every launch) for the roofline of the dominant kernel ----
    roofline, breakdown = None, None
    peaks = measured_peaks()
    if rank == 0 and not args.no_prof:
        ops.prof_enable(True, device)
        step_device()
        prof = ops.prof_summary()
        ops.prof_enable(False, device)
        tot = sum(v[1] for v in prof.values())
        breakdown = {k: {"launches": v[0], "ms": round(v[1], 3), "share": round(v[1] / tot, 4)} for k, v in
                     sorted(prof.items(), key=lambda kv: -kv[1][1])}
        chunk = min(ldm.cell_chunk, B)
        chunks = [min(chunk, B - c0) for c0 in range(0, B, chunk)]   # the last chunk of a step may be ragged
        gemm = {k: v for k, v in prof.items() if kernel_flops(k, 1, 1) is not None}
        top = max(gemm.items(), key=lambda kv: kv[1][1])
        name, (cnt, tms) = top
        # 3 forwards per cell (CFG) x 16 tokens; every chunk launches the kernel the same number of times
        n_evals = {"euler": 1, "heun2": 2, "midpoint": 2}.get(args.method, 1) * (NUM_STEPS - 1)
        fl = sum(kernel_flops(name, 3 * c * 16, 1 + c, n_evals) for c in chunks) / len(chunks)   # mean algorithmic FLOPs per launch
        ach = fl / (tms / cnt * 1e-3) / 1e12
        peak = peaks["bf16_tflops_sustained"]
        traffic = None
        tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.isfile(tp):
            tj = json.load(open(tp))
            if tj.get("ode_chunk_cells") == chunk:
                traffic = tj["per_launch_bytes"].get(name)
        roofline = {"bound": "tensor", "kernel": name, "achieved": round(ach, 1), "peak": peak, "unit": "TFLOP/s",
                    "frac": round(ach / peak, 4), "traffic": traffic, "peak_source": peaks["source"] + " (sustained cuBLAS bf16)",
                    "avg_launch_us": round(1e3 * tms / cnt, 2), "flops_per_launch": fl}
        # the decode-stage kernels against their own bounds (algorithmic work of SURVEY 8(d): 23 104 FLOP per (row, gene) token for the MCAB + head,
        # 8 B per (row, gene) for the NB finalisation: logits in, counts out)
        secondary = []
        rows_step = 2 * B
        if "mcab_decode_tc" in prof:
            c2, t2 = prof["mcab_decode_tc"]
            a2 = rows_step * G * 23104 / (t2 * 1e-3) / 1e12
            secondary.append({"bound": "tensor", "kernel": "mcab_decode_tc (mma.sync bf16)", "achieved": round(a2, 1), "peak": peak, "unit": "TFLOP/s",
                              "frac": round(a2 / peak, 4), "avg_launch_us": round(1e3 * t2 / c2, 2)})
        if "nb_finalize" in prof:
            c3, t3 = prof["nb_finalize"]
            a3 = rows_step * G * 8 / (t3 * 1e-3) / 1e9
            secondary.append({"bound": "hbm", "kernel": "nb_finalize", "achieved": round(a3, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                              "frac": round(a3 / peaks["hbm_gbs"], 4), "avg_launch_us": round(1e3 * t3 / c3, 2),
                              "note": "bound by the Philox / NB-inversion arithmetic, not by HBM (ncu: issue slots 78 %)"})
        roofline["secondary"] = secondary

    cpu_baseline = None
    if rank == 0 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        stepc = cpu_sample(args.cpu_batch, dcfg, vcfg, threads)
        t0 = time.perf_counter()
        stepc()
        dtc = time.perf_counter() - t0
        cpu_baseline = {"value": 2 * args.cpu_batch / dtc, "unit": "cells/s", "cores": threads, "kind": reference_kind(),
                        "sample": f"1 x sample() of {args.cpu_batch} cells ({2 * args.cpu_batch} rows): same model, 49-eval Euler + CFG + decode + NB draw, fp32, "
                                  + ("the reference's own modules (oracle/_ref)" if reference_kind() == "reference" else "oracle port"),
                        "seconds": round(dtc, 2)}

    gpu_eager = None
    if rank == 0 and not args.no_gpu_eager:
        prev = torch.get_float32_matmul_precision()
        try:
            torch.set_float32_matmul_precision("high")
            stepg = gpu_eager_sample(args.eager_batch, dcfg, vcfg, device)
            stepg()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); stepg(); b.record(); b.synchronize()
            gpu_eager = {"value": 2 * args.eager_batch / (a.elapsed_time(b) / 1e3), "unit": "cells/s", "kind": reference_kind(),
                         "sample": f"1 x sample() of {args.eager_batch} cells: the reference modules ({reference_kind()}) run eagerly on the same GPU "
                                   "(PyTorch library kernels, float32 matmul precision 'high' as the reference's inference script)"}
        except Exception as e:  # a comparator must never take the bench line down
            gpu_eager = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
        finally:
            torch.set_float32_matmul_precision(prev)
            torch.cuda.empty_cache()

    if rank == 0:
        evals = {"euler": NUM_STEPS - 1, "heun2": 2 * (NUM_STEPS - 1), "midpoint": 2 * (NUM_STEPS - 1)}.get(args.method)
        if evals is None:   # adaptive: function evaluations of the last solve
            evals = int(getattr(ldm.transport_sampler, "last_nfe", 0))
        fl = algorithmic_flops_per_row(G, evals)
        line = {
            "metric": METRIC, "value": value, "unit": "cells/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": workload_name(args.dataset, dcfg, vcfg, args.method),
                       "cells_per_step_per_gpu": B, "rows_per_step": rows_per_step, "ode_chunk_cells": min(ldm.cell_chunk, B),
                       "l2": "256 MB flush buffer written between timed steps", "parallelism": f"dist.sample_sharded: global batch of {Bg} cells split contiguously over {world} GPU(s), no collective in the step (outputs stay rank-local)",
                       "algorithmic_gflop_per_row": round(fl["row_cfg"] / 1e9, 3), "dit_evaluations_per_solve": evals},
            "clocks": clocks,
            "gpu_launches": int(launches),
            "model_tflops": round(value * fl["row_cfg"] / 1e12, 1),
        }
        if e2e_ms is not None:
            h2d = sum(v.numel() * v.element_size() for v in labels_h.values()) + genes_row_h.numel() * 8
            d2h = counts_h.numel() * 4 + z_h.numel() * 4
            line["e2e"] = {"value": rows_per_step * args.steps / (e2e_ms_max / 1e3), "unit": "cells/s", "h2d_bytes_per_step": h2d,
                           "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms_max / args.steps}
            if e2e_csr_ms is not None:   # rank-0 time; informational (the contract's `e2e` keeps the reference's dense return)
                line["e2e_csr"] = {"value": 2 * B * args.steps / (e2e_csr_ms / 1e3) * world, "unit": "cells/s", "d2h_bytes_per_step": csr_bytes[0],
                                   "note": "LatentDiffusion.sample_csr: CSR built on the device, pageable D2H of indptr/indices/data"}
        if world > 1:
            line["gathered"] = {"value": rows_per_step * args.steps / (gather_ms / 1e3), "unit": "cells/s", "ms_per_step": gather_ms / args.steps,
                                "allgather_bytes_per_rank": rows_per_step * (G + 256) * 4,
                                "note": "same step through dist.sample_sharded(gather=True): NCCL all_gather of counts (2B x G fp32) and z to every rank"}
        if roofline:
            line["roofline"] = roofline
            line["kernel_breakdown"] = breakdown
            dec_ms = sum(v["ms"] for k, v in breakdown.items() if k in ("mcab_decode_tc", "mcab_decode", "nb_finalize", "dec_latent", "qside"))
            ode_ms = sum(v["ms"] for k, v in breakdown.items()) - dec_ms
            line["stages"] = {"ode_only_cells_per_s": round(2 * B / (ode_ms / 1e3), 1), "decode_only_cells_per_s": round(2 * B / (dec_ms / 1e3), 1),
                              "note": "per GPU, from the per-kernel CUDA-event times of one profiled step (rows per second of kernel time)"}
        if cpu_baseline:
            line["cpu_baseline"] = cpu_baseline
        if gpu_eager:
            line["gpu_eager_baseline"] = gpu_eager
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


if __name__ == "__main__":
    main()
