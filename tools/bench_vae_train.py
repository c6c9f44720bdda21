"""Census-vocabulary VAE TRAINING step throughput (BASELINE configs[4]): G = 36 130 genes, S = 8 000 encoder
tokens per cell, per-GPU batch 128 (`experiments/configs/model/ldm_base.yaml:58`).  One step = `VAETrainer.training_step`:
encode -> decode of every gene -> NB loss -> backward -> (NCCL all-reduce of the flat gradient when launched under torchrun) ->
clip + AdamWLegacy.  Prints one JSON line (rank 0).

    python tools/bench_vae_train.py [--cells 128] [--dataset census] [--steps 10] [--warmup 3] [--exact] [--no-eager]
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/bench_vae_train.py

`gpu_eager_baseline`: the reference's own modules (oracle/_ref, unmodified) + torch autograd + its AdamWLegacy on the same GPU at
the reference's training precision (TF32 "high", scripts/train.py:18) on a bounded batch - the library-call comparator."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from scldm_b200 import ops, synthetic
from scldm_b200.config import DATASETS, dataset_configs
from scldm_b200.vae import TransformerVAE
from scldm_b200.vae_training import VAETrainer

ap = argparse.ArgumentParser()
ap.add_argument("--dataset", default="census")
ap.add_argument("--cells", type=int, default=128)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--exact", action="store_true")
ap.add_argument("--no-eager", action="store_true")
ap.add_argument("--no-e2e", action="store_true")
ap.add_argument("--eager-cells", type=int, default=16)
args = ap.parse_args()

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device(f"cuda:{local}")
torch.cuda.set_device(dev)
if world > 1:
    torch.distributed.init_process_group("nccl", device_id=dev)

_, vcfg = dataset_configs(args.dataset)
G, S, B = vcfg.n_genes, DATASETS[args.dataset]["genes_seq_len"], args.cells
vae = TransformerVAE.from_config(vcfg)
sd = synthetic.vae_state_dict(vcfg, 1234)
vae.load_state_dict(sd)
vae = vae.to(dev).train()
trainer = VAETrainer(vae, lr=1e-3, exact=args.exact)


def make_batch(n, seed):
    gen = torch.Generator().manual_seed(seed)
    counts = torch.zeros(n, G)
    for i in range(n):   # "expressed"-mode cells: n_expr ~ U(0.05 G, 0.3 G), counts 1 + Poisson(2)
        k = min(S, int(torch.randint(int(0.05 * G), int(0.3 * G), (1,), generator=gen)))
        idx = torch.randperm(G, generator=gen)[:k]
        counts[i, idx] = 1.0 + torch.poisson(torch.full((k,), 2.0), generator=gen)
    counts = counts.to(dev)
    gene_row = torch.arange(1, G + 1, device=dev)
    tok = ops.tokenize_expressed(counts, gene_row, S)
    return dict(counts=counts, genes=gene_row.unsqueeze(0).expand(n, -1), library_size=tok["library_size"], counts_subset=tok["counts_subset"],
                genes_subset=tok["genes_subset"])


batch = make_batch(B, 7 + rank)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
losses = []
for _ in range(args.warmup):
    losses.append(trainer.training_step(batch))
torch.cuda.synchronize()
if world > 1:
    torch.distributed.barrier()
tot = 0.0
for _ in range(args.steps):
    flush.fill_(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    losses.append(trainer.training_step(batch))
    b.record()
    b.synchronize()
    tot += a.elapsed_time(b)
ms = torch.tensor([tot / args.steps], device=dev)
if world > 1:
    torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
ms = float(ms)
loss_first, loss_last = float(losses[0]), float(losses[-1])

# end to end through the public API from HOST buffers: pinned dense counts -> H2D -> device tokenizer -> training step -> D2H of the loss
e2e_ms = None
if not args.no_e2e:
    counts_host = batch["counts"].cpu().pin_memory()
    gene_row = torch.arange(1, G + 1, device=dev)
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()

    # double-buffered upload, as a DataLoader with pinned memory + non_blocking copies feeds a training loop: while step k runs, the dense
    # counts of step k + 1 go host -> device on a copy stream; every timed step contains exactly one upload and waits for its own inputs
    copy_stream = torch.cuda.Stream(device=dev)
    dev_bufs = [torch.empty_like(batch["counts"]) for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    state = {"k": 0}

    def upload(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])            # the step that last read this buffer has finished
            dev_bufs[slot].copy_(counts_host, non_blocking=True)
            ready[slot].record(copy_stream)

    for ev in consumed:
        ev.record(torch.cuda.current_stream(dev))
    upload(0)

    def e2e_step():
        k = state["k"]
        cur, nxt = k & 1, (k + 1) & 1
        upload(nxt)
        main = torch.cuda.current_stream(dev)
        main.wait_event(ready[cur])
        c = dev_bufs[cur]
        tok = ops.tokenize_expressed(c, gene_row, S)
        loss = trainer.training_step(dict(counts=c, genes=batch["genes"], library_size=tok["library_size"], counts_subset=tok["counts_subset"],
                                          genes_subset=tok["genes_subset"]))
        consumed[cur].record(main)
        loss_host.copy_(loss, non_blocking=True)
        state["k"] = k + 1

    for _ in range(2):
        e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    tot = 0.0
    for _ in range(args.steps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        e2e_step()
        b.record()
        b.synchronize()
        tot += a.elapsed_time(b)
    e = torch.tensor([tot / args.steps], device=dev)
    if world > 1:
        torch.distributed.all_reduce(e, op=torch.distributed.ReduceOp.MAX)
    e2e_ms = float(e)

breakdown = None
if rank == 0:
    saved, trainer.world = trainer.world, 1
    ops.prof_enable(True, dev)
    trainer.training_step(batch)
    prof = ops.prof_summary()
    ops.prof_enable(False, dev)
    trainer.world = saved
    t_all = sum(v[1] for v in prof.values())
    breakdown = {k: {"launches": v[0], "ms": round(v[1], 3), "share": round(v[1] / t_all, 4)} for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])}

eager = None
if rank == 0 and not args.no_eager:
    try:
        from oracle import ref_loader

        ref = ref_loader.load_reference()
        import scldm.distributions as ref_dist
        import scldm.optimizers as ref_opt

        torch.set_float32_matmul_precision("high")
        n = args.eager_cells
        rvae = ref_loader.build_reference_vae(vcfg, sd).to(dev).train()
        opt = ref_opt.AdamWLegacy([p for p in rvae.parameters() if p.requires_grad], lr=1e-3, weight_decay=0.0)
        eb = make_batch(n, 99)
        genes_full = eb["genes"].contiguous()

        def ref_step():
            opt.zero_grad(set_to_none=True)
            params, _ = rvae(eb["counts"], genes_full, eb["library_size"], eb["counts_subset"], eb["genes_subset"])
            loss = (-ref_dist.log_nb_positive(eb["counts"], params["mu"], params["theta"])).sum(dim=1).mean()
            loss.backward()
            torch.nn.utils.clip_grad_norm_(rvae.parameters(), 10.0)
            opt.step()
            return loss

        ref_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 2
        for _ in range(reps):
            ref_step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        eager = {"value": round(n / dt, 1), "unit": "cells/s", "kind": "reference", "ms_per_step": round(dt * 1e3, 1),
                 "sample": f"{reps} x training step of {n} cells: the reference's own modules (oracle/_ref) + torch autograd + AdamWLegacy, eager on the same GPU, "
                           f"float32 matmul precision 'high' (scripts/train.py:18)", "peak_mem_gb": round(torch.cuda.max_memory_allocated(dev) / 2**30, 1)}
    except Exception as e:  # noqa: BLE001
        eager = {"unavailable": f"{type(e).__name__}: {str(e)[:200]}"}

if rank == 0:
    E, H, M = 32, 88, 16
    tok_flop = 2 * (E * E + 3 * E * H) + 4 * M * E          # per decoder gene token, forward: c_proj + SwiGLU + 16-key attention (Q side cached)
    dec_fwd = G * tok_flop
    enc_fwd = S * (4 * E * E + 4 * M * E)
    step_flop = B * 3 * (dec_fwd + enc_fwd)                 # forward + dgrad + wgrad (the recomputed forward of the backward kernel not counted)
    dec_ms = sum(v["ms"] for k, v in (breakdown or {}).items() if k.startswith("vtr_dec_mcab"))
    line = {
        "metric": "VAE training cells/sec (encode + decode + NB loss + backward + clip + AdamWLegacy)", "value": round(B * world / ms * 1e3, 1), "unit": "cells/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "weak",
        "dtype": "tf32x3 (fp32-grade)" if args.exact else "tf32 (decoder MCAB GEMMs) / fp32", "data": "synthetic",
        "config": {"workload": f"{args.dataset}-vocabulary VAE training step: G={G}, S={S}, {B} cells per GPU, E=32, 8+8 layers", "l2": "256 MB flush buffer written between timed steps",
                   "parallelism": f"dp{world}: one NCCL all-reduce of the flat fp32 gradient ({trainer.n_params * 4 / 1e6:.1f} MB)"},
        "loss_first": round(loss_first, 3), "loss_last": round(loss_last, 3),
        "algorithmic_gflop_per_step": round(step_flop / 1e9, 1), "model_tflops": round(step_flop / ms / 1e9, 1),
        "roofline": {"bound": "tensor (mma.sync TF32)", "kernel": "vtr::dec_mcab_train_kernel (forward + backward launches)",
                     "achieved": round(B * 3 * dec_fwd / max(dec_ms, 1e-9) / 1e9, 1), "peak": 1393.8 / 2, "unit": "TFLOP/s",
                     "frac": round(B * 3 * dec_fwd / max(dec_ms, 1e-9) / 1e9 / (1393.8 / 2), 4),
                     "peak_source": "half the measured sustained cuBLAS bf16 rate of MEASURED_PEAKS.json (TF32 tensor-core rate = bf16 / 2); no TF32 measurement on this pool",
                     "traffic": None},
        "e2e": None if e2e_ms is None else {"value": round(B * world / e2e_ms * 1e3, 1), "unit": "cells/s", "ms_per_step": round(e2e_ms, 3),
                                            "h2d_bytes_per_step": B * G * 4, "d2h_bytes_per_step": 4,
                                            "note": "pinned dense counts -> H2D (double-buffered on a copy stream: the upload of step k + 1 overlaps step k; one upload inside every timed step) -> scldm_tokenize_expressed -> VAETrainer.training_step -> D2H of the loss"},
        "gpu_launches": 17 * (args.steps + args.warmup),
        "kernel_breakdown": breakdown, "gpu_eager_baseline": eager,
    }
    print(json.dumps(line))
if world > 1:
    torch.distributed.destroy_process_group()
