"""Worker of tests/test_gpu_vae_train_ddp.py (launched with torch.distributed.run, one rank per GPU, NCCL).

Every rank runs the VAE training step on its shard of a fixed global batch; after the NCCL all-reduce of the flat gradient the
gradient must equal that of one process that saw the whole batch, and the replicas must stay identical after AdamWLegacy."""

import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from oracle.make_golden import vae_train_inputs
    from scldm_b200 import synthetic
    from scldm_b200.config import VAEConfig
    from scldm_b200.vae import TransformerVAE
    from scldm_b200.vae_training import VAETrainer

    cfg = VAEConfig(n_genes=900, n_layer=2)
    per = 4
    counts, genes, lib, cs, gs = [a.to(dev) for a in vae_train_inputs(cfg, per * world, 200)]

    def fresh(dist_on):
        vae = TransformerVAE.from_config(cfg)
        vae.load_state_dict(synthetic.vae_state_dict(cfg, 1234))
        vae = vae.to(dev).train()
        tr = VAETrainer(vae, lr=1e-3, exact=True)
        if not dist_on:
            tr.world = 1
        return vae, tr

    sl = slice(rank * per, (rank + 1) * per)
    _, tr = fresh(True)
    tr.forward_backward(counts[sl], genes[sl], lib[sl], cs[sl], gs[sl], global_batch=per * world)
    tr.allreduce_grads()
    torch.cuda.synchronize()
    g_ddp = tr.grad.clone()
    _, tr1 = fresh(False)
    tr1.forward_backward(counts, genes, lib, cs, gs)
    torch.cuda.synchronize()
    rel = float((g_ddp - tr1.grad).norm() / tr1.grad.norm())

    # a few full steps: replicas must stay bit-identical (same all-reduced gradient, same deterministic clip norm)
    _, tr2 = fresh(True)
    losses = []
    for _ in range(5):
        batch = dict(counts=counts[sl], genes=genes[sl], library_size=lib[sl], counts_subset=cs[sl], genes_subset=gs[sl])
        losses.append(float(tr2.training_step(batch)))
    flat = tr2.flat.clone()
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    diff = max(float((g - gathered[0]).abs().max()) for g in gathered)
    out = {"rank": rank, "grad_rel_l2_vs_single_process": rel, "max_abs_diff_across_ranks": diff, "losses": losses}
    allout = [None] * world
    dist.all_gather_object(allout, out)
    if rank == 0:
        print("DDP_RESULT " + json.dumps(allout))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
