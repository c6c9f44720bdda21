"""Worker of tests/test_gpu_train_ddp.py (launched with torch.distributed.run, one rank per GPU, NCCL).

Every rank trains the same DiT replica on its shard of a fixed global batch; after the bucketed all-reduce + AdamW the
weights must equal those of one process that saw the whole batch (mean-of-means = global mean for equal shards)."""

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
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)
    from scldm_b200 import synthetic
    from scldm_b200.config import DiTConfig
    from scldm_b200.nnets import DiT
    from scldm_b200.training import DiTTrainer
    from scldm_b200.transport import create_transport

    cfg = DiTConfig(class_vocab_sizes={"clusters": 14}, n_layer=4)
    per = 16
    Bg = per * world
    z = synthetic.randn("ddp.z", (Bg, 16, 16)).to(dev)
    x0 = synthetic.randn("ddp.x0", (Bg, 16, 16)).to(dev)
    t = torch.rand(Bg, generator=torch.Generator().manual_seed(11)).to(dev)
    lab = synthetic.randint("ddp.lab", 14, (Bg,)).to(dev)
    transport = create_transport("Linear", "velocity")

    def fresh(pg_on):
        m = DiT(**cfg.kwargs())
        m.load_state_dict(synthetic.dit_state_dict(cfg, 1234))
        m = m.to(dev).eval()   # eval: no label dropout, so both runs see the same labels
        tr = DiTTrainer(m, lr=1e-3, max_grad_norm=10.0, n_buckets=3)
        if not pg_on:
            tr.world = 1
        return m, tr

    sl = slice(rank * per, (rank + 1) * per)

    def grads_of(tr, zz, ll, tt, xx):
        B = zz.shape[0]
        te = tt.view(-1, 1, 1)
        v = tr.forward(te * zz + (1 - te) * xx, tt, tr.cls_rows({"clusters": ll}, B))
        diff = v - (zz - xx)
        tr.backward(diff * (2.0 / (B * 256)))
        tr.allreduce_grads()
        torch.cuda.synchronize()
        return tr.grad / tr.world        # the optimizer applies this 1 / world (grad_scale)

    _, tr = fresh(True)
    g_ddp = grads_of(tr, z[sl], lab[sl], t[sl], x0[sl]).clone()
    out = {"rank": rank, "world": world}
    if rank == 0:
        _, tr1 = fresh(False)
        g_one = grads_of(tr1, z, lab, t, x0)
        out["grad_rel_l2_vs_single_process"] = float((g_ddp - g_one).double().norm() / g_one.double().norm())
        out["buckets"] = tr.bucket_bounds
    losses = []
    for _ in range(3):
        losses.append(float(tr.fm_step(z[sl], {"clusters": lab[sl]}, transport, t=t[sl], x0=x0[sl])))
    torch.cuda.synchronize()
    # every rank must hold bit-identical weights after the all-reduced, clipped steps
    ref = tr.flat.clone()
    dist.broadcast(ref, src=0)
    out["max_abs_diff_across_ranks"] = float((tr.flat - ref).abs().max())
    out["losses"] = losses
    gathered = [None] * world
    dist.all_gather_object(gathered, out)
    if rank == 0:
        print("DDP_RESULT " + json.dumps(gathered))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
