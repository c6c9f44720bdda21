"""Worker of tests/test_gpu_dist.py (torch.distributed.run, one rank per GPU, NCCL): `dist.sample_sharded(gather=True)` on 2 ranks
must return, on every rank, exactly the rows a single-process `LatentDiffusion.sample` of the same global batch returns."""

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
    from scldm_b200 import dist as sdist
    from scldm_b200 import synthetic
    from scldm_b200.config import DiTConfig, VAEConfig
    from scldm_b200.models import LatentDiffusion
    from scldm_b200.nnets import DiT
    from scldm_b200.transport import create_transport
    from scldm_b200.vae import TransformerVAE

    dcfg = DiTConfig(class_vocab_sizes={"clusters": 14}, n_layer=2)
    vcfg = VAEConfig(n_genes=1200, n_layer=2)

    def build():
        dit = DiT(**dcfg.kwargs())
        dit.load_state_dict(synthetic.dit_state_dict(dcfg, 1234))
        vae = TransformerVAE.from_config(vcfg)
        vae.load_state_dict(synthetic.vae_state_dict(vcfg, 1234))
        mu_t, sd_t = synthetic.size_factor_tables(dcfg.class_vocab_sizes, 1234)
        return LatentDiffusion(vae.to(dev).eval(), dit.to(dev).eval(), create_transport("Linear", "velocity"), mu_size_factor=mu_t,
                               sd_size_factor=sd_t, sampling_method="euler", num_steps=8, seed=4321)

    B = 37   # ragged split: 19 + 18
    lab = {"clusters": synthetic.randint("dist.lab", 14, (B,)).to(dev)}
    genes = torch.arange(1, vcfg.n_genes + 1, device=dev).unsqueeze(0).expand(B, -1)
    w = {"clusters": 2.0}
    ldm = build()
    counts, z = sdist.sample_sharded(ldm, lab, w, B, genes, gather=True)
    counts2, _ = sdist.sample_sharded(ldm, lab, w, B, genes, gather=True)     # second call: fresh Philox offsets
    ref = build()
    rc, rz = ref.sample(lab, w, B, genes)
    rc2, _ = ref.sample(lab, w, B, genes)
    torch.cuda.synchronize()
    out = {"rank": rank, "counts_equal": bool(torch.equal(counts, rc)), "z_equal": bool(torch.equal(z, rz)), "second_call_equal": bool(torch.equal(counts2, rc2)),
           "calls_differ": bool(not torch.equal(counts, counts2)), "shape": list(counts.shape), "nnz": int((counts > 0).sum())}
    gathered = [None] * world
    dist.all_gather_object(gathered, out)
    if rank == 0:
        print("DIST_RESULT " + json.dumps(gathered))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
