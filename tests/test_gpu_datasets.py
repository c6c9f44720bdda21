"""GPU parity at the gene-vocabulary / class-table shapes of every dataset the reference ships a datamodule entry for
(`datamodule/default.yaml:41-137`; BASELINE.json configs[2], [3], and the census VAE of configs[4]): the full
`LatentDiffusion.sample` chain (CFG ODE -> decode -> NB mean) against the oracle on the same weights, noise, labels and
library sizes.  Small batches / few ODE steps keep the CPU oracle at seconds; the kernels and code paths are those of
the full-size run (vocabulary-sized tables, joint vs mutually-exclusive conditioning, G-sized softmax).

Tolerances (<= ~3x the errors measured on B200: latents 2.3e-3, NB means 3.3e-3): latents rel-L2 <= 5e-3 (bf16 tensor-core DiT, fp32 residual
stream), NB mean rel-L2 <= 1e-2 on the bf16 decode path,
|sum_g mu - library| / library <= 1e-4 (fp32 softmax over genes)."""

import pytest
import torch

from oracle import scldm_oracle as O
from oracle.make_golden import WEIGHT_SEED
from scldm_b200 import synthetic
from scldm_b200.config import DATASETS, dataset_configs

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize("dataset", ["hlca", "tabula_muris", "parse1m", "replogle"])
def test_sample_matches_oracle_at_dataset_shape(dataset):
    from scldm_b200.models import LatentDiffusion
    from scldm_b200.nnets import DiT
    from scldm_b200.transport import create_transport
    from scldm_b200.vae import TransformerVAE

    dcfg, vcfg = dataset_configs(dataset)
    dsd, vsd = synthetic.dit_state_dict(dcfg, WEIGHT_SEED), synthetic.vae_state_dict(vcfg, WEIGHT_SEED)
    dit = DiT(**dcfg.kwargs())
    dit.load_state_dict(dsd, strict=True)
    vae = TransformerVAE.from_config(vcfg)
    vae.load_state_dict(vsd, strict=True)
    steps = 6
    ldm = LatentDiffusion(vae.cuda().eval(), dit.cuda().eval(), create_transport("Linear", "velocity"), sampling_method="euler", num_steps=steps)
    B = 5
    lab = {k: synthetic.randint(f"ds.{dataset}.{k}", v, (B,)) for k, v in dcfg.class_vocab_sizes.items()}
    lab = {k: torch.cat([v[:-1], torch.tensor([dcfg.class_vocab_sizes[k] - 1])]) for k, v in lab.items()}   # include the last class id
    w = {k: 1.0 + 0.5 * (i + 1) for i, k in enumerate(sorted(lab))}      # distinct, non-trivial guidance weights
    z0 = synthetic.randn(f"ds.{dataset}.z0", (B, 16, 16))
    lsf = 8.0 + 0.3 * synthetic.randn(f"ds.{dataset}.lsf", (B,))
    genes = torch.arange(1, vcfg.n_genes + 1).unsqueeze(0).repeat(B, 1)
    counts, z, mu = ldm.sample({k: v.cuda() for k, v in lab.items()}, w, B, genes.cuda(), z0=z0.cuda(), log_size_factors=lsf.cuda(),
                               return_mu=True)
    with torch.no_grad():
        mu_o, _, z_o = O.latent_diffusion_sample(z0, lab, w, genes, lsf, dsd, dcfg, vsd, vcfg, num_steps=steps, method="euler")
    e_z, e_mu = rel_l2(z, z_o), rel_l2(mu, mu_o)
    print(f"{dataset}: G={vcfg.n_genes} classes={dcfg.class_vocab_sizes} {dcfg.condition_strategy}: z {e_z:.2e} mu {e_mu:.2e}")
    assert counts.shape == (2 * B, vcfg.n_genes) and bool(torch.isfinite(counts).all()) and bool((counts >= 0).all())
    assert e_z < 5e-3 and e_mu < 1e-2, (e_z, e_mu)
    lib = torch.exp(lsf)
    assert torch.allclose(mu.sum(1).cpu(), torch.cat([lib, lib]), rtol=1e-4)


def test_census_vocabulary_decode_and_encode_match_oracle():
    """census-shaped VAE (G = 36 130, no classes): MCAB decode and MCAB encode at the largest vocabulary the reference names."""
    from scldm_b200.vae import TransformerVAE

    _, vcfg = dataset_configs("census")
    vsd = synthetic.vae_state_dict(vcfg, WEIGHT_SEED)
    vae = TransformerVAE.from_config(vcfg)
    vae.load_state_dict(vsd, strict=True)
    vae = vae.cuda().eval()
    B, S = 3, DATASETS["census"]["genes_seq_len"]
    z = synthetic.randn("ds.census.z", (B, 16, 16))
    lib = torch.exp(8.0 + 0.3 * synthetic.randn("ds.census.lsf", (B, 1)))
    genes = torch.arange(1, vcfg.n_genes + 1).unsqueeze(0).repeat(B, 1)
    dist = vae.decode(z.cuda(), genes.cuda(), lib.cuda())
    with torch.no_grad():
        mu_o, th_o = O.vae_decode(z, genes, lib, vsd, vcfg)
    e_mu, e_th = rel_l2(dist.mu, mu_o), rel_l2(dist.theta, th_o)
    print(f"census decode: mu {e_mu:.2e} theta {e_th:.2e}")
    assert e_mu < 1e-2 and e_th < 1e-5, (e_mu, e_th)
    assert torch.allclose(dist.mu.sum(1).cpu(), lib.reshape(-1), rtol=1e-4)
    # encode: S = 8000 tokens per cell, ragged numbers of expressed genes, mask-padded (gene 0 / count 0) as the tokenizer does
    gs = torch.zeros(B, S, dtype=torch.int64)
    cs = torch.zeros(B, S)
    for i, n in enumerate((S, 2500, 1)):
        perm = synthetic.randint(f"ds.census.perm{i}", vcfg.n_genes, (n,)) + 1
        gs[i, :n] = perm
        cs[i, :n] = 1.0 + synthetic.randint(f"ds.census.cnt{i}", 6, (n,)).float()
    z_enc = vae.encode(None, None, cs.cuda(), gs.cuda())
    z_enc = z_enc[0] if isinstance(z_enc, (tuple, list)) else z_enc
    with torch.no_grad():
        z_o = O.vae_encode(cs, gs, vsd, vcfg)
    e = rel_l2(z_enc, z_o)
    print(f"census encode: z {e:.2e}")
    assert z_enc.shape == (B, 16, 16) and e < 1e-2, e   # tensor-core pooling with bf16 operands; z is LayerNorm-ed (unit variance)


def test_baseline_config0_plain_sampling_no_cfg():
    """BASELINE.json configs[0] at full size: dentate_gyrus-shaped model, batch 64, `sample_ode('euler', num_steps=50)` with the plain
    conditional `DiT.forward` as the model (no classifier-free guidance), then decode.  The fused loop (`FusedForwardModel`: one
    C-ABI call) and the generic host loop over an opaque lambda agree, and both match the oracle run on the CPU."""
    from scldm_b200.nnets import DiT
    from scldm_b200.transport import Sampler, create_transport
    from scldm_b200.transport.transport import FusedForwardModel
    from scldm_b200.vae import TransformerVAE

    dcfg, vcfg = dataset_configs("dentate_gyrus")
    dsd, vsd = synthetic.dit_state_dict(dcfg, WEIGHT_SEED), synthetic.vae_state_dict(vcfg, WEIGHT_SEED)
    dit = DiT(**dcfg.kwargs())
    dit.load_state_dict(dsd, strict=True)
    dit = dit.cuda().eval()
    vae = TransformerVAE.from_config(vcfg)
    vae.load_state_dict(vsd, strict=True)
    vae = vae.cuda().eval()
    B = 64
    z0 = synthetic.randn("c0.z0", (B, 16, 16))
    lab = {"clusters": synthetic.randint("c0.lab", 14, (B,))}
    lib = torch.exp(8.0 + 0.3 * synthetic.randn("c0.lsf", (B, 1)))
    genes = torch.arange(1, vcfg.n_genes + 1).unsqueeze(0).repeat(B, 1)
    fn = Sampler(create_transport("Linear", "velocity")).sample_ode(sampling_method="euler", num_steps=50)
    cond = {"clusters": lab["clusters"].cuda()}
    z_fused = fn(z0.cuda(), FusedForwardModel(dit), condition=cond)[-1]
    z_loop = fn(z0.cuda(), lambda x, t, **kw: dit.forward(x, t, **kw, force_drop_ids=False), condition=cond)[-1]
    mu = vae.decode(z_fused, genes.cuda(), lib.cuda()).mu
    with torch.no_grad():
        z_o = O.sample_ode(z0, lambda x, t: O.dit_forward(x, t, lab, dsd, dcfg), num_steps=50, method="euler")[-1]
        mu_o, _ = O.vae_decode(z_o, genes, lib, vsd, vcfg)
    e_z, e_l, e_mu = rel_l2(z_fused, z_o), rel_l2(z_loop, z_fused), rel_l2(mu, mu_o)
    print(f"configs[0]: z fused-vs-oracle {e_z:.2e}, host-loop-vs-fused {e_l:.2e}, mu {e_mu:.2e}")
    assert e_z < 5e-3 and e_l < 5e-3 and e_mu < 1e-2, (e_z, e_l, e_mu)   # the two loops differ in how the next input projection is evaluated (fp32 kernel vs hi/lo bf16 tensor-core form)
    assert torch.allclose(mu.sum(1).cpu(), lib.reshape(-1), rtol=1e-4)
