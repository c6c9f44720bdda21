"""GPU parity of the census-scale VAE (n_embed = 256; SURVEY.md 8d: "benchmark E in {32, 256}") against vectors minted from the
unmodified reference modules (`oracle.make_golden vae256`) and against the oracle at the census shapes.

Every MCAB contraction takes bf16 operands with fp32 accumulation (tcgen05), LayerNorm / softmax / residuals are fp32:
  mu rel-L2 <= 1e-2, |sum_g mu - library| / library <= 1e-4, theta (fp32 exp of a table) <= 1e-5, encoder latents rel-L2 <= 1e-2."""

import os

import numpy as np
import pytest
import torch

from oracle import scldm_oracle as O
from oracle.make_golden import WEIGHT_SEED, vae_inputs
from scldm_b200 import synthetic
from scldm_b200.config import VAEConfig

pytestmark = pytest.mark.gpu
TOL_MU, TOL_Z = 1e-2, 1e-2


def rel_l2(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def make_vae(cfg):
    from scldm_b200.vae import TransformerVAE

    vae = TransformerVAE.from_config(cfg)
    sd = synthetic.vae_state_dict(cfg, WEIGHT_SEED)
    vae.load_state_dict(sd, strict=True)
    return vae.cuda().eval(), sd


@pytest.mark.parametrize("name,G,B,S", [("vae256_small", 1500, 3, 400), ("vae256_census", 36130, 1, 0)])
def test_decode_vs_reference_golden(golden_dir, name, G, B, S):
    g = dict(np.load(os.path.join(golden_dir, name + ".npz")))
    cfg = VAEConfig(n_genes=G, n_embed=256)
    vae, sd = make_vae(cfg)
    z, genes, lib, _, _ = vae_inputs(name, cfg, B, max(S, 8))
    nb = vae.decode(z.cuda(), genes.cuda(), lib.cuda())
    e_mu, e_th = rel_l2(nb.mu, g["mu"]), rel_l2(nb.theta[0], g["theta"])
    tot = (nb.mu.sum(1).cpu() / lib[:, 0] - 1).abs().max().item()
    print(name, f"mu rel-L2 {e_mu:.2e} theta {e_th:.2e} |sum/lib-1| {tot:.2e}")
    assert e_mu < TOL_MU and e_th < 1e-5 and tot < 1e-4, (e_mu, e_th, tot)
    assert nb.mu.shape == (B, G)


def test_encode_vs_reference_golden(golden_dir):
    g = dict(np.load(os.path.join(golden_dir, "vae256_small.npz")))
    cfg = VAEConfig(n_genes=1500, n_embed=256)
    vae, _ = make_vae(cfg)
    z = vae.encode(None, None, torch.from_numpy(g["counts_subset"]).cuda(), torch.from_numpy(g["genes_subset"]).cuda())
    e = rel_l2(z, g["z_enc"])
    print(f"encode E=256: z rel-L2 {e:.2e}")
    assert z.shape == (3, 16, 16) and e < TOL_Z, e


def test_many_cells_ragged_genes_and_chunking():
    """37 cells (not a multiple of 8), G = 1000 (not a multiple of 128), forced into several cell chunks: vs the oracle."""
    from scldm_b200 import ops

    cfg = VAEConfig(n_genes=1000, n_embed=256, n_layer=2)
    vae, sd = make_vae(cfg)
    B = 37
    z = synthetic.randn("v256.z", (B, 16, 16))
    lib = torch.exp(8.0 + 0.3 * synthetic.randn("v256.lib", (B, 1)))
    genes = torch.arange(1, 1001).unsqueeze(0).expand(B, -1)
    with torch.no_grad():
        mu_o, th_o = O.vae_decode(z, genes, lib, sd, cfg)
    packed = vae.packed_decoder()
    mu, theta, counts = ops.vae256_decode(packed, z.cuda(), genes[0].contiguous().cuda(), lib.cuda(), want_mu=True, want_counts=True, seed=3, max_rows=16 * 1024)
    e = rel_l2(mu, mu_o)
    print(f"E=256 many cells: mu {e:.2e}")
    assert e < TOL_MU and rel_l2(theta, th_o[0]) < 1e-5, e
    assert bool((counts >= 0).all()) and bool((counts == counts.round()).all())
    assert abs(float(counts.sum()) / float(mu.sum()) - 1) < 0.02        # 37 000 NB draws: totals agree to a few sd
    mu2, _, _ = ops.vae256_decode(packed, z.cuda(), genes[0].contiguous().cuda(), lib.cuda(), want_mu=True)
    assert torch.equal(mu2, mu)                                          # chunking does not change a cell's result


def test_census_shape_encode_decode_vs_oracle():
    """BASELINE configs[4] shapes: G = 36 130 decode and S = 8 000 encode, two cells, against the fp32 oracle."""
    cfg = VAEConfig(n_genes=36130, n_embed=256)
    vae, sd = make_vae(cfg)
    B, S = 2, 8000
    z, genes, lib, cs, gs = vae_inputs("v256.census", cfg, B, S)
    z_enc = vae.encode(None, None, cs.cuda(), gs.cuda())
    nb = vae.decode(z.cuda(), genes.cuda(), lib.cuda())
    with torch.no_grad():
        z_o = O.vae_encode(cs, gs, sd, cfg)
        mu_o, _ = O.vae_decode(z, genes, lib, sd, cfg)
    e_z, e_mu = rel_l2(z_enc, z_o), rel_l2(nb.mu, mu_o)
    print(f"census E=256: encode z {e_z:.2e}, decode mu {e_mu:.2e}")
    assert e_z < TOL_Z and e_mu < TOL_MU, (e_z, e_mu)
