"""CPU checks of the VAE training step's host side: the oracle's autograd is pinned to the reference-minted golden
(`tests/golden/vae_train_step.npz`: unmodified reference modules + torch autograd + AdamWLegacy on a B200), and the flat
parameter layout of `VAETrainer` matches the group sizes the kernels assume (include/scldm_b200.h)."""

import os

import numpy as np
import torch

from oracle import scldm_oracle as O
from oracle.make_golden import VAE_TRAIN_FULL, WEIGHT_SEED, vae_train_inputs
from scldm_b200 import synthetic
from scldm_b200.config import VAEConfig


def test_oracle_vae_autograd_matches_reference_golden(golden_dir):
    g = dict(np.load(os.path.join(golden_dir, "vae_train_step.npz")))
    cfg, B, S = VAEConfig(n_genes=1500, n_layer=2), 3, 400
    sd = {k: v.clone().requires_grad_(k != "encoder.pos_embed") for k, v in synthetic.vae_state_dict(cfg, WEIGHT_SEED).items()}
    counts, genes, lib, cs, gs = vae_train_inputs(cfg, B, S)
    mu, theta, h_z, per_cell, llh = O.vae_forward_loss(counts, genes, lib, cs, gs, sd, cfg)
    llh.backward()
    assert abs(float(llh) - float(g["loss"])) / abs(float(g["loss"])) < 2e-5
    assert float((h_z.detach() - torch.from_numpy(g["h_z"])).norm() / torch.from_numpy(g["h_z"]).norm()) < 2e-5
    floor = 1e-5 * float(g["total_norm"])    # decoder_head.params.bias: its gradient is exactly zero in exact arithmetic (softmax shift invariance)
    for name, norm in zip(g["names"].tolist(), g["grad_norms"].tolist()):
        grad = sd[name].grad
        ours = grad if name in VAE_TRAIN_FULL else grad.reshape(-1)[::53]
        ref = torch.from_numpy(g["grad." + name])
        if norm < floor:
            assert float(grad.norm()) < floor, name
            continue
        err = float((ours.double() - ref.double()).norm() / ref.double().norm())
        assert err < 1e-4, (name, err)
        assert abs(float(grad.norm()) - norm) <= 1e-4 * norm + floor
    # clip_grad_norm_(10) + one AdamW step (AdamWLegacy == torch.optim.AdamW for amsgrad=False, caution=False: optimizers.py:72-141)
    params = [sd[n] for n in g["names"].tolist()]
    total = torch.nn.utils.clip_grad_norm_(params, 10.0)
    assert abs(float(total) - float(g["total_norm"])) < 1e-4 * float(g["total_norm"])
    torch.optim.AdamW(params, lr=1e-3, weight_decay=0.0).step()
    noise = {n for n, nrm in zip(g["names"].tolist(), g["grad_norms"].tolist()) if nrm < floor}   # Adam normalises a noise-level gradient to a full-size step
    for name in sorted(set(g["names"].tolist()) - noise):
        w = sd[name].detach()
        ours = w if name in VAE_TRAIN_FULL else w.reshape(-1)[::53]
        assert float((ours - torch.from_numpy(g["new." + name])).abs().max()) < 2e-5, name


def test_vae_flat_layout_matches_kernel_groups():
    from scldm_b200.vae import TransformerVAE
    from scldm_b200.vae_training import BLOCK_SIZE, MCAB_SIZE, flat_layout

    cfg = VAEConfig(n_genes=1500, n_layer=3)
    vae = TransformerVAE.from_config(cfg)
    shapes = {k: tuple(v.shape) for k, v in vae.named_parameters() if v.requires_grad}
    assert "encoder.pos_embed" not in shapes                       # frozen in the reference (nnets.py:103-106)
    names, off, goff, n = flat_layout(cfg.n_layer, shapes)
    assert set(names) == set(shapes) and len(names) == len(set(names))
    assert goff["dec_ca"] - goff["enc_ca"] == MCAB_SIZE == goff["inducing"] - goff["dec_ca"]
    assert goff["dec_blocks"] - goff["enc_blocks"] == cfg.n_layer * BLOCK_SIZE == goff["enc_lat"] - goff["dec_blocks"]
    assert all(o % 4 == 0 for o in off.values()) and n % 4 == 0   # float4 loads of every tensor
    # offsets inside a Block group as vae_train_kernels.cuh lays them out
    b0 = goff["enc_blocks"]
    assert off["encoder.encoder_layers.0.attn.c_attn.weight"] - b0 == 64 and off["encoder.encoder_layers.0.mlp.c_proj.weight"] - b0 == BLOCK_SIZE - 32 * 88
    c0 = goff["dec_ca"]
    assert off["decoder.decoder_cross_attention.attn.c_attn_q.weight"] - c0 == 192 + 2048 and off["decoder.decoder_cross_attention.mlp.w1.weight"] - c0 == 192 + 2048 + 2048


def test_vae_trainer_has_no_cpu_path():
    import pytest

    from scldm_b200.vae import TransformerVAE
    from scldm_b200.vae_training import VAETrainer

    with pytest.raises(RuntimeError, match="CUDA"):
        VAETrainer(TransformerVAE.from_config(VAEConfig(n_genes=100, n_layer=1)))
