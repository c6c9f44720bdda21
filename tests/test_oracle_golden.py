"""Pins the oracle restatement (`oracle/scldm_oracle.py`) against golden vectors minted from the
unmodified reference (`oracle/make_golden.py`).  fp32 CPU both sides: tolerance 2e-5 relative-L2
(only op-ordering differences)."""

import json
import os

import numpy as np
import pytest
import torch

from oracle import scldm_oracle as O
from oracle.make_golden import WEIGHT_SEED, dit_inputs, golden_cases, vae_inputs
from scldm_b200 import synthetic
from scldm_b200.config import VAEConfig

TOL = 2e-5


def rel_l2(a, b):
    a, b = torch.as_tensor(a, dtype=torch.float64), torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name + ".npz")))


@pytest.mark.parametrize("name", list(golden_cases().keys()))
def test_dit_forward_and_cfg(golden_dir, name):
    case = golden_cases()[name]
    cfg, B = case["cfg"], case["B"]
    g = load(golden_dir, name)
    sd = synthetic.dit_state_dict(cfg, WEIGHT_SEED)
    x, t, labels = dit_inputs(name, cfg, B)
    assert np.array_equal(x.numpy(), g["x"])  # input recipe is reproducible
    with torch.no_grad():
        if cfg.condition_strategy == "joint":
            fwd = O.dit_forward(x, t, labels, sd, cfg)
        else:
            first = sorted(labels)[0]
            fwd = O.dit_forward(x, t, {first: labels[first]}, sd, cfg)
        assert rel_l2(fwd, g["out_forward"]) < TOL
        out = O.dit_forward_with_cfg(x, t, labels, case["scales"], sd, cfg)
        assert rel_l2(out, g["out_cfg"]) < TOL
        out_none = O.dit_forward_with_cfg(x, t, None, None, sd, cfg)
        assert rel_l2(out_none, g["out_cfg_none"]) < TOL
    # guided half differs from unguided base => conditioning is exercised (not vacuous)
    assert rel_l2(g["out_cfg"], g["out_cfg_none"]) > 1e-2


@pytest.mark.parametrize("method,steps,w", [("euler", 50, 2.0), ("euler", 50, 1.0), ("heun2", 10, 2.0), ("midpoint", 10, 2.0)])
def test_ode_sampler(golden_dir, method, steps, w):
    g = load(golden_dir, "ode_me1")
    assert tuple(g["t0t1"]) == (0.0, 1.0)  # eps forced to 0 for Linear+velocity (transport/__init__.py:55-57)
    cfg = golden_cases()["dit_me1"]["cfg"]
    sd = synthetic.dit_state_dict(cfg, WEIGHT_SEED)
    z0 = torch.from_numpy(g["z0"])
    lab = {"clusters": torch.from_numpy(g["label"])}
    lab2 = {k: torch.cat([v, v]) for k, v in lab.items()}
    with torch.no_grad():
        traj = O.sample_ode(torch.cat([z0, z0]), lambda x, t: O.dit_forward_with_cfg(x, t, lab2, {"clusters": w}, sd, cfg),
                            num_steps=steps, method=method)
    assert traj.shape[0] == steps  # N grid points = N-1 steps (quirk 4)
    assert rel_l2(traj[-1], g[f"z_{method}_{steps}_w{w}"]) < 1e-4


@pytest.mark.parametrize("name,G,B,S", [("vae_small", 1500, 3, 400), ("vae_dentate", 17002, 2, 600)])
def test_vae_decode_encode(golden_dir, name, G, B, S):
    g = load(golden_dir, name)
    cfg = VAEConfig(n_genes=G)
    sd = synthetic.vae_state_dict(cfg, WEIGHT_SEED)
    z, genes, lib, cs, gs = vae_inputs(name, cfg, B, S)
    with torch.no_grad():
        h = O.decoder_forward(z, genes, sd, cfg)
        assert rel_l2(h[:, :64], g["h_first64"]) < TOL
        mu, theta = O.vae_decode(z, genes, lib, sd, cfg)
        assert rel_l2(mu, g["mu"]) < 1e-4
        assert rel_l2(theta[0], g["theta"]) < 1e-6
        assert torch.allclose(mu.sum(1), lib[:, 0], rtol=1e-5)
        z_enc = O.vae_encode(cs, gs, sd, cfg)
        assert rel_l2(z_enc, g["z_enc"]) < 1e-4


@pytest.mark.parametrize("agg", ["log1pzero", "anscombe", "sqrt"])
def test_encode_count_transforms(golden_dir, agg):
    """The other multiplicative count transforms of InputTransformerVAE (reference layers.py:28-44) through the oracle's encoder vs
    the reference-minted vectors (`python -m oracle.make_golden vae_agg`)."""
    g = load(golden_dir, "vae_agg")
    cfg = VAEConfig(n_genes=1500, agg_func=agg)
    sd = synthetic.vae_state_dict(cfg, WEIGHT_SEED)
    with torch.no_grad():
        z_enc = O.vae_encode(torch.from_numpy(g["counts_subset"]), torch.from_numpy(g["genes_subset"]), sd, cfg)
    assert rel_l2(z_enc, g["z_enc_" + agg]) < 1e-4


def test_full_sample(golden_dir):
    g = load(golden_dir, "sample_me1")
    cfg = golden_cases()["dit_me1"]["cfg"]
    vcfg = VAEConfig(n_genes=1500)
    dsd = synthetic.dit_state_dict(cfg, WEIGHT_SEED)
    vsd = synthetic.vae_state_dict(vcfg, WEIGHT_SEED)
    B = 2
    genes = torch.arange(1, vcfg.n_genes + 1, dtype=torch.int64).unsqueeze(0).repeat(B, 1)
    with torch.no_grad():
        mu, theta, zf = O.latent_diffusion_sample(
            torch.from_numpy(g["z0"]), {"clusters": torch.from_numpy(g["label"])}, {"clusters": 2.0}, genes,
            torch.from_numpy(g["log_size_factors"]), dsd, cfg, vsd, vcfg, num_steps=50, method="euler")
    assert rel_l2(zf, g["z_final"]) < 1e-4
    assert rel_l2(mu, g["mu"]) < 1e-3
    assert rel_l2(theta[0], g["theta"]) < 1e-6


def test_state_dict_contract(golden_dir):
    """synthetic specs == the reference modules' own state_dict keys/shapes."""
    with open(os.path.join(golden_dir, "state_dict_keys.json")) as f:
        keys = json.load(f)
    for name, case in golden_cases().items():
        spec = synthetic.dit_state_spec(case["cfg"])
        assert {k: list(s) for k, (s, _) in spec.items()} == keys[name]
    for name, G in (("vae_small", 1500), ("vae_dentate", 17002)):
        spec = synthetic.vae_state_spec(VAEConfig(n_genes=G))
        assert {k: list(s) for k, (s, _) in spec.items()} == keys[name]


def test_nb_sample_moments():
    """restated scvi Gamma-Poisson: mean mu, variance mu + mu^2/theta."""
    gen = torch.Generator().manual_seed(7)
    mu = torch.tensor([[0.3, 2.0, 15.0, 120.0]]).repeat(200000, 1)
    theta = torch.tensor([0.5, 1.5, 3.0, 8.0])
    x = O.nb_sample(mu, theta, generator=gen)
    m, v = x.mean(0), x.var(0)
    exp_v = mu[0] + mu[0] ** 2 / theta
    assert torch.allclose(m, mu[0], rtol=0.03)
    assert torch.allclose(v, exp_v, rtol=0.06)


def test_fm_training_losses(golden_dir):
    """oracle restatement of Transport.training_losses (Linear path, velocity) vs the reference's own Transport."""
    g = load(golden_dir, "fm_loss_me1")
    cfg = golden_cases()["dit_me1"]["cfg"]
    sd = synthetic.dit_state_dict(cfg, WEIGHT_SEED)
    lab = {"clusters": torch.from_numpy(g["label"])}
    with torch.no_grad():
        out = O.fm_training_losses(torch.from_numpy(g["x1"]), torch.from_numpy(g["t"]), torch.from_numpy(g["x0"]),
                                   lambda x, t: O.dit_forward(x, t, lab, sd, cfg))
    assert rel_l2(out["loss"], g["loss"]) < 1e-5 and rel_l2(out["pred"], g["pred"]) < TOL


def test_oracle_tokenizer_and_csr_small_cases():
    """hand-checkable cases for the numpy restatements of the data-format steps either side of the path
    (datamodule.py:708-731, _utils.py:186-200)."""
    import numpy as np

    from oracle import scldm_oracle as O

    counts = np.array([[0, 3, 0, 1, 0], [0, 0, 0, 0, 0], [2, 2, 2, 0, 0]], dtype=np.float32)
    genes = np.array([11, 12, 13, 14, 15], dtype=np.int64)
    t = O.tokenize_cells_expressed(counts, genes, 3, mask_idx=0)
    assert t["genes_subset"].tolist() == [[12, 14, 0], [0, 0, 0], [11, 12, 13]]
    assert t["counts_subset"].tolist() == [[3, 1, 0], [0, 0, 0], [2, 2, 2]]
    assert t["library_size"].ravel().tolist() == [4, 0, 6]
    import pytest

    with pytest.raises(ValueError):
        O.tokenize_cells_expressed(counts, genes, 2)
    indptr, indices, data = O.counts_to_csr(counts)
    assert indptr.tolist() == [0, 2, 2, 5] and indices.tolist() == [1, 3, 0, 1, 2] and data.tolist() == [3, 1, 2, 2, 2]


def test_vae_forward_nb_loss(golden_dir):
    """oracle `log_nb_positive` / `TransformerVAE.forward` + `VAE.loss` vs the reference's own outputs
    (distributions.py:6-42, vae.py:29-56, models.py:233-247)."""
    import numpy as np
    import torch

    from oracle import scldm_oracle as O
    from oracle.make_golden import WEIGHT_SEED
    from scldm_b200 import synthetic
    from scldm_b200.config import VAEConfig

    g = dict(np.load(os.path.join(golden_dir, "vae_loss_small.npz")))
    cfg = VAEConfig(n_genes=1500)
    sd = synthetic.vae_state_dict(cfg, WEIGHT_SEED)
    counts, mu, theta = (torch.from_numpy(g[k]) for k in ("counts", "mu", "theta"))
    ll = O.log_nb_positive(counts, mu, theta)
    assert torch.allclose(ll, torch.from_numpy(g["log_nb"]), rtol=1e-5, atol=1e-5)
    genes = torch.arange(1, cfg.n_genes + 1).unsqueeze(0).repeat(counts.shape[0], 1)
    with torch.no_grad():
        mu_o, th_o, hz_o, per_cell, llh = O.vae_forward_loss(counts, genes, torch.from_numpy(g["lib"]), torch.from_numpy(g["counts_subset"]),
                                                            torch.from_numpy(g["genes_subset"]), sd, cfg)
    assert rel_l2(mu_o, g["mu"]) < TOL and rel_l2(hz_o, g["h_z"]) < TOL and rel_l2(th_o, g["theta"]) < 1e-6
    assert rel_l2(per_cell, g["per_cell"]) < 1e-5 and abs(float(llh) - float(g["llh"])) < 1e-3 * abs(float(g["llh"]))


def test_vae_decode_unshared_theta(golden_dir):
    """oracle decode with the unshared-theta NB head (stochastic_layers.py:91-98,106-113) vs the reference's output."""
    import numpy as np
    import torch

    from oracle import scldm_oracle as O
    from oracle.make_golden import WEIGHT_SEED
    from scldm_b200 import synthetic
    from scldm_b200.config import VAEConfig

    g = dict(np.load(os.path.join(golden_dir, "vae_unshared_theta.npz")))
    cfg = VAEConfig(n_genes=1500, shared_theta=False)
    sd = synthetic.vae_state_dict(cfg, WEIGHT_SEED)
    B = g["z"].shape[0]
    genes = torch.arange(1, cfg.n_genes + 1).unsqueeze(0).repeat(B, 1)
    with torch.no_grad():
        mu, theta = O.vae_decode(torch.from_numpy(g["z"]), genes, torch.from_numpy(g["lib"]), sd, cfg)
    assert theta.shape == (B, cfg.n_genes)
    assert rel_l2(mu, g["mu"]) < TOL and rel_l2(theta, g["theta"]) < TOL


def test_training_step_oracle_matches_reference_autograd(golden_dir):
    """`tests/golden/train_step_me1.npz` was minted on a B200 from the UNMODIFIED reference modules (staged by oracle/build_ref.py):
    `Transport.training_losses(DiT.train())` -> autograd -> clip_grad_norm_(10) -> torch.optim.AdamW(lr=5e-4).  torch's flex_attention
    has no CPU backward, so the reference itself cannot be differentiated in the dev container; the oracle restatement can, and must
    reproduce loss, every gradient and the updated weights (fp32 GPU vs fp32 CPU: 2e-5 relative)."""
    import numpy as np

    from scldm_b200.config import DiTConfig

    g = dict(np.load(os.path.join(golden_dir, "train_step_me1.npz")))
    cfg = DiTConfig(class_vocab_sizes={"clusters": 14}, n_layer=2)
    sd = synthetic.dit_state_dict(cfg, WEIGHT_SEED)
    sdg = {k: v.clone().requires_grad_(k != "pos_embed") for k, v in sd.items()}
    z, x0, t = (torch.from_numpy(g[k]) for k in ("z", "x0", "t"))
    lab = torch.from_numpy(g["label"]).clone()
    lab[torch.from_numpy(g["drop"])] = 14                      # CFG label dropout (nnets.py:402-417): dropped labels -> the null row
    out = O.fm_training_losses(z, t, x0, lambda xt, tt: O.dit_forward(xt, tt, {"clusters": lab}, sdg, cfg))
    loss = out["loss"].mean()
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) < 2e-5 * float(g["loss"])
    assert rel_l2(out["pred"].detach(), g["pred"]) < 2e-5
    names = [str(n) for n in g["names"]]
    for n, gn in zip(names, g["grad_norms"]):
        grad = sdg[n].grad
        assert abs(float(grad.norm()) - gn) < 1e-4 * gn, n
        ref = g["grad." + n]
        mine = grad.numpy() if ref.shape == tuple(grad.shape) else grad.reshape(-1)[::97].numpy()
        assert rel_l2(mine, ref) < 1e-4, n
    params = [sdg[n] for n in names]
    total = torch.nn.utils.clip_grad_norm_(params, 10.0)
    assert abs(float(total) - float(g["total_norm"])) < 1e-4 * float(g["total_norm"])
    torch.optim.AdamW(params, lr=5e-4, weight_decay=0.0).step()
    for n in names:
        ref = g["new." + n]
        w = sdg[n].detach()
        mine = w.numpy() if ref.shape == tuple(w.shape) else w.reshape(-1)[::97].numpy()
        # the first Adam step moves a weight by lr * g / (|g| + eps): where |g| ~ eps = 1e-8 (the k bias, whose gradient is zero up
        # to rounding because it cancels in the softmax) the step is rounding noise of size <= lr, so those elements only get the
        # lr bound; everywhere else the update must agree to 2e-6
        gref = g["grad." + n]
        big = np.abs(gref) > 1e-6
        d = np.abs(mine - ref)
        assert d.max() <= 5.1e-4 and (not big.any() or d[big].max() < 2e-6), n
