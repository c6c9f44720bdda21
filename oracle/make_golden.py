"""ORACLE — TEST INFRASTRUCTURE ONLY.  Mints `tests/golden/*.npz` from the REAL reference.

Run in the dev container (needs /root/reference):   python -m oracle.make_golden

Every case builds the unmodified reference modules (`oracle/ref_loader.py`), loads the
deterministic synthetic weights of `scldm_b200/synthetic.py` with `strict=True` (which also
proves the key/shape contract), runs the reference in fp32 on CPU and stores inputs+outputs.
Weights are NOT stored: they are regenerated from (seed, tensor name) on both sides.
"""

from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import ref_loader  # noqa: E402
from scldm_b200 import synthetic  # noqa: E402
from scldm_b200.config import DiTConfig, VAEConfig  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
WEIGHT_SEED = 1234


def golden_cases() -> dict:
    """Case table shared with the tests (configs + input recipes are derived from this)."""
    return {
        "dit_me1": dict(cfg=DiTConfig(class_vocab_sizes={"clusters": 14}), B=4, scales={"clusters": 2.0}),
        "dit_me2": dict(cfg=DiTConfig(class_vocab_sizes={"cell_type": 5, "tissue": 3}, n_layer=2), B=3,
                        scales={"cell_type": 1.5, "tissue": 0.5}),
        "dit_joint": dict(cfg=DiTConfig(class_vocab_sizes={"cell_line": 4, "gene": 7}, n_layer=2,
                                        condition_strategy="joint"), B=3, scales={"cell_line": 1.0, "gene": 3.0}),
    }


def dit_inputs(name: str, cfg: DiTConfig, B: int):
    x = synthetic.randn(name + ".x", (2 * B, cfg.seq_len, cfg.n_embed_input))
    t = torch.from_numpy(np.linspace(0.05, 0.95, 2 * B).astype(np.float32))
    labels = {k: synthetic.randint(name + ".label." + k, v, (2 * B,)) for k, v in cfg.class_vocab_sizes.items()}
    return x, t, labels


def vae_inputs(name: str, cfg: VAEConfig, B: int, S: int):
    z = synthetic.randn(name + ".z", (B, cfg.n_inducing_points, cfg.n_embed_latent))
    genes = torch.arange(1, cfg.n_genes + 1, dtype=torch.int64).unsqueeze(0).repeat(B, 1)
    lib = torch.exp(8.0 + 0.3 * synthetic.randn(name + ".lib", (B, 1)))
    # "expressed"-mode encoder inputs (reference datamodule.py:708-731): distinct ids packed left, zero padded
    rng = np.random.default_rng(99)
    gs = np.zeros((B, S), dtype=np.int64)
    cs = np.zeros((B, S), dtype=np.float32)
    for b in range(B):
        n = int(rng.integers(S // 3, S))
        gs[b, :n] = rng.choice(np.arange(1, cfg.n_genes + 1), size=n, replace=False)
        cs[b, :n] = 1 + rng.poisson(2.0, size=n)
    return z, genes, lib, torch.from_numpy(cs), torch.from_numpy(gs)


@torch.no_grad()
def main() -> None:
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.manual_seed(0)
    ref = ref_loader.load_reference()
    keys: dict[str, dict] = {}

    # ---- DiT forward / forward_with_cfg -------------------------------------------------
    for name, case in golden_cases().items():
        cfg, B = case["cfg"], case["B"]
        sd = synthetic.dit_state_dict(cfg, WEIGHT_SEED)
        model = ref_loader.build_reference_dit(cfg, sd)
        keys[name] = {k: list(v.shape) for k, v in model.state_dict().items()}
        x, t, labels = dit_inputs(name, cfg, B)
        # (a DiT without class tables cannot run in the reference: `_get_condition_embedding`
        #  indexes `condition.values()` unconditionally, nnets.py:381 -- so every case has classes)
        if cfg.condition_strategy == "joint":
            out_fwd = model.forward(x, t, labels, force_drop_ids=False)
        else:
            first = sorted(labels)[0]
            out_fwd = model.forward(x, t, {first: labels[first]}, force_drop_ids=False)
        arrays = dict(x=x.numpy(), t=t.numpy())
        for k, v in labels.items():
            arrays["label." + k] = v.numpy()
        arrays["out_forward"] = out_fwd.numpy()
        arrays["out_cfg"] = model.forward_with_cfg(x, t, labels, case["scales"]).numpy()
        arrays["out_cfg_none"] = model.forward_with_cfg(x, t, None, None).numpy()
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **arrays)
        print(name, {k: v.shape for k, v in arrays.items()})

    # ---- ODE sampler through the reference's Sampler / ode classes -----------------------
    case = golden_cases()["dit_me1"]
    cfg = case["cfg"]
    sd = synthetic.dit_state_dict(cfg, WEIGHT_SEED)
    model = ref_loader.build_reference_dit(cfg, sd)
    transport = ref.transport.create_transport(path_type="Linear", prediction="velocity", loss_weight="velocity",
                                               train_eps=1e-5, sample_eps=1e-5)
    sampler = ref.transport.Sampler(transport)
    B = 2
    z0 = synthetic.randn("ode.z0", (B, cfg.seq_len, cfg.n_embed_input))
    lab = {"clusters": synthetic.randint("ode.label", 14, (B,))}
    z_cfg = torch.cat([z0, z0])
    lab_cfg = {k: torch.cat([v, v]) for k, v in lab.items()}
    arrays = dict(z0=z0.numpy(), label=lab["clusters"].numpy())
    for method, steps, w in (("euler", 50, 2.0), ("euler", 50, 1.0), ("heun2", 10, 2.0), ("midpoint", 10, 2.0)):
        fn = sampler.sample_ode(sampling_method=method, num_steps=steps)
        model_fn = lambda x, t, **kw: model.forward_with_cfg(x, t, **kw, cfg_scale={"clusters": w})  # noqa: E731
        traj = fn(z_cfg, model_fn, condition=lab_cfg)
        arrays[f"z_{method}_{steps}_w{w}"] = traj[-1].numpy()
        print("ode", method, steps, w, traj.shape, float(traj[-1].abs().mean()))
    t0t1 = transport.check_interval(transport.train_eps, transport.sample_eps, sde=False, eval=True)
    arrays["t0t1"] = np.asarray(t0t1, dtype=np.float64)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "ode_me1.npz"), **arrays)

    # ---- flow-matching loss through the reference's Transport.training_losses (eval mode, as in shared_step) ----
    B = 6
    x1 = synthetic.randn("fm.x1", (B, cfg.seq_len, cfg.n_embed_input))
    labf = {"clusters": synthetic.randint("fm.label", 14, (B,))}
    torch.manual_seed(77)
    x0_ref = torch.randn_like(x1)            # same draw order as Transport.sample (transport.py:103-107)
    t_ref = torch.rand((B,))
    torch.manual_seed(77)
    terms = transport.training_losses(model, x1, {"condition": labf})
    np.savez_compressed(os.path.join(GOLDEN_DIR, "fm_loss_me1.npz"), x1=x1.numpy(), label=labf["clusters"].numpy(), x0=x0_ref.numpy(),
                        t=t_ref.numpy(), loss=terms["loss"].numpy(), pred=terms["pred"].numpy())
    print("fm loss", terms["loss"].tolist())

    # ---- VAE decode / encode ----------------------------------------------------------------
    for name, G, B, S in (("vae_small", 1500, 3, 400), ("vae_dentate", 17002, 2, 600)):
        vcfg = VAEConfig(n_genes=G)
        vsd = synthetic.vae_state_dict(vcfg, WEIGHT_SEED)
        vae = ref_loader.build_reference_vae(vcfg, vsd)
        keys[name] = {k: list(v.shape) for k, v in vae.state_dict().items()}
        z, genes, lib, cs, gs = vae_inputs(name, vcfg, B, S)
        nb = vae.decode(z, genes, lib)
        h = vae.decoder(z, vae.input_layer.gene_embedding(genes))
        z_enc = vae.encode(None, None, cs, gs)
        arrays = dict(z=z.numpy(), lib=lib.numpy(), counts_subset=cs.numpy(), genes_subset=gs.numpy(),
                      mu=nb.mu.numpy(), theta=nb.theta[0].numpy(), z_enc=z_enc.numpy(),
                      h_first64=h[:, :64].numpy())
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **arrays)
        print(name, {k: v.shape for k, v in arrays.items()}, "mu.sum/lib", (nb.mu.sum(1) / lib[:, 0]).tolist())

    # ---- full sample(): reference pieces composed as LatentDiffusion.sample composes them ----
    cfg = golden_cases()["dit_me1"]["cfg"]
    vcfg = VAEConfig(n_genes=1500)
    model = ref_loader.build_reference_dit(cfg, synthetic.dit_state_dict(cfg, WEIGHT_SEED))
    vae = ref_loader.build_reference_vae(vcfg, synthetic.vae_state_dict(vcfg, WEIGHT_SEED))
    B = 2
    z0 = synthetic.randn("sample.z0", (B, cfg.seq_len, cfg.n_embed_input))
    lab = {"clusters": synthetic.randint("sample.label", 14, (B,))}
    lsf = 8.0 + 0.3 * synthetic.randn("sample.lsf", (B,))
    genes = torch.arange(1, vcfg.n_genes + 1, dtype=torch.int64).unsqueeze(0).repeat(B, 1)
    fn = sampler.sample_ode(sampling_method="euler", num_steps=50)
    w = {"clusters": 2.0}
    model_fn = lambda x, t, **kw: model.forward_with_cfg(x, t, **kw, cfg_scale=w)  # noqa: E731
    z_fin = fn(torch.cat([z0, z0]), model_fn, condition={k: torch.cat([v, v]) for k, v in lab.items()})[-1]
    lib = torch.exp(lsf).view(-1, 1)
    nb = vae.decode(z_fin, torch.cat([genes, genes]), torch.cat([lib, lib]))
    np.savez_compressed(os.path.join(GOLDEN_DIR, "sample_me1.npz"), z0=z0.numpy(), label=lab["clusters"].numpy(),
                        log_size_factors=lsf.numpy(), z_final=z_fin.numpy(), mu=nb.mu.numpy(), theta=nb.theta[0].numpy())
    print("sample", z_fin.shape, nb.mu.shape)

    with open(os.path.join(GOLDEN_DIR, "state_dict_keys.json"), "w") as f:
        json.dump(keys, f, indent=0, sort_keys=True)


@torch.no_grad()
def nb_loss_golden() -> None:
    """`log_nb_positive` and `TransformerVAE.forward` + NB loss from the reference itself (distributions.py:6-42, vae.py:29-56,
    models.py:233-247) on a small VAE: tests/golden/vae_loss_small.npz."""
    ref = ref_loader.load_reference()
    import scldm.distributions as ref_dist

    cfg, B, S = VAEConfig(n_genes=1500), 3, 400
    sd = synthetic.vae_state_dict(cfg, WEIGHT_SEED)
    vae = ref_loader.build_reference_vae(cfg, sd)
    _, genes, lib, cs, gs = vae_inputs("vae_small", cfg, B, S)
    # dense counts consistent with the subset tokens (gene id g sits in column g-1), plus a few large values for the lgamma terms
    counts = torch.zeros(B, cfg.n_genes)
    for i in range(B):
        m = gs[i] > 0
        counts[i, gs[i][m] - 1] = cs[i][m]
    counts[0, :5] = torch.tensor([0.0, 1.0, 17.0, 250.0, 4000.0])
    lib = counts.sum(1, keepdim=True)
    params, h_z = vae.forward(counts, genes, lib, cs, gs)
    ll = ref_dist.log_nb_positive(counts, params["mu"], params["theta"])
    per_cell = (-ll).sum(dim=1)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "vae_loss_small.npz"), counts=counts.numpy(), lib=lib.numpy(), counts_subset=cs.numpy(),
                        genes_subset=gs.numpy(), mu=params["mu"].numpy(), theta=params["theta"].numpy(), h_z=h_z.numpy(),
                        log_nb=ll.numpy(), per_cell=per_cell.numpy(), llh=per_cell.mean().numpy())
    print("vae_loss_small", float(per_cell.mean()), per_cell.tolist())


@torch.no_grad()
def unshared_theta_golden() -> None:
    """`TransformerVAE.decode` with an unshared-theta NB head (`shared_theta=False`: `params` is Linear(E->2) and theta =
    exp(second channel), stochastic_layers.py:91-98,106-113) from the reference: tests/golden/vae_unshared_theta.npz."""
    cfg, B, S = VAEConfig(n_genes=1500, shared_theta=False), 3, 400
    sd = synthetic.vae_state_dict(cfg, WEIGHT_SEED)
    vae = ref_loader.build_reference_vae(cfg, sd)
    z, genes, lib, _, _ = vae_inputs("vae_unshared", cfg, B, S)
    d = vae.decode(z, genes, lib)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "vae_unshared_theta.npz"), z=z.numpy(), lib=lib.numpy(), mu=d.mu.numpy(), theta=d.theta.numpy())
    print("vae_unshared_theta", d.mu.shape, d.theta.shape, float(d.theta.min()), float(d.theta.max()))


@torch.no_grad()
def label_dropout_golden() -> None:
    """Training-mode CFG label dropout of the reference (`nnets.py:389-456`): the summed class embedding the reference DiT
    builds in train mode under a fixed CPU seed, for the mutually-exclusive 2-class and the joint model:
    tests/golden/dit_label_dropout.npz."""
    arrays = {}
    for name in ("dit_me2", "dit_joint"):
        case = golden_cases()[name]
        cfg = case["cfg"]
        sd = synthetic.dit_state_dict(cfg, WEIGHT_SEED)
        model = ref_loader.build_reference_dit(cfg, sd).train()
        labels = {k: synthetic.randint(f"drop.{name}.{k}", v, (64,)) for k, v in cfg.class_vocab_sizes.items()}
        for rep in range(3):
            torch.manual_seed(1000 + rep)
            emb = model._get_condition_embedding(labels, force_drop_ids=True)
            arrays[f"{name}.emb{rep}"] = emb.squeeze(1).numpy()
        for k, v in labels.items():
            arrays[f"{name}.label.{k}"] = v.numpy()
    np.savez_compressed(os.path.join(GOLDEN_DIR, "dit_label_dropout.npz"), **arrays)
    print("dit_label_dropout", {k: v.shape for k, v in arrays.items()})


TRAIN_GOLDEN_FULL = ("final_layer.linear.weight", "final_layer.linear.bias", "blocks.0.attn.c_attn.bias", "blocks.1.attn.c_proj.bias",
                     "class_embeddings.clusters.weight", "input_proj.weight", "t_embedder.mlp.2.bias", "blocks.0.adaln_modulation.1.bias")


def train_step_golden() -> None:
    """One LDM training step of the REFERENCE (`LatentDiffusion.training_step`, models.py:634-666, minus the Lightning shell):
    `Transport.training_losses(DiT.train(), z, {"condition": labels})` under a fixed CPU seed -> `loss.mean().backward()` through
    the unmodified reference modules -> `clip_grad_norm_(10)` (training/default.yaml:15) -> `torch.optim.AdamW(lr=5e-4)`
    (ldm_base.yaml:36-40).  Stores the RNG draws (x0, t, label-dropout mask, re-derived by replaying the reference's sequence of
    RNG calls), the loss, every gradient's norm, a few gradients in full, strided slices of the rest, and the same for the
    updated weights: tests/golden/train_step_me1.npz."""
    ref = ref_loader.load_reference()
    # torch's flex_attention has no CPU backward ("FlexAttention does not support backward on CPU"), so this fixture is minted on the
    # GPU box from the staged reference modules (oracle/build_ref.py):  SCLDM_GOLDEN_DEVICE=cuda SCLDM_GOLDEN_DIR=gpurun_out/golden
    dev = torch.device(os.environ.get("SCLDM_GOLDEN_DEVICE", "cuda" if torch.cuda.is_available() else "cpu"))
    out_dir = os.environ.get("SCLDM_GOLDEN_DIR", GOLDEN_DIR)
    os.makedirs(out_dir, exist_ok=True)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.set_float32_matmul_precision("highest")
    cfg = DiTConfig(class_vocab_sizes={"clusters": 14}, n_layer=2)
    sd = synthetic.dit_state_dict(cfg, WEIGHT_SEED)
    model = ref_loader.build_reference_dit(cfg, sd).to(dev).train()
    for p in model.parameters():
        p.requires_grad_(True)
    model.pos_embed.requires_grad_(False)
    B = 8
    z = synthetic.randn("train.z", (B, cfg.seq_len, cfg.n_embed_input)).to(dev)
    lab = {"clusters": synthetic.randint("train.lab", 14, (B,)).to(dev)}
    transport = ref.transport.create_transport(path_type="Linear", prediction="velocity", loss_weight="velocity", train_eps=1e-5, sample_eps=1e-5)
    # replay of the reference's RNG calls: Transport.sample (randn_like on the data's device, rand on the CPU: transport.py:104-106),
    # then DiT's randint + rand on the labels' device (nnets.py:395,402)
    torch.manual_seed(4242)
    x0 = torch.randn_like(z)
    t = torch.rand((B,))
    torch.randint(0, 1, (), device=dev)
    drop = torch.rand(B, device=dev) < cfg.cfg_dropout_prob
    torch.manual_seed(4242)
    terms = transport.training_losses(model, z, {"condition": lab})
    loss = terms["loss"].mean()
    loss.backward()
    c = lambda a: a.detach().cpu().numpy().copy()  # noqa: E731
    arrays = dict(z=c(z), label=c(lab["clusters"]), x0=c(x0), t=c(t), drop=c(drop), loss=np.float32(loss.item()),
                  loss_per_cell=c(terms["loss"]), pred=c(terms["pred"]), device=np.array(str(dev)))
    params = dict(model.named_parameters())
    names = [n for n, p in params.items() if p.requires_grad]
    arrays["names"] = np.array(names)
    arrays["grad_norms"] = np.array([float(params[n].grad.norm()) for n in names], dtype=np.float64)
    for n in names:
        g = params[n].grad
        arrays["grad." + n] = c(g) if n in TRAIN_GOLDEN_FULL else c(g.reshape(-1)[::97])
    total = torch.nn.utils.clip_grad_norm_([p for p in model.parameters() if p.requires_grad], 10.0)
    arrays["total_norm"] = np.float64(float(total))
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=5e-4, weight_decay=0.0)
    opt.step()
    for n in names:
        w = params[n].detach()
        arrays["new." + n] = c(w) if n in TRAIN_GOLDEN_FULL else c(w.reshape(-1)[::97])
    np.savez_compressed(os.path.join(out_dir, "train_step_me1.npz"), **arrays)
    print("train_step_me1 loss", float(loss), "total grad norm", float(total), "dropped", int(drop.sum()))



VAE_TRAIN_FULL = ("decoder_head.params.weight", "decoder_head.params.bias", "encoder.ca_layer.inducing_points", "encoder.ca_layer.attn.c_attn_q.weight",
                  "decoder.decoder_cross_attention.mlp.w1.weight", "decoder.decoder_cross_attention.attn.c_attn_q.weight",
                  "encoder.encoder_latent_input.0.weight", "decoder.decoder_latent_input.1.weight", "decoder.decoder_layers.1.attn.c_attn.weight",
                  "encoder.encoder_layers.0.ln_1.weight", "encoder.ca_layer.ln_1.bias")


def vae_train_inputs(cfg: VAEConfig, B: int, S: int):
    """Seeded batch of the VAE training step: "expressed"-mode encoder tokens + the dense counts they come from (+ a few large counts
    for the lgamma / digamma terms), library size = row sum (datamodule.py:708-731)."""
    _, genes, _, cs, gs = vae_inputs("vae_train", cfg, B, S)
    counts = torch.zeros(B, cfg.n_genes)
    for i in range(B):
        m = gs[i] > 0
        counts[i, gs[i][m] - 1] = cs[i][m]
    counts[0, :5] = torch.tensor([0.0, 1.0, 17.0, 250.0, 40.0])
    lib = counts.sum(1, keepdim=True)
    return counts, genes, lib, cs, gs


def vae_train_step_golden() -> None:
    """One VAE training step of the REFERENCE (`VAE.training_step`, models.py:249-287, minus the Lightning shell): unmodified
    `TransformerVAE.forward` -> `-log_nb_positive(...).sum(1).mean()` -> autograd -> `clip_grad_norm_(10)` (training/default.yaml:15)
    -> `AdamWLegacy(lr=1e-3, weight_decay=0)` (vae_base.yaml:56-60; the class itself needs no third-party package).  Stores loss, every
    gradient's norm, a few gradients in full, strided slices of the rest, and the same for the updated weights:
    tests/golden/vae_train_step.npz.  flex_attention has no CPU backward -> minted on the GPU box like train_step_me1."""
    ref = ref_loader.load_reference()
    import scldm.distributions as ref_dist
    dev = torch.device(os.environ.get("SCLDM_GOLDEN_DEVICE", "cuda" if torch.cuda.is_available() else "cpu"))
    out_dir = os.environ.get("SCLDM_GOLDEN_DIR", GOLDEN_DIR)
    os.makedirs(out_dir, exist_ok=True)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.set_float32_matmul_precision("highest")
    cfg, B, S = VAEConfig(n_genes=1500, n_layer=2), 3, 400
    sd = synthetic.vae_state_dict(cfg, WEIGHT_SEED)
    vae = ref_loader.build_reference_vae(cfg, sd).to(dev).train()
    counts, genes, lib, cs, gs = [a.to(dev) for a in vae_train_inputs(cfg, B, S)]
    params, h_z = vae.forward(counts, genes, lib, cs, gs)
    per_cell = (-ref_dist.log_nb_positive(counts, params["mu"], params["theta"])).sum(dim=1)
    loss = per_cell.mean()
    loss.backward()
    c = lambda a: a.detach().cpu().numpy().copy()  # noqa: E731
    arrays = dict(loss=np.float32(loss.item()), per_cell=c(per_cell), h_z=c(h_z), mu=c(params["mu"]), device=np.array(str(dev)))
    named = dict(vae.named_parameters())
    names = [n for n, p in named.items() if p.requires_grad]
    arrays["names"] = np.array(names)
    arrays["grad_norms"] = np.array([float(named[n].grad.norm()) for n in names], dtype=np.float64)
    for n in names:
        g = named[n].grad
        arrays["grad." + n] = c(g) if n in VAE_TRAIN_FULL else c(g.reshape(-1)[::53])
    total = torch.nn.utils.clip_grad_norm_([named[n] for n in names], 10.0)
    arrays["total_norm"] = np.float64(float(total))
    try:
        import scldm.optimizers as ref_opt
        opt = ref_opt.AdamWLegacy([named[n] for n in names], lr=1e-3, weight_decay=0.0)
        arrays["optimizer"] = np.array("scldm.optimizers.AdamWLegacy")
    except Exception as e:  # noqa: BLE001
        print("AdamWLegacy not importable (", e, ") -> torch.optim.AdamW (same update for amsgrad=False, caution=False)")
        opt = torch.optim.AdamW([named[n] for n in names], lr=1e-3, weight_decay=0.0)
        arrays["optimizer"] = np.array("torch.optim.AdamW")
    opt.step()
    for n in names:
        w = named[n].detach()
        arrays["new." + n] = c(w) if n in VAE_TRAIN_FULL else c(w.reshape(-1)[::53])
    np.savez_compressed(os.path.join(out_dir, "vae_train_step.npz"), **arrays)
    print("vae_train_step loss", float(loss), "total grad norm", float(total))


@torch.no_grad()
def vae256_golden() -> None:
    """Census-scale VAE width (n_embed = 256: 8 heads, 4 cross heads, SwiGLU hidden 684) through the REFERENCE modules:
    decode + encode at G = 1500 (B = 3, S = 400) and decode at the census vocabulary G = 36 130 (B = 1):
    tests/golden/vae256_small.npz, vae256_census.npz."""
    for name, G, B, S in (("vae256_small", 1500, 3, 400), ("vae256_census", 36130, 1, 0)):
        vcfg = VAEConfig(n_genes=G, n_embed=256)
        vsd = synthetic.vae_state_dict(vcfg, WEIGHT_SEED)
        vae = ref_loader.build_reference_vae(vcfg, vsd)
        z, genes, lib, cs, gs = vae_inputs(name, vcfg, B, max(S, 8))
        nb = vae.decode(z, genes, lib)
        arrays = dict(z=z.numpy(), lib=lib.numpy(), mu=nb.mu.numpy(), theta=nb.theta[0].numpy())
        if S > 0:
            arrays.update(counts_subset=cs.numpy(), genes_subset=gs.numpy(), z_enc=vae.encode(None, None, cs, gs).numpy())
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **arrays)
        print(name, {k: v.shape for k, v in arrays.items()}, "mu.sum/lib", (nb.mu.sum(1) / lib[:, 0]).tolist())


@torch.no_grad()
def sde_golden() -> None:
    """`Sampler.sample_sde` of the REFERENCE (`transport.py:269-322`, `integrators.py:7-75`) on the me1 DiT with classifier-free
    guidance: Euler-Maruyama and Heun, diffusion_form="sigma" (the default "SBDM" is infinite at t0 = 0 for the eps = 0 that
    `create_transport` forces), 10 grid points, last_step="Mean".  The Brownian increments the reference drew (torch.randn on the
    CPU) are recorded by wrapping torch.randn for the duration of the call: tests/golden/sde_me1.npz."""
    ref = ref_loader.load_reference()
    case = golden_cases()["dit_me1"]
    cfg = case["cfg"]
    model = ref_loader.build_reference_dit(cfg, synthetic.dit_state_dict(cfg, WEIGHT_SEED))
    transport = ref.transport.create_transport(path_type="Linear", prediction="velocity", loss_weight="velocity", train_eps=1e-5, sample_eps=1e-5)
    sampler = ref.transport.Sampler(transport)
    B = 2
    z0 = synthetic.randn("sde.z0", (B, cfg.seq_len, cfg.n_embed_input))
    lab = synthetic.randint("sde.label", 14, (B,))
    w = {"clusters": 2.0}
    model_fn = lambda x, t, **kw: model.forward_with_cfg(x, t, **kw, cfg_scale=w)  # noqa: E731
    arrays = dict(z0=z0.numpy(), label=lab.numpy())
    for method in ("Euler", "Heun"):
        fn = sampler.sample_sde(sampling_method=method, diffusion_form="sigma", diffusion_norm=1.0, last_step="Mean", last_step_size=0.04, num_steps=10)
        drawn, orig = [], torch.randn
        import scldm.transport.integrators as integ

        def rec(*a, **k):
            t = orig(*a, **k)
            drawn.append(t.clone())
            return t

        integ.th.randn = rec
        try:
            torch.manual_seed(31)
            xs = fn(torch.cat([z0, z0]), model_fn, condition={"clusters": torch.cat([lab, lab])})
        finally:
            integ.th.randn = orig
        arrays[f"{method}.noise"] = torch.stack(drawn).numpy()
        arrays[f"{method}.states"] = torch.stack(xs).numpy()
        print("sde", method, len(drawn), torch.stack(xs).shape, float(xs[-1].abs().mean()))
    np.savez_compressed(os.path.join(GOLDEN_DIR, "sde_me1.npz"), **arrays)


def vae_agg_golden() -> None:
    """The other multiplicative count transforms of `InputTransformerVAE` (reference `layers.py:28-44`: log1pzero, anscombe, sqrt)
    through the reference's own encoder (the shipped yaml uses log1p, which `vae_small` / `vae_dentate` cover)."""
    arrays = {}
    for agg in ("log1pzero", "anscombe", "sqrt"):
        vcfg = VAEConfig(n_genes=1500, agg_func=agg)
        vae = ref_loader.build_reference_vae(vcfg, synthetic.vae_state_dict(vcfg, WEIGHT_SEED))
        _, _, _, cs, gs = vae_inputs("vae_small", vcfg, 3, 400)
        with torch.no_grad():
            arrays["z_enc_" + agg] = vae.encode(None, None, cs, gs).numpy()
    arrays.update(counts_subset=cs.numpy(), genes_subset=gs.numpy())
    np.savez_compressed(os.path.join(GOLDEN_DIR, "vae_agg.npz"), **arrays)
    print("vae_agg", {k: v.shape for k, v in arrays.items()})


if __name__ == "__main__":
    later = {"vae_agg": vae_agg_golden, "nb_loss": nb_loss_golden, "unshared_theta": unshared_theta_golden, "label_dropout": label_dropout_golden, "train_step": train_step_golden, "vae_train_step": vae_train_step_golden, "vae256": vae256_golden, "sde": sde_golden}   # fixtures added after the first set; minted
    if len(sys.argv) > 1 and sys.argv[1] in later:                                 # alone so the others stay byte-identical
        later[sys.argv[1]]()
    else:
        main()
        for fn in later.values():
            fn()
