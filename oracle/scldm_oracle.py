"""ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.

CPU restatement (plain PyTorch, fp32 by default, fp64 on request) of the scLDM generation
hot path, written function by function from the reference's algorithm with the reference
file:line each function follows.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import this module, and only as
the checker / the CPU baseline -- never as the thing shipped.

Pinning: the reference's own tests hold NO golden vectors for this path (SURVEY.md §4,
§8c), so the restatement is pinned against *outputs of the reference itself run in the dev
container*: `oracle/make_golden.py` imports the unmodified reference modules from
`/root/reference/src/scldm` (namespace stub, see `oracle/ref_loader.py`), runs them on
seeded weights/inputs and commits the vectors under `tests/golden/`;
`tests/test_oracle_golden.py` checks this file against them on every CPU run.

Third-party arithmetic that is not under /root/reference is restated from its published
algorithm and flagged as such:
  * torchdiffeq.odeint fixed-grid solvers (unpinned transitive dependency; call site
    `transport/integrators.py:4,111`): euler / midpoint / heun2 one-step formulas.
  * scvi.distributions.NegativeBinomial.sample (scvi-tools>=1.2, `pyproject.toml:41`; call
    sites `vae.py:87`, `models.py:819`): Gamma-Poisson mixture
    counts ~ Poisson(Gamma(concentration=theta, rate=theta/mu)).

Weights are passed as a flat `state_dict` with the reference's key names.
"""

from __future__ import annotations

import math

import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------
# small pieces
# ----------------------------------------------------------------------------------------


def layer_norm(x, weight=None, bias=None, eps=1e-8):
    """nn.LayerNorm over the last dim, biased variance (torch semantics)."""
    mean = x.mean(dim=-1, keepdim=True)
    var = ((x - mean) ** 2).mean(dim=-1, keepdim=True)
    y = (x - mean) / torch.sqrt(var + eps)
    if weight is not None:
        y = y * weight
    if bias is not None:
        y = y + bias
    return y


def linear(x, sd, name):
    """nn.Linear with (out,in) weight; bias optional."""
    y = x @ sd[name + ".weight"].T
    b = sd.get(name + ".bias")
    return y if b is None else y + b


def swiglu_mlp(x, sd, p):
    """`MLP.forward` reference `layers.py:173-174`: c_proj(silu(w1 x) * (w2 x)), no biases."""
    return (F.silu(x @ sd[p + "w1.weight"].T) * (x @ sd[p + "w2.weight"].T)) @ sd[p + "c_proj.weight"].T


def attention(q, k, v):
    """`flex_attention(q,k,v, block_mask=None, score_mod=None)` eager math path:
    softmax(q k^T / sqrt(head_dim)) v, no mask (reference `layers.py:153,260`)."""
    scale = 1.0 / math.sqrt(q.shape[-1])
    s = (q @ k.transpose(-1, -2)) * scale
    return torch.softmax(s, dim=-1) @ v


def self_attention(x, sd, p, n_head):
    """`SelfAttention.forward` reference `layers.py:143-158`: q,k,v = split(c_attn(x)) (q first)."""
    B, S, D = x.shape
    qkv = linear(x, sd, p + "c_attn")
    q, k, v = qkv.split(D, dim=2)
    hd = D // n_head
    q = q.view(B, S, n_head, hd).transpose(1, 2)
    k = k.view(B, S, n_head, hd).transpose(1, 2)
    v = v.view(B, S, n_head, hd).transpose(1, 2)
    y = attention(q, k, v).transpose(1, 2).contiguous().view(B, S, D)
    return linear(y, sd, p + "c_proj")


def cross_attention(x, q, sd, p, n_head):
    """`CrossAttention.forward` reference `layers.py:248-264`: k,v = split(c_attn(x)) (k first)."""
    B, S, _ = x.shape
    _, M, D = q.shape
    k, v = linear(x, sd, p + "c_attn").split(D, dim=-1)
    qq = linear(q, sd, p + "c_attn_q")
    hd = D // n_head
    k = k.view(B, S, n_head, hd).transpose(1, 2)
    v = v.view(B, S, n_head, hd).transpose(1, 2)
    qq = qq.view(B, M, n_head, hd).transpose(1, 2)
    y = attention(qq, k, v).transpose(1, 2).contiguous().view(B, M, D)
    return linear(y, sd, p + "c_proj")


# ----------------------------------------------------------------------------------------
# DiT  (reference nnets.py:216-492, layers.py:177-226, 339-401)
# ----------------------------------------------------------------------------------------


def timestep_embedding(t, dim=256, max_period=10000):
    """`TimestepEmbedder.timestep_embedding` reference `layers.py:351-360`: [cos | sin], t unscaled."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half).to(t.device)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1).to(t.dtype if t.dtype == torch.float64 else torch.float32)


def t_embedder(t, sd):
    """`TimestepEmbedder.forward` reference `layers.py:362-364`: Linear -> SiLU -> Linear."""
    h = timestep_embedding(t).to(sd["t_embedder.mlp.0.weight"].dtype)
    h = F.silu(linear(h, sd, "t_embedder.mlp.0"))
    return linear(h, sd, "t_embedder.mlp.2")


def condition_embedding(labels: dict, sd, class_vocab_sizes: dict, batch_size: int, device):
    """Sum over *every* class table (sorted names) of emb[label] where a class that is absent from
    `labels` contributes its null row (index = vocab size).

    Restates the eval-mode result of `_get_mutually_exclusive_condition_embedding`
    (`nnets.py:389-426`) when at most one class is passed (which is how `forward_with_cfg`
    calls it, `nnets.py:374`, making the `torch.randint` class pick deterministic) and of
    `_get_joint_condition_embedding` (`nnets.py:428-456`) when all classes are passed.
    """
    total = None
    for name in sorted(class_vocab_sizes.keys()):
        null = class_vocab_sizes[name]
        if name in labels:
            idx = labels[name].long()
        else:
            idx = torch.full((batch_size,), null, dtype=torch.long, device=device)
        e = sd[f"class_embeddings.{name}.weight"][idx]
        total = e if total is None else total + e
    if total is None:
        return None
    return total.unsqueeze(1)


def dit_block(x, c, sd, p, n_head, eps):
    """adaLN `Block.forward` reference `layers.py:208-221`.

    chunks c0..c5 of Linear(SiLU(c)); the reference *calls* modulate(LN, scale_attn, shift_attn)
    where the second chunk is named scale: effective h = LN(x)*(1+c0)+c1  (SURVEY quirk 1).
    """
    mod = linear(F.silu(c), sd, p + "adaln_modulation.1")
    c0, c1, c2, c3, c4, c5 = mod.chunk(6, dim=-1)
    h = layer_norm(x, eps=eps) * (1 + c0) + c1
    x = x + c2 * self_attention(h, sd, p + "attn.", n_head)
    h = layer_norm(x, eps=eps) * (1 + c3) + c4
    x = x + c5 * swiglu_mlp(h, sd, p + "mlp.")
    return x


def dit_final_layer(x, c, sd, eps):
    """`FinalLayerDit.forward` reference `layers.py:397-401`: (shift, scale) = chunk2;
    LN(x)*(1+scale)+shift -- multiplicative is chunk 1 here (opposite of the blocks)."""
    shift, scale = linear(F.silu(c), sd, "final_layer.adaln_modulation.1").chunk(2, dim=-1)
    h = layer_norm(x, eps=eps) * (1 + scale) + shift
    return linear(h, sd, "final_layer.linear")


def dit_forward(x, t, labels: dict, sd, cfg):
    """`DiT.forward` (eval, force_drop_ids=False) reference `nnets.py:273-297`."""
    c = t_embedder(t, sd).unsqueeze(1)
    ce = condition_embedding(labels, sd, cfg.class_vocab_sizes, x.shape[0], x.device)
    if ce is not None:
        c = c + ce
    h = linear(x, sd, "input_proj") + sd["pos_embed"]
    for i in range(cfg.n_layer):
        h = dit_block(h, c, sd, f"blocks.{i}.", cfg.n_head, cfg.layernorm_eps)
    return dit_final_layer(h, c, sd, cfg.layernorm_eps)


def dit_forward_with_cfg(x, t, labels: dict | None, cfg_scale: dict | None, sd, cfg):
    """`DiT.forward_with_cfg` reference `nnets.py:336-378`.

    All 2B rows get the unconditional prediction; rows [B,2B) are guided:
      mutually_exclusive: v = v_u + sum_c w_c (f(x, only class c) - v_u)
      joint:              v = v_u + mean(w) (f(x, all classes) - v_u)
    """
    n = x.shape[0]
    half = n // 2
    uncond = dit_forward(x, t, {}, sd, cfg)
    u_half, g_base = uncond[:half], uncond[half:]
    guided = g_base.clone()
    if labels is not None and cfg_scale is not None:
        xh, th = x[half:], t[half:]
        if cfg.condition_strategy == "joint":
            pred = dit_forward(xh, th, {k: v[half:] for k, v in labels.items()}, sd, cfg)
            avg = sum(cfg_scale.values()) / len(cfg_scale)
            guided = guided + avg * (pred - g_base)
        else:
            for name, scale in cfg_scale.items():
                pred = dit_forward(xh, th, {name: labels[name][half:]}, sd, cfg)
                guided = guided + scale * (pred - g_base)
    return torch.cat([u_half, guided], dim=0)


# ----------------------------------------------------------------------------------------
# ODE sampler (reference transport/transport.py:324-369, integrators.py:78-112; torchdiffeq restated)
# ----------------------------------------------------------------------------------------


def ode_time_grid(num_steps: int, t0: float = 0.0, t1: float = 1.0):
    """`ode.__init__` reference `integrators.py:95`: th.linspace(t0, t1, num_steps).
    For Linear path + velocity prediction `create_transport` forces eps=0 so (t0,t1)=(0,1)
    (`transport/__init__.py:55-57`, `transport.py:69-95`)."""
    return torch.linspace(t0, t1, num_steps)


def odeint_fixed(fn, x, t_grid, method="euler"):
    """RESTATEMENT of torchdiffeq's fixed-grid solvers on the grid itself (no sub-stepping):
    num_steps grid points => num_steps-1 steps. Returns all states (T, ...)."""
    out = [x]
    for k in range(len(t_grid) - 1):
        t0, t1 = t_grid[k], t_grid[k + 1]
        dt = t1 - t0
        if method == "euler":
            x = x + dt * fn(t0, x)
        elif method == "midpoint":
            k1 = fn(t0, x)
            x = x + dt * fn(t0 + 0.5 * dt, x + 0.5 * dt * k1)
        elif method == "heun2":
            k1 = fn(t0, x)
            k2 = fn(t1, x + dt * k1)
            x = x + dt * 0.5 * (k1 + k2)
        else:
            raise NotImplementedError(method)
        out.append(x)
    return torch.stack(out)


def sample_ode(x, model_fn, num_steps=50, method="euler"):
    """`Sampler.sample_ode(...)(x, model, **kw)` + `ode.sample` reference `integrators.py:100-112`:
    scalar t is broadcast to a (rows,) vector before each model call; velocity model => drift = model."""
    grid = ode_time_grid(num_steps).to(x.dtype)

    def fn(t, xx):
        tv = torch.ones(xx.shape[0], dtype=xx.dtype, device=xx.device) * t
        return model_fn(xx, tv)

    return odeint_fixed(fn, x, grid, method)


def fm_training_losses(x1, t, x0, model_fn):
    """`Transport.training_losses` for Linear path + velocity prediction, reference `transport.py:110-150` with
    `ICPlan.plan` (`path.py:129-151`): x_t = t x1 + (1-t) x0, u_t = x1 - x0, loss = mean over (tokens, channels) of
    (model(x_t, t) - u_t)^2 per cell.  t and x0 are passed in (the reference draws them with the global RNG)."""
    te = t.view(-1, *([1] * (x1.dim() - 1)))
    xt = te * x1 + (1 - te) * x0
    ut = x1 - x0
    out = model_fn(xt, t)
    return {"loss": ((out - ut) ** 2).flatten(1).mean(1), "pred": out}


# ----------------------------------------------------------------------------------------
# VAE decoder + NB head (reference nnets.py:147-208, layers.py:267-330, stochastic_layers.py:76-116)
# ----------------------------------------------------------------------------------------


def vae_block(x, sd, p, n_head, eps):
    """non-adaLN `Block.forward` reference `layers.py:222-226` (affine LN)."""
    x = x + self_attention(layer_norm(x, sd[p + "ln_1.weight"], sd[p + "ln_1.bias"], eps), sd, p + "attn.", n_head)
    x = x + swiglu_mlp(layer_norm(x, sd[p + "ln_2.weight"], sd[p + "ln_2.bias"], eps), sd, p + "mlp.")
    return x


def mcab(x, q, sd, p, n_head, eps):
    """`CrossAttentionBlock.forward` (use_adaln=False) reference `layers.py:325-330`:
    x = q + attn(LN1(x), LN1q(q));  x = x + MLP(LN2(x))   -- residual is the raw query (quirk 2)."""
    a = cross_attention(
        layer_norm(x, sd[p + "ln_1.weight"], sd[p + "ln_1.bias"], eps),
        layer_norm(q, sd[p + "ln_1q.weight"], sd[p + "ln_1q.bias"], eps),
        sd, p + "attn.", n_head,
    )
    h = q + a
    return h + swiglu_mlp(layer_norm(h, sd[p + "ln_2.weight"], sd[p + "ln_2.bias"], eps), sd, p + "mlp.")


def decoder_latents(z, sd, cfg):
    """`Decoder.forward` up to the cross attention, reference `nnets.py:203-205`:
    LN(no affine) -> Linear(L->E) -> n_layer Blocks."""
    x = layer_norm(z, eps=cfg.layernorm_eps)
    x = linear(x, sd, "decoder.decoder_latent_input.1")
    for i in range(cfg.n_layer):
        x = vae_block(x, sd, f"decoder.decoder_layers.{i}.", cfg.n_head, cfg.layernorm_eps)
    return x


def decoder_forward(z, genes, sd, cfg):
    """`TransformerVAE.decode` front half, reference `vae.py:78-81` + `nnets.py:200-208`:
    q = shared gene embedding lookup; MCAB unpools 16 latents to G gene tokens."""
    x = decoder_latents(z, sd, cfg)
    q = sd["input_layer.gene_embedding.weight"][genes.long()]
    return mcab(x, q, sd, "decoder.decoder_cross_attention.", cfg.n_head_cross, cfg.layernorm_eps)


def nb_head(h, genes, library_size, sd, cfg, temperature=1.0):
    """`NegativeBinomialTransformerLayer.forward` reference `stochastic_layers.py:102-116`:
    mu = softmax over genes(logit / t) * library_size; theta = exp(theta_emb[gene]) (shared) or
    exp(second output channel) (unshared)."""
    out = linear(h, sd, "decoder_head.params")
    if cfg.shared_theta:
        logit = out.squeeze(-1)
        theta = torch.exp(sd["decoder_head.theta.weight"][genes.long()].squeeze(-1))
    else:
        logit, log_theta = out[..., 0], out[..., 1]
        theta = torch.exp(log_theta)
    mu = torch.softmax(logit / temperature, dim=1) * library_size
    return mu, theta


def vae_decode(z, genes, library_size, sd, cfg):
    """`TransformerVAE.decode` reference `vae.py:71-87` -> (mu, theta) of the NB."""
    return nb_head(decoder_forward(z, genes, sd, cfg), genes, library_size, sd, cfg)


def nb_sample(mu, theta, generator=None):
    """RESTATEMENT of scvi `NegativeBinomial.sample`: Poisson(Gamma(theta, rate=theta/mu)),
    gamma draw clamped to <= 1e8."""
    gamma = torch._standard_gamma(theta.expand_as(mu).contiguous(), generator=generator) * (mu / theta)
    return torch.poisson(torch.clamp(gamma, max=1e8), generator=generator)


# ----------------------------------------------------------------------------------------
# VAE encoder (reference layers.py:97-118, nnets.py:81-144)
# ----------------------------------------------------------------------------------------


def count_transform(counts, agg_func: str = "log1p"):
    """The multiplicative count transforms of `InputTransformerVAE` (reference `layers.py:28-44`, `PROJ_FUNC`)."""
    if agg_func == "log1p":
        return torch.log1p(counts)
    if agg_func == "log1pzero":
        return torch.where(counts == 0, torch.tensor(-1.0), torch.log1p(counts))
    if agg_func == "anscombe":
        return torch.asinh(torch.sqrt(counts + 1.0))
    if agg_func == "sqrt":
        return torch.sqrt(counts + 1.0)
    raise NotImplementedError(agg_func)


def vae_encode(counts_subset, genes_subset, sd, cfg):
    """`TransformerVAE.encode` reference `vae.py:58-69`: emb[gene]*log1p(count) (`layers.py:28-31`) ->
    MCAB pool with inducing points (no key masking, quirk 3) -> +pos_embed -> blocks ->
    Linear(E->L) -> LN(no affine)."""
    emb = sd["input_layer.gene_embedding.weight"][genes_subset.long()]
    x = emb * count_transform(counts_subset, getattr(cfg, "agg_func", "log1p")).unsqueeze(-1)
    B = x.shape[0]
    q = sd["encoder.ca_layer.inducing_points"].unsqueeze(0).expand(B, -1, -1)
    h = mcab(x, q, sd, "encoder.ca_layer.", cfg.n_head_cross, cfg.layernorm_eps)
    if "encoder.pos_embed" in sd:
        h = h + sd["encoder.pos_embed"]
    for i in range(cfg.n_layer):
        h = vae_block(h, sd, f"encoder.encoder_layers.{i}.", cfg.n_head, cfg.layernorm_eps)
    h = linear(h, sd, "encoder.encoder_latent_input.0")
    return layer_norm(h, eps=cfg.layernorm_eps)


# ----------------------------------------------------------------------------------------
# LatentDiffusion.sample (reference models.py:766-819, restated: the Lightning base is not importable)
# ----------------------------------------------------------------------------------------


def sample_log_size_factors(labels: dict | None, mu_tbl: dict | None, sd_tbl: dict | None, batch_size: int,
                            eps_normal: torch.Tensor, key: str | None = None):
    """Independent path of `_sample_log_size_factors` reference `models.py:552-597`:
    l_i = mu[key][label_i] + sd[key][label_i] * eps_i  (Normal(loc, scale).sample() with the
    standard-normal draw supplied by the caller so both sides share it); zeros when tables are missing."""
    out = torch.zeros(batch_size, dtype=eps_normal.dtype)
    if labels is None or mu_tbl is None or sd_tbl is None:
        return out
    if key is None:
        inter = sorted(set(labels) & set(mu_tbl) & set(sd_tbl))
        if not inter:
            return out
        key = inter[0]
    for i in range(batch_size):
        c = int(labels[key][i])
        m, s = mu_tbl[key].get(c), sd_tbl[key].get(c)
        if m is None or s is None:
            continue
        out[i] = m + s * eps_normal[i]
    return out


def sample_joint_log_size_factors(labels: dict, components: list, joint_idx_2_classes: dict, mu_vec: dict, sd_vec: dict,
                                  batch_size: int, eps_normal: torch.Tensor):
    """Joint path of `_sample_log_size_factors` reference `models.py:510-550`: key "{i}_{j}" of the component labels ->
    `joint_idx_2_classes` -> class index -> N(mu, sd); missing key or statistics leave 0."""
    out = torch.zeros(batch_size, dtype=eps_normal.dtype)
    for b in range(batch_size):
        key = "_".join(str(int(labels[k][b])) for k in components)
        if key not in joint_idx_2_classes:
            continue
        c = joint_idx_2_classes[key]
        m, s = mu_vec.get(c), sd_vec.get(c)
        if m is None or s is None:
            continue
        out[b] = m + s * eps_normal[b]
    return out


def latent_diffusion_sample(z0, labels, guidance_weight, genes, log_size_factors, dit_sd, dit_cfg, vae_sd, vae_cfg,
                            num_steps=50, method="euler"):
    """`LatentDiffusion.sample` reference `models.py:766-819` with the noise z0 (B,M,L) and the
    log size factors given (so both sides share them).  Returns (mu, theta, z_final) for the 2B rows:
    [0,B) unconditional, [B,2B) guided.  (The reference then draws counts = NB(mu,theta).sample().)"""
    z_cfg = torch.cat([z0, z0], dim=0)
    labels_cfg = {k: torch.cat([v, v], dim=0) for k, v in (labels or {}).items()}

    def model_fn(x, t):
        return dit_forward_with_cfg(x, t, labels_cfg, guidance_weight, dit_sd, dit_cfg)

    z_final = sample_ode(z_cfg, model_fn, num_steps=num_steps, method=method)[-1]
    genes2 = torch.cat([genes, genes], dim=0)
    lib = torch.exp(log_size_factors).view(-1, 1)
    lib2 = torch.cat([lib, lib], dim=0)
    mu, theta = vae_decode(z_final, genes2, lib2, vae_sd, vae_cfg)
    return mu, theta, z_final


# ----------------------------------------------------------------------------------------
# data formats either side of the path: numpy restatements, test infrastructure only
# ----------------------------------------------------------------------------------------


def tokenize_cells_expressed(counts, gene_idx_row, genes_seq_len: int, mask_idx: int = 0):
    """`tokenize_cells(..., sample_genes="expressed")` (`src/scldm/datamodule.py:708-731`): per cell the expressed genes
    (count > 0) packed left in gene order, padded with the mask token / zero counts; raises when a cell expresses more
    than `genes_seq_len` genes.  counts (N, G) float, gene_idx_row (G,) int64 -> dict as the reference returns."""
    import numpy as np

    counts = np.asarray(counts)
    gene_idx = np.tile(np.asarray(gene_idx_row), (len(counts), 1))
    library_size = counts.sum(1, keepdims=True)
    N, _ = counts.shape
    expressed = counts > 0
    if (expressed.sum(axis=1) > genes_seq_len).any():
        raise ValueError("genes_seq_len is smaller than number of expressed genes")
    pos_order = expressed.cumsum(axis=1) - 1
    genes_out = np.full((N, genes_seq_len), mask_idx, dtype=gene_idx.dtype)
    counts_out = np.zeros((N, genes_seq_len), dtype=counts.dtype)
    ii, jj = np.where(expressed)
    pp = pos_order[expressed]
    genes_out[ii, pp] = gene_idx[ii, jj]
    counts_out[ii, pp] = counts[ii, jj]
    return {"genes": gene_idx, "counts": counts, "genes_subset": genes_out, "counts_subset": counts_out, "library_size": library_size}


def counts_to_csr(counts):
    """`sparse.csr_matrix(counts.numpy())` as used by `process_generation_output` (`src/scldm/_utils.py:186-200`)."""
    from scipy import sparse

    m = sparse.csr_matrix(counts)
    return m.indptr, m.indices, m.data


def log_nb_positive(x, mu, theta, eps: float = 1e-8):
    """NB log-likelihood per entry (`src/scldm/distributions.py:6-42`)."""
    log_theta_mu_eps = torch.log(theta + mu + eps)
    return (theta * (torch.log(theta + eps) - log_theta_mu_eps) + x * (torch.log(mu + eps) - log_theta_mu_eps)
            + torch.lgamma(x + theta) - torch.lgamma(theta) - torch.lgamma(x + 1))


def vae_forward_loss(counts, genes, library_size, counts_subset, genes_subset, sd, cfg):
    """`TransformerVAE.forward` (`src/scldm/vae.py:29-56`) followed by `VAE.loss` for the NB head
    (`src/scldm/models.py:233-247`): returns (mu, theta, h_z, per-cell NLL, llh = mean over cells)."""
    h_z = vae_encode(counts_subset, genes_subset, sd, cfg)
    mu, theta = vae_decode(h_z, genes, library_size, sd, cfg)
    per_cell = (-log_nb_positive(counts, mu, theta)).sum(dim=1)
    return mu, theta, h_z, per_cell, per_cell.mean()
