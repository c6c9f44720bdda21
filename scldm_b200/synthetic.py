"""Deterministic synthetic weights and inputs for parity tests and benchmarks.

There is no network, hence no released checkpoint: every test and benchmark uses
random-init weights of the reference architecture.  The reference's own initialiser
(`nnets.py:458-492`) zeroes every adaLN Linear and the final Linear, so a fresh DiT is
identically 0 and any parity test on it is vacuous.  The
generator below therefore gives every tensor non-trivial values.  It is keyed by
(seed, crc32(tensor name)) through numpy's Philox bit generator, so that the dev
container (which can import the real reference to mint golden vectors) and the GPU box
(which cannot) build bit-identical state_dicts.

State-dict key/shape specs follow the reference modules' `state_dict()` (SURVEY.md §8b,
probed; the golden fixture `tests/golden/state_dict_keys.json` pins them).
"""

from __future__ import annotations

import zlib

import numpy as np
import torch

from .config import DiTConfig, VAEConfig


def sincos_pos_embed(embed_dim: int, seq_len: int) -> np.ndarray:
    """1-D sin/cos table, *sin half first* (reference `layers.py:367-385`)."""
    pos = np.arange(seq_len, dtype=np.float32).reshape(-1, 1)
    omega = np.arange(embed_dim // 2, dtype=np.float32)
    omega /= embed_dim / 2.0
    omega = 1.0 / (10000**omega)
    out = pos * omega.reshape(1, -1)
    return np.concatenate([np.sin(out), np.cos(out)], axis=1).astype(np.float32)


def dit_state_spec(cfg: DiTConfig) -> dict[str, tuple[tuple[int, ...], str]]:
    """name -> (shape, init-kind) for `scldm.nnets.DiT.state_dict()` (reference `nnets.py:219-271`)."""
    D, L, H = cfg.n_embed, cfg.n_embed_input, cfg.hidden
    spec: dict[str, tuple[tuple[int, ...], str]] = {}
    spec["pos_embed"] = ((1, cfg.seq_len, D), "sincos")
    n_null = int(cfg.cfg_dropout_prob > 0)
    for name, vocab in cfg.class_vocab_sizes.items():
        spec[f"class_embeddings.{name}.weight"] = ((vocab + n_null, D), "class_emb")
    spec["t_embedder.mlp.0.weight"] = ((D, 256), "linear")
    spec["t_embedder.mlp.0.bias"] = ((D,), "bias")
    spec["t_embedder.mlp.2.weight"] = ((D, D), "linear")
    spec["t_embedder.mlp.2.bias"] = ((D,), "bias")
    for i in range(cfg.n_layer):
        p = f"blocks.{i}."
        spec[p + "attn.c_attn.weight"] = ((3 * D, D), "linear")
        if cfg.bias:
            spec[p + "attn.c_attn.bias"] = ((3 * D,), "bias")
        spec[p + "attn.c_proj.weight"] = ((D, D), "linear")
        if cfg.bias:
            spec[p + "attn.c_proj.bias"] = ((D,), "bias")
        spec[p + "mlp.w1.weight"] = ((H, D), "linear")
        spec[p + "mlp.w2.weight"] = ((H, D), "linear")
        spec[p + "mlp.c_proj.weight"] = ((D, H), "linear")
        spec[p + "adaln_modulation.1.weight"] = ((6 * D, D), "adaln")
        spec[p + "adaln_modulation.1.bias"] = ((6 * D,), "adaln_bias")
    spec["input_proj.weight"] = ((D, L), "linear")
    if cfg.bias:
        spec["input_proj.bias"] = ((D,), "bias")
    spec["final_layer.linear.weight"] = ((L, D), "linear")
    if cfg.bias:
        spec["final_layer.linear.bias"] = ((L,), "bias")
    spec["final_layer.adaln_modulation.1.weight"] = ((2 * D, D), "adaln")
    if cfg.bias:
        spec["final_layer.adaln_modulation.1.bias"] = ((2 * D,), "adaln_bias")
    return spec


def _block_spec(spec: dict, p: str, E: int, H: int, bias: bool) -> None:
    spec[p + "ln_1.weight"] = ((E,), "ln_w")
    spec[p + "ln_1.bias"] = ((E,), "ln_b")
    spec[p + "ln_2.weight"] = ((E,), "ln_w")
    spec[p + "ln_2.bias"] = ((E,), "ln_b")
    spec[p + "attn.c_attn.weight"] = ((3 * E, E), "linear")
    spec[p + "attn.c_proj.weight"] = ((E, E), "linear")
    if bias:
        spec[p + "attn.c_attn.bias"] = ((3 * E,), "bias")
        spec[p + "attn.c_proj.bias"] = ((E,), "bias")
    spec[p + "mlp.w1.weight"] = ((H, E), "linear")
    spec[p + "mlp.w2.weight"] = ((H, E), "linear")
    spec[p + "mlp.c_proj.weight"] = ((E, H), "linear")


def _mcab_spec(spec: dict, p: str, E: int, H: int, bias: bool, n_inducing: int) -> None:
    if n_inducing > 0:
        spec[p + "inducing_points"] = ((n_inducing, E), "normal1")
    for ln in ("ln_1", "ln_1q", "ln_2"):
        spec[p + ln + ".weight"] = ((E,), "ln_w")
        spec[p + ln + ".bias"] = ((E,), "ln_b")
    spec[p + "attn.c_attn.weight"] = ((2 * E, E), "linear")
    spec[p + "attn.c_attn_q.weight"] = ((E, E), "linear")
    spec[p + "attn.c_proj.weight"] = ((E, E), "linear")
    if bias:
        spec[p + "attn.c_attn.bias"] = ((2 * E,), "bias")
        spec[p + "attn.c_attn_q.bias"] = ((E,), "bias")
        spec[p + "attn.c_proj.bias"] = ((E,), "bias")
    spec[p + "mlp.w1.weight"] = ((H, E), "linear")
    spec[p + "mlp.w2.weight"] = ((H, E), "linear")
    spec[p + "mlp.c_proj.weight"] = ((E, H), "linear")


def vae_state_spec(cfg: VAEConfig) -> dict[str, tuple[tuple[int, ...], str]]:
    """name -> (shape, init-kind) for `scldm.vae.TransformerVAE.state_dict()` (reference `vae.py:15-27`)."""
    assert cfg.shared_embedding and not cfg.use_adaln, "only the shipped VAE topology is specified"
    E, L, H, M = cfg.n_embed, cfg.n_embed_latent, cfg.hidden, cfg.n_inducing_points
    spec: dict[str, tuple[tuple[int, ...], str]] = {}
    if cfg.positional_encoding:
        spec["encoder.pos_embed"] = ((1, M, E), "small")
    for i in range(cfg.n_layer):
        _block_spec(spec, f"encoder.encoder_layers.{i}.", E, H, cfg.bias)
    _mcab_spec(spec, "encoder.ca_layer.", E, H, cfg.bias, M)
    spec["encoder.encoder_latent_input.0.weight"] = ((L, E), "linear")
    if cfg.bias:
        spec["encoder.encoder_latent_input.0.bias"] = ((L,), "bias")
    for i in range(cfg.n_layer):
        _block_spec(spec, f"decoder.decoder_layers.{i}.", E, H, cfg.bias)
    spec["decoder.decoder_latent_input.1.weight"] = ((E, L), "linear")
    if cfg.bias:
        spec["decoder.decoder_latent_input.1.bias"] = ((E,), "bias")
    _mcab_spec(spec, "decoder.decoder_cross_attention.", E, H, cfg.bias, 0)
    if cfg.shared_theta:
        spec["decoder_head.theta.weight"] = ((cfg.n_genes + 1, 1), "theta")
        spec["decoder_head.params.weight"] = ((1, E), "linear")
        spec["decoder_head.params.bias"] = ((1,), "bias")
    else:
        spec["decoder_head.params.weight"] = ((2, E), "linear")
        spec["decoder_head.params.bias"] = ((2,), "bias")
    spec["input_layer.gene_embedding.weight"] = ((cfg.n_genes + 1, E), "normal1")
    return spec


def _rng(seed: int, name: str) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(key=[seed & 0xFFFFFFFFFFFFFFFF, zlib.crc32(name.encode())]))


def _draw(name: str, shape: tuple[int, ...], kind: str, seed: int) -> np.ndarray:
    g = _rng(seed, name)
    if kind == "sincos":
        return sincos_pos_embed(shape[2], shape[1]).reshape(shape)
    if kind == "linear":  # xavier-uniform scale, like the reference's `_basic_init`
        fan_out, fan_in = shape
        a = float(np.sqrt(6.0 / (fan_in + fan_out)))
        return g.uniform(-a, a, size=shape).astype(np.float32)
    if kind == "bias":
        return (0.02 * g.standard_normal(shape)).astype(np.float32)
    if kind == "adaln":  # de-zeroed (reference zero-inits these, `nnets.py:480-489`)
        return (0.02 * g.standard_normal(shape)).astype(np.float32)
    if kind == "adaln_bias":
        return (0.05 * g.standard_normal(shape)).astype(np.float32)
    if kind == "class_emb":  # larger than the reference's 0.02 so that conditioning visibly matters
        return (0.5 * g.standard_normal(shape)).astype(np.float32)
    if kind == "ln_w":
        return (1.0 + 0.1 * g.standard_normal(shape)).astype(np.float32)
    if kind == "ln_b":
        return (0.1 * g.standard_normal(shape)).astype(np.float32)
    if kind == "small":
        return (0.02 * g.standard_normal(shape)).astype(np.float32)
    if kind == "normal1":
        return g.standard_normal(shape).astype(np.float32)
    if kind == "theta":
        return (0.5 * g.standard_normal(shape)).astype(np.float32)
    raise KeyError(kind)


def make_state_dict(spec: dict[str, tuple[tuple[int, ...], str]], seed: int) -> dict[str, torch.Tensor]:
    return {name: torch.from_numpy(_draw(name, shape, kind, seed)) for name, (shape, kind) in spec.items()}


def dit_state_dict(cfg: DiTConfig, seed: int = 1234) -> dict[str, torch.Tensor]:
    return make_state_dict(dit_state_spec(cfg), seed)


def vae_state_dict(cfg: VAEConfig, seed: int = 1234) -> dict[str, torch.Tensor]:
    return make_state_dict(vae_state_spec(cfg), seed)


def size_factor_tables(class_vocab_sizes: dict[str, int], seed: int = 1234) -> tuple[dict, dict]:
    """Synthetic per-class log-size-factor tables: mu ~ U(7,9), sd ~ U(0.2,0.5).

    Layout mirrors the pickles the reference loads into `VocabularyEncoderSimplified`
    (`encoder.py:96-134`): {condition_key: {class_idx: value}}.
    """
    mu: dict[str, dict[int, float]] = {}
    sd: dict[str, dict[int, float]] = {}
    for name, vocab in class_vocab_sizes.items():
        g = _rng(seed, "size_factor." + name)
        mu[name] = {i: float(v) for i, v in enumerate(g.uniform(7.0, 9.0, size=vocab))}
        sd[name] = {i: float(v) for i, v in enumerate(g.uniform(0.2, 0.5, size=vocab))}
    return mu, sd


def randn(name: str, shape: tuple[int, ...], seed: int = 4321) -> torch.Tensor:
    return torch.from_numpy(_rng(seed, name).standard_normal(shape).astype(np.float32))


def randint(name: str, high: int, shape: tuple[int, ...], seed: int = 4321) -> torch.Tensor:
    return torch.from_numpy(_rng(seed, name).integers(0, high, size=shape, dtype=np.int64))
