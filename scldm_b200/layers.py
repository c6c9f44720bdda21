"""Parameter containers with the reference's module tree (`scldm/layers.py`) so that reference
checkpoints load with identical `state_dict` keys and shapes.

These modules hold weights only.  The arithmetic lives in the sm_100a kernels behind
`scldm_b200.nnets` / `scldm_b200.vae`; calling a container's `forward` raises instead of silently
running a PyTorch implementation (no CPU / eager fallback by design).
"""

from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from .config import swiglu_hidden


def weights_key(module: nn.Module, cache_attr: str, select=None) -> tuple:
    """(data_ptr, version) of the module's parameters and buffers: the validity key of a packed-weights cache.  The tensor list is
    collected once (a `state_dict()` walk per call costs a millisecond of host time, which a synchronous generation step waits
    for); in-place updates (`load_state_dict`, optimizer steps) bump `_version`, `.to()` / `.cuda()` change `data_ptr`, and a
    re-registered parameter changes the module's tensor count."""
    cached = module.__dict__.get(cache_attr)
    n_now = sum(len(m._parameters) + len(m._buffers) for m in module.modules())
    if cached is None or cached[0] != n_now:
        named = list(module.named_parameters()) + list(module.named_buffers())
        cached = (n_now, [t for k, t in named if select is None or select(k)])
        module.__dict__[cache_attr] = cached
    return tuple((t.data_ptr(), t._version) for t in cached[1])


class _KernelOnly(nn.Module):
    def forward(self, *args, **kwargs):  # pragma: no cover - guard
        raise RuntimeError(
            f"{type(self).__name__} is a weight container; its math runs inside the fused sm_100a kernels "
            "(call the owning DiT / TransformerVAE). scldm_b200 has no eager fallback."
        )


class SelfAttention(_KernelOnly):
    """weights of reference `SelfAttention` (`layers.py:121-141`): c_attn (D->3D, q|k|v), c_proj (D->D)."""

    def __init__(self, n_embed: int, n_head: int, dropout: float, bias: bool):
        super().__init__()
        assert n_embed % n_head == 0
        self.n_head, self.n_embed, self.dropout = n_head, n_embed, dropout
        self.c_attn = nn.Linear(n_embed, 3 * n_embed, bias=bias)
        self.c_proj = nn.Linear(n_embed, n_embed, bias=bias)


class CrossAttention(_KernelOnly):
    """weights of reference `CrossAttention` (`layers.py:229-246`): c_attn (k|v), c_attn_q, c_proj."""

    def __init__(self, n_embed: int, n_head: int, dropout: float, bias: bool):
        super().__init__()
        self.n_head, self.n_embed = n_head, n_embed
        self.c_attn = nn.Linear(n_embed, 2 * n_embed, bias=bias)
        self.c_attn_q = nn.Linear(n_embed, n_embed, bias=bias)
        self.c_proj = nn.Linear(n_embed, n_embed, bias=bias)


class MLP(_KernelOnly):
    """weights of the SwiGLU MLP (`layers.py:161-171`): w1, w2 (E->H), c_proj (H->E), no biases."""

    def __init__(self, n_embed: int, multiple_of: int):
        super().__init__()
        hidden = swiglu_hidden(n_embed, multiple_of)
        self.w1 = nn.Linear(n_embed, hidden, bias=False)
        self.w2 = nn.Linear(n_embed, hidden, bias=False)
        self.c_proj = nn.Linear(hidden, n_embed, bias=False)


class Block(_KernelOnly):
    """weights of reference `Block` (`layers.py:177-206`)."""

    def __init__(self, n_embed, n_head, dropout, bias, norm_layer, multiple_of, layernorm_eps, use_adaln=False,
                 elementwise_affine=True):
        super().__init__()
        assert norm_layer == "layernorm"
        self.ln_1 = nn.LayerNorm(n_embed, eps=layernorm_eps, elementwise_affine=elementwise_affine)
        self.ln_2 = nn.LayerNorm(n_embed, eps=layernorm_eps, elementwise_affine=elementwise_affine)
        self.attn = SelfAttention(n_embed=n_embed, n_head=n_head, dropout=dropout, bias=bias)
        self.mlp = MLP(n_embed=n_embed, multiple_of=multiple_of)
        self.use_adaln = use_adaln
        if use_adaln:
            self.adaln_modulation = nn.Sequential(nn.SiLU(), nn.Linear(n_embed, 6 * n_embed, bias=True))


class CrossAttentionBlock(_KernelOnly):
    """weights of the MCAB (`layers.py:267-303`)."""

    def __init__(self, n_embed, n_inducing_points, n_head, dropout, bias, norm_layer, multiple_of, layernorm_eps,
                 use_adaln=False):
        super().__init__()
        assert norm_layer == "layernorm"
        if use_adaln:
            raise NotImplementedError("adaLN MCAB (use_adaln=True) is not on the shipped VAE path (vae_base.yaml:35)")
        self.inducing_points = None if n_inducing_points == 0 else nn.Parameter(torch.randn(n_inducing_points, n_embed))
        self.ln_1 = nn.LayerNorm(n_embed, eps=layernorm_eps)
        self.ln_1q = nn.LayerNorm(n_embed, eps=layernorm_eps)
        self.attn = CrossAttention(n_embed=n_embed, n_head=n_head, dropout=dropout, bias=bias)
        self.ln_2 = nn.LayerNorm(n_embed, eps=layernorm_eps)
        self.mlp = MLP(n_embed=n_embed, multiple_of=multiple_of)
        self.use_adaln = use_adaln


class TimestepEmbedder(_KernelOnly):
    """weights of `TimestepEmbedder` (`layers.py:339-349`)."""

    def __init__(self, hidden_size: int, frequency_embedding_size: int = 256):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(frequency_embedding_size, hidden_size), nn.SiLU(), nn.Linear(hidden_size, hidden_size))
        self.frequency_embedding_size = frequency_embedding_size


class FinalLayerDit(_KernelOnly):
    """weights of `FinalLayerDit` (`layers.py:388-395`)."""

    def __init__(self, n_embed: int, n_embed_input: int, bias: bool, layernorm_eps: float):
        super().__init__()
        self.norm_final = nn.LayerNorm(n_embed, elementwise_affine=False, eps=layernorm_eps)
        self.linear = nn.Linear(n_embed, n_embed_input, bias=bias)
        self.adaln_modulation = nn.Sequential(nn.SiLU(), nn.Linear(n_embed, 2 * n_embed, bias=bias))


class InputTransformerVAE(_KernelOnly):
    """weights of `InputTransformerVAE` (`layers.py:97-109`).  The multiplicative count transforms ('log1p', the shipped one
    (vae_base.yaml:40), 'log1pzero', 'anscombe', 'sqrt': `layers.py:28-44`) run inside the encoder kernels; the three variants
    with learned count embeddings ('proj', 'projconcat', 'softbin') are not built."""

    AGG_CODES = {"log1p": 0, "log1pzero": 1, "anscombe": 2, "sqrt": 3}

    def __init__(self, n_genes: int, n_embed: int, agg_func: str = "log1p"):
        super().__init__()
        if agg_func not in self.AGG_CODES:
            raise NotImplementedError(f"agg_func='{agg_func}': implemented: {sorted(self.AGG_CODES)}")
        self.gene_embedding = nn.Embedding(n_genes + 1, n_embed)
        self.agg_func = agg_func


def get_1d_sincos_pos_embed(embed_dim: int, seq_len: int) -> np.ndarray:
    """sin half first, then cos (reference `layers.py:367-385`)."""
    assert embed_dim % 2 == 0
    pos = np.arange(seq_len, dtype=np.float32).reshape(-1, 1)
    omega = np.arange(embed_dim // 2, dtype=np.float32) / (embed_dim / 2.0)
    out = pos * (1.0 / (10000**omega)).reshape(1, -1)
    return np.concatenate([np.sin(out), np.cos(out)], axis=1)
