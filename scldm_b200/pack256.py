"""Weight packing for the census-scale VAE (n_embed = 256): reference `state_dict` -> the device blobs of
`scldm_vae256_weights` (include/scldm_b200.h).  The latent Blocks are packed exactly like DiT blocks (`pack.PackedDiT`: fused
attention / MLP weight streams for `dit_blocks_kernel`) with zero biases and a constant modulation row that encodes the affine
LayerNorms; the MCAB matrices are plain K-major UMMA tiles for the slab GEMM."""

from __future__ import annotations

import torch

from . import _lib
from .config import DiTConfig, VAEConfig
from .pack import PackedDiT, pack_kmajor_tiles

E = 256


def supported(cfg: VAEConfig) -> bool:
    return (cfg.n_embed, cfg.n_embed_latent, cfg.n_inducing_points, cfg.n_head, cfg.n_head_cross, cfg.hidden) == (256, 16, 16, 8, 4, 684)


def _check(cfg: VAEConfig) -> None:
    if not supported(cfg):
        raise NotImplementedError(f"the n_embed = 256 kernels cover 8 heads, 4 cross heads, 16 latents x 16, hidden 684; got {cfg}")
    if cfg.bias or cfg.use_adaln or not cfg.shared_embedding or not cfg.shared_theta or cfg.agg_func != "log1p":
        raise NotImplementedError("the n_embed = 256 kernels cover bias=False, use_adaln=False, shared_embedding, shared_theta, agg_func='log1p'")


def _blocks_as_dit(g, prefix: str, n_layer: int, eps: float, device):
    """The latent Blocks `prefix{i}.*` (layers.py:177-226 without adaLN) as a `PackedDiT` + the constant modulation row."""
    H = 684
    sd = {"pos_embed": torch.zeros(1, 16, E), "t_embedder.mlp.0.weight": torch.zeros(E, 256), "t_embedder.mlp.2.weight": torch.zeros(E, E),
          "input_proj.weight": torch.zeros(E, 16), "final_layer.linear.weight": torch.zeros(16, E),
          "final_layer.adaln_modulation.1.weight": torch.zeros(2 * E, E), "t_embedder.mlp.0.bias": torch.zeros(E), "t_embedder.mlp.2.bias": torch.zeros(E)}
    mod = []
    for i in range(n_layer):
        p, q = f"{prefix}{i}.", f"blocks.{i}."
        for n in ("attn.c_attn.weight", "attn.c_proj.weight", "mlp.w1.weight", "mlp.w2.weight", "mlp.c_proj.weight"):
            sd[q + n] = g(p + n)
        assert sd[q + "mlp.w1.weight"].shape == (H, E)
        sd[q + "adaln_modulation.1.weight"] = torch.zeros(6 * E, E)
        one = torch.ones(E)
        mod += [g(p + "ln_1.weight") - 1, g(p + "ln_1.bias"), one, g(p + "ln_2.weight") - 1, g(p + "ln_2.bias"), one]
    mod.append(torch.zeros(2 * E))
    cfg = DiTConfig(n_layer=n_layer, class_vocab_sizes={}, layernorm_eps=eps)
    return PackedDiT(sd, cfg, device), torch.cat(mod).to(device=device, dtype=torch.float32).contiguous()


def _w12_tiles(w1: torch.Tensor, w2: torch.Tensor) -> torch.Tensor:
    H, T = w1.shape[0], -(-w1.shape[0] // 128)
    a, b = torch.zeros(T * 128, E), torch.zeros(T * 128, E)
    a[:H], b[:H] = w1, w2
    return pack_kmajor_tiles(torch.stack([a.view(T, 128, E), b.view(T, 128, E)], 1).reshape(-1, E), 256)


class _Packed256:
    def _finish(self, cfg, device, names):
        s = _lib.Vae256Weights()
        s.n_layer, s.n_ids, s.mlp_tiles, s.has_pos = cfg.n_layer, self.emb.shape[0], 6, int(getattr(self, "has_pos", False))
        s.eps, s.head_b = float(cfg.layernorm_eps), float(getattr(self, "head_b", 0.0))
        s.blocks = self.blocks.struct
        for n in names:
            setattr(s, n, getattr(self, n).data_ptr())
        self.struct = s
        self.device = torch.device(device)
        self.cfg = cfg


class PackedVAE256Decoder(_Packed256):
    def __init__(self, sd: dict, cfg: VAEConfig, device):
        _check(cfg)
        f32 = lambda t: t.detach().to(device=device, dtype=torch.float32).contiguous()  # noqa: E731
        dev = lambda t: t.to(device).contiguous()  # noqa: E731
        g = lambda n: sd[n].detach().float().cpu()  # noqa: E731
        c = "decoder.decoder_cross_attention."
        self.emb = f32(g("input_layer.gene_embedding.weight"))
        self.blocks, self.blocks_mod = _blocks_as_dit(g, "decoder.decoder_layers.", cfg.n_layer, cfg.layernorm_eps, device)
        self.ln1_mod = f32(torch.cat([g(c + "ln_1.weight") - 1, g(c + "ln_1.bias")]))
        self.w_kv = dev(pack_kmajor_tiles(g(c + "attn.c_attn.weight"), 256))
        self.ln1q_w, self.ln1q_b = f32(g(c + "ln_1q.weight")), f32(g(c + "ln_1q.bias"))
        self.w_q = dev(pack_kmajor_tiles(g(c + "attn.c_attn_q.weight"), 256))
        self.w_proj = dev(pack_kmajor_tiles(g(c + "attn.c_proj.weight"), 256))
        self.ln2_w, self.ln2_b = f32(g(c + "ln_2.weight")), f32(g(c + "ln_2.bias"))
        self.w_12 = dev(_w12_tiles(g(c + "mlp.w1.weight"), g(c + "mlp.w2.weight")))
        self.lat_w = f32(g("decoder.decoder_latent_input.1.weight"))
        hw = g("decoder_head.params.weight").reshape(-1)
        self.head_w = f32(hw)
        self.head_b = float(g("decoder_head.params.bias").reshape(-1)[0])
        v = torch.zeros(6 * 128, dtype=torch.float64)
        v[:684] = g(c + "mlp.c_proj.weight").double().T @ hw.double()     # logit = w.x + (W3^T w).s + b
        self.head_v = f32(v.float())
        self.theta_tbl = f32(g("decoder_head.theta.weight").reshape(-1))
        self.shared_theta = True
        self.qp_bf16 = None
        self._finish(cfg, device, ["emb", "blocks_mod", "ln1_mod", "w_kv", "ln1q_w", "ln1q_b", "w_q", "w_proj", "ln2_w", "ln2_b", "w_12", "lat_w", "head_w",
                                   "head_v", "theta_tbl"])


class PackedVAE256Encoder(_Packed256):
    def __init__(self, sd: dict, cfg: VAEConfig, device):
        _check(cfg)
        f32 = lambda t: t.detach().to(device=device, dtype=torch.float32).contiguous()  # noqa: E731
        dev = lambda t: t.to(device).contiguous()  # noqa: E731
        g = lambda n: sd[n].detach().float().cpu()  # noqa: E731
        c = "encoder.ca_layer."
        eps = float(cfg.layernorm_eps)
        self.emb = f32(g("input_layer.gene_embedding.weight"))
        self.blocks, self.blocks_mod = _blocks_as_dit(g, "encoder.encoder_layers.", cfg.n_layer, eps, device)
        self.ln1_mod = f32(torch.cat([g(c + "ln_1.weight"), g(c + "ln_1.bias")]))     # the token kernel takes (weight | bias) as is
        self.w_kv = dev(pack_kmajor_tiles(g(c + "attn.c_attn.weight"), 256))
        ind = g(c + "inducing_points")
        qn = torch.nn.functional.layer_norm(ind, (E,), g(c + "ln_1q.weight"), g(c + "ln_1q.bias"), eps)
        self.q_tbl = f32(qn @ g(c + "attn.c_attn_q.weight").T)      # cell invariant (layers.py:312-313)
        self.inducing = f32(ind)
        self.w_proj = dev(pack_kmajor_tiles(g(c + "attn.c_proj.weight"), 256))
        self.ln2_mod = f32(torch.cat([g(c + "ln_2.weight") - 1, g(c + "ln_2.bias")]))
        self.w_12 = dev(_w12_tiles(g(c + "mlp.w1.weight"), g(c + "mlp.w2.weight")))
        w3 = torch.zeros(E, 6 * 128)
        w3[:, :684] = g(c + "mlp.c_proj.weight")
        self.w_3 = dev(pack_kmajor_tiles(w3, 256))
        self.has_pos = "encoder.pos_embed" in sd
        self.pos = f32(g("encoder.pos_embed").reshape(16, E)) if self.has_pos else f32(torch.zeros(16, E))
        self.ones = f32(torch.ones(E))
        self.out_w = f32(g("encoder.encoder_latent_input.0.weight"))
        self._finish(cfg, device, ["emb", "blocks_mod", "ln1_mod", "w_kv", "q_tbl", "inducing", "w_proj", "ln2_mod", "w_12", "w_3", "pos", "ones", "out_w"])
