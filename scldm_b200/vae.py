"""Drop-in for `scldm.vae.TransformerVAE` (`vae.py:15-87`): same constructor, same `state_dict`
keys (`encoder.*`, `decoder.*`, `decoder_head.*`, `input_layer.*`), `.decode` runs the fused
sm_100a decoder (latent blocks -> MCAB unpool -> NB head -> optional Gamma-Poisson draw)."""

from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .config import VAEConfig
from .layers import InputTransformerVAE, weights_key
from .nnets import Decoder, Encoder
from .pack import PackedVAEDecoder, PackedVAEEncoder
from .pack256 import PackedVAE256Decoder, PackedVAE256Encoder
from .stochastic_layers import NegativeBinomial, NegativeBinomialTransformerLayer


def shared_gene_vector(genes: torch.Tensor) -> torch.Tensor:
    """The reference feeds `genes` as a (B,G) int64 tile of ONE row (`datamodule.py:689`).
    The kernels take that row once; a batch with differing rows is rejected loudly."""
    if genes.dim() == 1:
        return genes.contiguous()
    if genes.shape[0] > 1 and genes.stride(0) != 0 and not bool((genes == genes[0:1]).all()):
        raise NotImplementedError("decode with per-cell gene orderings: only the tiled layout the reference's tokenizer emits is supported")
    return genes[0].contiguous()


class TransformerVAE(nn.Module):
    def __init__(self, encoder: Encoder, decoder: Decoder, decoder_head: NegativeBinomialTransformerLayer,
                 input_layer: InputTransformerVAE):
        super().__init__()
        self.encoder = encoder
        self.decoder = decoder
        self.decoder_head = decoder_head
        self.input_layer = input_layer
        self._packed_dec: PackedVAEDecoder | None = None
        self._packed_key = None
        self.sample_seed = 0
        self.sample_offset = 0
        self.decode_precision = "bf16"  # "bf16": MCAB on tensor cores; "fp32": exact CUDA-core variant

    @classmethod
    def from_config(cls, cfg: VAEConfig) -> "TransformerVAE":
        return cls(encoder=Encoder(**cfg.encoder_kwargs()), decoder=Decoder(**cfg.decoder_kwargs()),
                   decoder_head=NegativeBinomialTransformerLayer(**cfg.head_kwargs()),
                   input_layer=InputTransformerVAE(**cfg.input_kwargs()))

    def config(self) -> VAEConfig:
        d, e = self.decoder.hparams_, self.encoder.hparams_
        return VAEConfig(n_genes=d["n_genes"], n_embed=d["n_embed"], n_embed_latent=d["n_embed_latent"],
                         n_inducing_points=d["n_inducing_points"], n_layer=d["n_layer"], n_head=d["n_head"],
                         n_head_cross=d["n_head_cross"], bias=d["bias"], multiple_of=d["multiple_of"],
                         layernorm_eps=d["layernorm_eps"], positional_encoding=e["positional_encoding"],
                         shared_embedding=d["shared_embedding"], use_adaln=d["use_adaln"],
                         shared_theta=self.decoder_head.shared_theta, agg_func=self.input_layer.agg_func)

    def packed_decoder(self) -> PackedVAEDecoder:
        key = weights_key(self, "_wkey_dec", lambda k: not k.startswith("encoder."))
        dev = self.input_layer.gene_embedding.weight.device
        if dev.type != "cuda":
            raise RuntimeError("scldm_b200.TransformerVAE runs on CUDA only (no CPU fallback): call .cuda() first")
        if self._packed_dec is None or self._packed_key != key:
            cfg = self.config()
            cls = PackedVAE256Decoder if cfg.n_embed == 256 else PackedVAEDecoder      # kernels exist for n_embed 32 (vae_base.yaml) and 256 (census scale)
            self._packed_dec = cls({k: v.detach() for k, v in self.state_dict(keep_vars=True).items()}, cfg, dev)
            self._packed_key = key
        return self._packed_dec

    def _decode_op(self, packed):
        """The decode entry point of the packed weights' kernel family (same contract)."""
        if isinstance(packed, PackedVAE256Decoder):
            return lambda z, g, lib, precision="bf16", **kw: ops.vae256_decode(packed, z, g, lib, **kw)     # tensor cores only at n_embed = 256
        return lambda z, g, lib, **kw: ops.vae_decode(packed, z, g, lib, **kw)

    def decode(self, z: torch.Tensor, genes: torch.Tensor, library_size: torch.Tensor,
               condition: dict[str, torch.Tensor] | None = None) -> NegativeBinomial:
        """`TransformerVAE.decode` (`vae.py:71-87`) -> NB(mu, theta); `.sample()` draws counts on device."""
        if condition is not None:
            raise NotImplementedError("conditioned decoder (use_adaln=True) is not on the shipped path")
        packed = self.packed_decoder()
        gvec = shared_gene_vector(genes)
        zc = z.contiguous().float()
        prec = self.decode_precision
        dec = self._decode_op(packed)
        mu, theta, _ = dec(zc, gvec, library_size, want_mu=True, want_counts=False, precision=prec)
        theta_full = theta.unsqueeze(0).expand(z.shape[0], -1) if theta.dim() == 1 else theta
        n_cells = z.shape[0]

        def sampler():
            # Philox streams are keyed by (sample_seed, cell offset): every `.sample()` call - of this or of any other `decode()` -
            # takes the next `n_cells` offsets, so repeated draws are independent; set `vae.sample_seed` / `vae.sample_offset` to replay
            offset = self.sample_offset
            self.sample_offset = offset + n_cells
            _, _, counts = dec(zc, gvec, library_size, want_mu=False, want_counts=True, seed=self.sample_seed, cell_offset=offset, precision=prec)
            return counts

        return NegativeBinomial(mu, theta_full, _sampler=sampler)

    def decode_counts(self, z, genes, library_size, seed: int = 0, cell_offset: int = 0, want_mu: bool = False,
                      out_counts=None, out_mu=None):
        """Fast path of `decode(...).sample()` (`models.py:818-819`): one pass, counts only (+mu on request),
        optionally written straight into caller-provided output rows."""
        packed = self.packed_decoder()
        mu, theta, counts = self._decode_op(packed)(z.contiguous().float(), shared_gene_vector(genes), library_size,
                                                    want_mu=want_mu, want_counts=True, seed=seed, cell_offset=cell_offset,
                                                    out_counts=out_counts, out_mu=out_mu, precision=self.decode_precision)
        return counts, mu, theta

    def packed_encoder(self) -> PackedVAEEncoder:
        key = weights_key(self, "_wkey_enc", lambda k: k.startswith("encoder.") or k.startswith("input_layer."))
        dev = self.input_layer.gene_embedding.weight.device
        if dev.type != "cuda":
            raise RuntimeError("scldm_b200.TransformerVAE runs on CUDA only (no CPU fallback): call .cuda() first")
        if getattr(self, "_packed_enc", None) is None or self._packed_enc_key != key:
            cfg = self.config()
            cls = PackedVAE256Encoder if cfg.n_embed == 256 else PackedVAEEncoder
            self._packed_enc = cls({k: v.detach() for k, v in self.state_dict(keep_vars=True).items()}, cfg, dev)
            self._packed_enc_key = key
        return self._packed_enc

    def encode(self, counts, genes, counts_subset=None, genes_subset=None):
        """`TransformerVAE.encode` (`vae.py:58-69`): the subset tensors are used when given, else (counts, genes)."""
        c = counts_subset if counts_subset is not None else counts
        g = genes_subset if genes_subset is not None else genes
        packed = self.packed_encoder()
        if isinstance(packed, PackedVAE256Encoder):
            return ops.vae256_encode(packed, g.contiguous(), c.contiguous())
        return ops.vae_encode(packed, g.contiguous(), c.contiguous())

    def forward(self, counts, genes, library_size, counts_subset=None, genes_subset=None):
        """`TransformerVAE.forward` (`vae.py:29-56`): encode the (subset) tokens, decode every gene -> `({"mu", "theta"}, h_z)`.
        In training mode with gradients enabled and a `vae_training.VAETrainer` attached (its constructor attaches itself) the outputs
        carry a grad_fn: `VAE.loss(...)` -> `.backward()` fills `p.grad` through the library's backward kernels, so the reference's
        Lightning `training_step` / optimizer loop runs unchanged.  Otherwise inference kernels, no autograd graph."""
        trainer = getattr(self, "_trainer", None)
        if self.training and torch.is_grad_enabled() and trainer is not None:
            from .vae_training import differentiable_forward

            return differentiable_forward(trainer, counts, genes, library_size, counts_subset if counts_subset is not None else counts,
                                          genes_subset if genes_subset is not None else genes)
        with torch.no_grad():
            h_z = self.encode(counts, genes, counts_subset, genes_subset)
            dist = self.decode(h_z, genes, library_size)
            return {"mu": dist.mu, "theta": dist.theta}, h_z

    @staticmethod
    def reconstruction_loss(counts: torch.Tensor, params: dict[str, torch.Tensor]) -> dict[str, torch.Tensor]:
        """`VAE.loss` for the NB head (`models.py:233-247`): `{"llh": (-log_nb_positive(...)).sum(dim=1).mean()}`
        (`LossEnum.LLH_LOSS`, `constants.py:35`); the per-cell sums come from one fused device pass (`ops.nb_nll`)."""
        per_cell = ops.nb_nll(counts.float(), params["mu"], params["theta"])
        return {"llh": per_cell.mean(), "per_cell": per_cell}
