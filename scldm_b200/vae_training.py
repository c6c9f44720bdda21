"""VAE training step on the device (reference: `VAE.training_step`, `src/scldm/models.py:249-287`; `TransformerVAE.forward`,
`vae.py:29-56`; `VAE.loss`, `models.py:233-247`; `log_nb_positive`, `distributions.py:6-42`; optimizer `AdamWLegacy`,
`optimizers.py:72-141` with `lr: 1e-3, weight_decay: 0` from `experiments/configs/model/vae_base.yaml:56-60`;
`gradient_clip_val: 10` from `configs/training/default.yaml:15`; LR schedule `wsd_schedule`, `_utils.py:19-60`).

`VAETrainer` owns ONE flat fp32 parameter buffer (the module's `nn.Parameter`s become views into it, so `state_dict()` is
unchanged), a flat gradient buffer of the same layout (`p.grad` are views into it) and the AdamW moments.  Forward, NB loss,
backward and the optimizer run in hand-written sm_100a kernels through the C-ABI (`scldm_vae_train_step`, `scldm_adamw_step`);
there is no eager fallback.  Data parallelism: one process per GPU, one NCCL all-reduce of the flat gradient (what DDP does for
the reference, `experiments/scripts/train.py`)."""

from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib
from .layers import InputTransformerVAE
from .vae import TransformerVAE, shared_gene_vector

BLOCK_TENSORS = ["ln_1.weight", "ln_1.bias", "attn.c_attn.weight", "attn.c_proj.weight", "ln_2.weight", "ln_2.bias", "mlp.w1.weight",
                 "mlp.w2.weight", "mlp.c_proj.weight"]
MCAB_TENSORS = ["ln_1.weight", "ln_1.bias", "ln_1q.weight", "ln_1q.bias", "ln_2.weight", "ln_2.bias", "attn.c_attn.weight",
                "attn.c_attn_q.weight", "attn.c_proj.weight", "mlp.w1.weight", "mlp.w2.weight", "mlp.c_proj.weight"]
BLOCK_SIZE, MCAB_SIZE = 12672, 12736


def flat_layout(n_layer: int, shapes: dict[str, tuple]):
    """Order and element offsets of the trainable VAE tensors in the flat buffers (groups as include/scldm_b200.h lays them out).
    Returns (names, offsets, group offsets, n_params)."""
    names, groups = [], {}

    def group(key, tensors):
        groups[key] = len(names)
        names.extend(tensors)

    group("emb", ["input_layer.gene_embedding.weight"])
    group("theta", ["decoder_head.theta.weight"])
    group("head_w", ["decoder_head.params.weight"])
    group("head_b", ["decoder_head.params.bias"])
    group("enc_ca", ["encoder.ca_layer." + t for t in MCAB_TENSORS])
    group("dec_ca", ["decoder.decoder_cross_attention." + t for t in MCAB_TENSORS])
    group("inducing", ["encoder.ca_layer.inducing_points"])
    group("enc_blocks", [f"encoder.encoder_layers.{l}.{t}" for l in range(n_layer) for t in BLOCK_TENSORS])
    group("dec_blocks", [f"decoder.decoder_layers.{l}.{t}" for l in range(n_layer) for t in BLOCK_TENSORS])
    group("enc_lat", ["encoder.encoder_latent_input.0.weight"])
    group("dec_lat", ["decoder.decoder_latent_input.1.weight"])
    missing = [n for n in names if n not in shapes]
    extra = [n for n in shapes if n not in set(names)]
    if missing or extra:
        raise NotImplementedError(f"unexpected VAE parameter set (missing {missing}, extra {extra})")
    off, offsets = 0, {}
    for name in names:
        offsets[name] = off
        off += (math.prod(shapes[name]) + 3) // 4 * 4
    goff = {k: offsets[names[i]] for k, i in groups.items()}
    return names, offsets, goff, off


class VAETrainer:
    """Flat-buffer training state of a `scldm_b200.vae.TransformerVAE` living on a CUDA device."""

    def __init__(self, vae: TransformerVAE, *, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_grad_norm=10.0, lr_lambda=None,
                 process_group=None, exact: bool = False):
        cfg = vae.config()
        dev = vae.input_layer.gene_embedding.weight.device
        if dev.type != "cuda":
            raise RuntimeError("VAETrainer needs the VAE on a CUDA device (there is no CPU training path)")
        if (cfg.n_embed, cfg.n_embed_latent, cfg.n_inducing_points, cfg.n_head, cfg.n_head_cross) != (32, 16, 16, 8, 4) or cfg.bias or cfg.use_adaln \
                or not cfg.shared_theta or not cfg.shared_embedding or cfg.agg_func not in InputTransformerVAE.AGG_CODES:
            raise NotImplementedError(f"sm_100a VAE training kernels cover the vae_base.yaml architecture (n_embed 32, 16 x 16 latents, 8 / 4 heads, "
                                      f"no bias, no adaLN, shared theta / embedding, multiplicative agg_func); got {cfg}")
        if cfg.n_layer > _lib.MAX_LAYERS:
            raise NotImplementedError("n_layer <= 32")
        if float(getattr(vae.decoder_head, "t", 1.0)) != 1.0:
            raise NotImplementedError("NB head temperature != 1")
        self.lib = _lib.load()
        self.vae, self.cfg, self.device = vae, cfg, dev
        self.lr, self.betas, self.eps, self.weight_decay, self.max_grad_norm = lr, betas, eps, weight_decay, max_grad_norm
        self.lr_lambda, self.pg, self.exact = lr_lambda, process_group, exact
        self.world = torch.distributed.get_world_size(process_group) if self._dist() else 1
        self.step_count = 0

        sd = dict(vae.named_parameters())
        shapes = {k: tuple(v.shape) for k, v in sd.items() if v.requires_grad}
        order, offsets, goff, self.n_params = flat_layout(cfg.n_layer, shapes)
        assert goff["dec_ca"] - goff["enc_ca"] == MCAB_SIZE and (cfg.n_layer == 0 or goff["dec_blocks"] - goff["enc_blocks"] == cfg.n_layer * BLOCK_SIZE)
        self.offsets = offsets
        self.flat = torch.zeros(self.n_params, dtype=torch.float32, device=dev)
        self.grad = torch.zeros_like(self.flat)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.scratch = torch.zeros(512, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for name in order:
                p = sd[name]
                view = self.flat[offsets[name]: offsets[name] + p.numel()].view(p.shape)
                view.copy_(p.detach().to(torch.float32))
                p.data = view                                          # the module's parameters now alias the flat buffer
                p.grad = self.grad[offsets[name]: offsets[name] + p.numel()].view(p.shape)
        pos = vae.encoder.pos_embed
        self.pos = pos.detach().reshape(16, 32).to(torch.float32).contiguous() if pos is not None else None

        s = _lib.VaeTrain()
        s.n_layer, s.n_ids, s.agg_func, s.has_pos = cfg.n_layer, cfg.n_genes + 1, InputTransformerVAE.AGG_CODES[cfg.agg_func], int(self.pos is not None)
        s.eps = float(cfg.layernorm_eps)
        s.params, s.grads = self.flat.data_ptr(), self.grad.data_ptr()
        for k, v in goff.items():
            setattr(s, k, v)
        s.n_params = self.n_params
        s.pos = self.pos.data_ptr() if self.pos is not None else None
        self.struct = s
        self._ws, self._ws_key = None, None
        self.last_nll = None
        self._named = [(name, sd[name]) for name in order]
        object.__setattr__(vae, "_trainer", self)     # TransformerVAE.forward in training mode routes through the autograd bridge

    def _attach_grads(self) -> bool:
        """`optimizer.zero_grad(set_to_none=True)` (Lightning's default) drops the `p.grad` views: re-create them.  Returns True when
        any was missing, i.e. the caller zeroed the gradients and the flat buffer has to be cleared before accumulating."""
        dropped = False
        for name, p in self._named:
            if p.grad is None or p.grad.data_ptr() != self.grad.data_ptr() + 4 * self.offsets[name]:
                dropped = dropped or p.grad is None
                p.grad = self.grad[self.offsets[name]: self.offsets[name] + p.numel()].view(p.shape)
        return dropped

    # ------------------------------------------------------------------------------------------------------------
    def _dist(self) -> bool:
        return torch.distributed.is_available() and torch.distributed.is_initialized()

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def workspace(self, B: int, S: int, G: int) -> torch.Tensor:
        if self._ws is None or self._ws_key != (B, S, G):
            nbytes = int(self.lib.scldm_vae_train_workspace_bytes(C.byref(self.struct), B, S, G))
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self._ws_key = (B, S, G)
        return self._ws

    def zero_grad(self) -> None:
        self.grad.zero_()

    def forward_backward(self, counts, genes, library_size, counts_subset, genes_subset, *, backward: bool = True, zero_grads: bool = True,
                         global_batch: int | None = None, want_mu: bool = False):
        """`TransformerVAE.forward` + `VAE.loss` (+ autograd): returns ({"llh": mean NLL over the local cells, "per_cell": ...}, h_z[, mu]).
        Gradients of `recon_loss.sum(dim=1).mean()` over `global_batch` cells (default: the local batch) land in the flat buffer."""
        B, G = counts.shape
        S = counts_subset.shape[1]
        gvec = shared_gene_vector(genes).to(torch.int64).contiguous()
        counts = counts.contiguous().float()
        cs = counts_subset.contiguous().float()
        gs = genes_subset.contiguous().to(torch.int64)
        lib = library_size.reshape(-1).contiguous().float()
        nll = torch.empty(B, dtype=torch.float32, device=self.device)
        z = torch.empty(B, 16, 16, dtype=torch.float32, device=self.device)
        mu = torch.empty(B, G, dtype=torch.float32, device=self.device) if want_mu else None
        ws = self.workspace(B, S, G)
        scale = 1.0 / float(global_batch if global_batch is not None else B)
        with torch.cuda.device(self.device):
            rc = self.lib.scldm_vae_train_step(C.byref(self.struct), gs.data_ptr(), cs.data_ptr(), S, gvec.data_ptr(), counts.data_ptr(), lib.data_ptr(),
                                               B, G, scale, int(backward), int(zero_grads), int(self.exact), nll.data_ptr(), z.data_ptr(),
                                               mu.data_ptr() if want_mu else None, ws.data_ptr(), ws.numel(), self._stream())
        _lib.check(rc, "scldm_vae_train_step")
        self.last_nll = nll
        out = {"llh": nll.mean(), "per_cell": nll}
        return (out, z, mu) if want_mu else (out, z)

    def backward_from(self, dmu: torch.Tensor, dtheta: torch.Tensor | None, inputs, *, zero_grads: bool = False) -> None:
        """Backward of the last forward-only `forward_backward(backward=False)` on `inputs` = (genes, library_size, counts_subset,
        genes_subset) from dLoss/dmu [B,G] (and dLoss/dtheta [B,G] of the expanded theta): the autograd bridge's backward."""
        genes, library_size, counts_subset, genes_subset = inputs
        zero_grads = self._attach_grads() or zero_grads
        B, G = dmu.shape
        S = counts_subset.shape[1]
        gvec = shared_gene_vector(genes).to(torch.int64).contiguous()
        cs = counts_subset.contiguous().float()
        gs = genes_subset.contiguous().to(torch.int64)
        lib = library_size.reshape(-1).contiguous().float()
        dmu = dmu.contiguous().float()
        dth = dtheta.contiguous().float() if dtheta is not None else None
        ws = self.workspace(B, S, G)
        with torch.cuda.device(self.device):
            rc = self.lib.scldm_vae_train_backward(C.byref(self.struct), gs.data_ptr(), cs.data_ptr(), S, gvec.data_ptr(), lib.data_ptr(), dmu.data_ptr(),
                                                   dth.data_ptr() if dth is not None else None, B, G, int(zero_grads), int(self.exact), ws.data_ptr(),
                                                   ws.numel(), self._stream())
        _lib.check(rc, "scldm_vae_train_backward")
        self.vae._packed_dec = None          # an optimizer of the caller's will move the weights: rebuild the inference packs lazily
        self.vae._packed_enc = None

    def allreduce_grads(self) -> None:
        if self._dist() and self.world > 1:
            torch.distributed.all_reduce(self.grad, group=self.pg)

    def optimizer_step(self) -> None:
        """clip_grad_norm_(max_grad_norm) + AdamWLegacy, one fused pass over the flat buffers."""
        self.step_count += 1
        lr = self.lr * (self.lr_lambda(self.step_count - 1) if self.lr_lambda is not None else 1.0)
        with torch.cuda.device(self.device):
            rc = self.lib.scldm_adamw_step(self.flat.data_ptr(), self.grad.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), self.n_params,
                                           float(lr), float(self.betas[0]), float(self.betas[1]), float(self.eps), float(self.weight_decay),
                                           self.step_count, float(self.max_grad_norm or 0.0), 1.0, self.scratch.data_ptr(), None, None, self._stream())
        _lib.check(rc, "scldm_adamw_step")
        # the inference-side packed weights are stale now (rebuilt lazily by TransformerVAE.packed_decoder / packed_encoder)
        self.vae._packed_dec = None
        self.vae._packed_enc = None

    def training_step(self, batch: dict[str, torch.Tensor]) -> torch.Tensor:
        """`VAE.training_step` (`models.py:249-287`) + backward + DDP all-reduce + clip + optimizer.  `batch` uses the reference's keys
        (`ModelEnum`, constants.py): counts, genes, library_size, counts_subset, genes_subset.  Returns the local mean loss (0-d tensor)."""
        B = batch["counts"].shape[0]
        out, _ = self.forward_backward(batch["counts"], batch["genes"], batch["library_size"], batch["counts_subset"], batch["genes_subset"],
                                       global_batch=B * self.world)
        self.allreduce_grads()
        self.optimizer_step()
        return out["llh"]


class _VAETrainFunction(torch.autograd.Function):
    """autograd bridge: `TransformerVAE.forward` in training mode -> `VAETrainer` forward kernels; `loss.backward()` ->
    `scldm_vae_train_backward`.  Parameter gradients land in the flat buffer (the `p.grad` views), as autograd would leave them."""

    @staticmethod
    def forward(ctx, anchor, trainer, counts, genes, library_size, counts_subset, genes_subset):
        ctx.trainer = trainer
        ctx.inputs = (genes, library_size, counts_subset, genes_subset)
        _, z, mu = trainer.forward_backward(counts, genes, library_size, counts_subset, genes_subset, backward=False, want_mu=True)
        gvec = shared_gene_vector(genes).to(torch.int64)
        theta = torch.exp(trainer.vae.decoder_head.theta.weight.detach()[gvec].squeeze(-1)).unsqueeze(0).expand(mu.shape[0], -1).contiguous()
        ctx.mark_non_differentiable(z)
        return mu, theta, z

    @staticmethod
    def backward(ctx, dmu, dtheta, _dz):
        ctx.trainer.backward_from(dmu, dtheta, ctx.inputs, zero_grads=False)
        return (None,) * 7


def differentiable_forward(trainer: VAETrainer, counts, genes, library_size, counts_subset, genes_subset):
    """`TransformerVAE.forward` (`vae.py:29-56`) whose outputs carry a grad_fn: `({"mu", "theta"}, h_z)`; any loss built from mu / theta
    (`VAE.loss`, `models.py:233-247`) can call `.backward()` - call `trainer.zero_grad()` first, gradients ACCUMULATE like autograd's.
    h_z is returned detached (the reference's training loss does not use it)."""
    anchor = torch.zeros((), device=trainer.device, requires_grad=True)   # makes autograd call backward although no input requires grad
    mu, theta, z = _VAETrainFunction.apply(anchor, trainer, counts, genes, library_size, counts_subset, genes_subset)
    return {"mu": mu, "theta": theta}, z
