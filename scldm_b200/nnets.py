"""Drop-in replacements for `scldm.nnets.{DiT, Encoder, Decoder}`: same constructor kwargs, same
`state_dict` layout, same call signatures -- the arithmetic runs in hand-written sm_100a kernels
through the C-ABI (`include/scldm_b200.h`).  `DiT.forward` is differentiable once a `training.DiTTrainer` is attached.
"""

from __future__ import annotations

from typing import Literal

import torch
import torch.nn as nn

from . import ops
from .config import DiTConfig
from .layers import Block, CrossAttentionBlock, FinalLayerDit, TimestepEmbedder, get_1d_sincos_pos_embed, weights_key
from .pack import PackedDiT


class Encoder(nn.Module):
    """weights of `scldm.nnets.Encoder` (`nnets.py:81-135`); driven by `TransformerVAE.encode`."""

    def __init__(self, n_layer, n_inducing_points, n_embed, n_embed_latent, n_head, n_head_cross, dropout, bias, multiple_of,
                 layernorm_eps, norm_layer, positional_encoding=False):
        super().__init__()
        self.latent_embedding = n_embed_latent
        self.latent_dim = n_inducing_points
        self.pos_embed = nn.Parameter(torch.zeros(1, n_inducing_points, n_embed), requires_grad=False) if positional_encoding else None
        self.ca_layer = CrossAttentionBlock(n_embed=n_embed, n_inducing_points=n_inducing_points, n_head=n_head_cross, dropout=dropout,
                                            bias=bias, norm_layer=norm_layer, multiple_of=multiple_of, layernorm_eps=layernorm_eps)
        self.encoder_layers = nn.ModuleList([
            Block(n_embed=n_embed, n_head=n_head, dropout=dropout, bias=bias, norm_layer=norm_layer, multiple_of=multiple_of,
                  layernorm_eps=layernorm_eps) for _ in range(n_layer)])
        self.encoder_latent_input = nn.Sequential(
            nn.Linear(n_embed, n_embed_latent, bias=bias),
            nn.LayerNorm(n_embed_latent, eps=layernorm_eps, elementwise_affine=False))
        self.hparams_ = dict(n_layer=n_layer, n_inducing_points=n_inducing_points, n_embed=n_embed, n_embed_latent=n_embed_latent,
                             n_head=n_head, n_head_cross=n_head_cross, bias=bias, multiple_of=multiple_of,
                             layernorm_eps=layernorm_eps, positional_encoding=positional_encoding)

    def forward(self, x):
        raise RuntimeError("Encoder is driven through TransformerVAE.encode (fused kernels); no eager fallback")


class Decoder(nn.Module):
    """weights of `scldm.nnets.Decoder` (`nnets.py:147-198`); driven by `TransformerVAE.decode`."""

    def __init__(self, n_genes, n_embed, n_embed_latent, n_head, n_head_cross, n_layer, n_inducing_points, dropout, bias,
                 multiple_of, layernorm_eps, norm_layer, shared_embedding, use_adaln=False):
        super().__init__()
        self.gene_embedding = nn.Embedding(n_genes + 1, n_embed) if not shared_embedding else nn.Identity()
        self.decoder_latent_input = nn.Sequential(
            nn.LayerNorm(n_embed_latent, eps=layernorm_eps, elementwise_affine=False),
            nn.Linear(n_embed_latent, n_embed, bias=bias))
        self.decoder_layers = nn.ModuleList([
            Block(n_embed=n_embed, n_head=n_head, dropout=dropout, bias=bias, norm_layer=norm_layer, multiple_of=multiple_of,
                  layernorm_eps=layernorm_eps, use_adaln=use_adaln) for _ in range(n_layer)])
        self.decoder_cross_attention = CrossAttentionBlock(
            n_embed=n_embed, n_inducing_points=0, n_head=n_head_cross, dropout=dropout, bias=bias, norm_layer=norm_layer,
            multiple_of=multiple_of, layernorm_eps=layernorm_eps, use_adaln=use_adaln)
        self.hparams_ = dict(n_genes=n_genes, n_embed=n_embed, n_embed_latent=n_embed_latent, n_head=n_head,
                             n_head_cross=n_head_cross, n_layer=n_layer, n_inducing_points=n_inducing_points, bias=bias,
                             multiple_of=multiple_of, layernorm_eps=layernorm_eps, shared_embedding=shared_embedding,
                             use_adaln=use_adaln)

    def forward(self, x, genes, condition=None):
        raise RuntimeError("Decoder is driven through TransformerVAE.decode (fused MCAB + NB-head kernel); no eager fallback")


class DiT(nn.Module):
    """Diffusion Transformer, drop-in for `scldm.nnets.DiT` (`nnets.py:216-492`)."""

    def __init__(self, n_embed: int, n_embed_input: int, n_layer: int, n_head: int, seq_len: int, dropout: float, bias: bool,
                 norm_layer: str, multiple_of: int, layernorm_eps: float, class_vocab_sizes: dict[str, int],
                 cfg_dropout_prob: float = 0.1, condition_strategy: Literal["mutually_exclusive", "joint"] = "mutually_exclusive"):
        super().__init__()
        self.class_vocab_sizes = dict(class_vocab_sizes)
        self.cfg_dropout_prob = cfg_dropout_prob
        self.condition_strategy = condition_strategy
        self.class_embeddings = nn.ModuleDict()
        for name, vocab in class_vocab_sizes.items():
            self.class_embeddings[name] = nn.Embedding(vocab + int(cfg_dropout_prob > 0), n_embed)
        self.t_embedder = TimestepEmbedder(n_embed)
        self.pos_embed = nn.Parameter(torch.zeros(1, seq_len, n_embed), requires_grad=False)
        self.blocks = nn.ModuleList([
            Block(n_embed=n_embed, n_head=n_head, dropout=dropout, bias=bias, norm_layer=norm_layer, multiple_of=multiple_of,
                  layernorm_eps=layernorm_eps, use_adaln=True, elementwise_affine=False) for _ in range(n_layer)])
        self.n_embed, self.seq_len = n_embed, seq_len
        self.input_proj = nn.Linear(n_embed_input, n_embed, bias=bias)
        self.final_layer = FinalLayerDit(n_embed, n_embed_input, bias, layernorm_eps)
        self.config = DiTConfig(n_embed=n_embed, n_embed_input=n_embed_input, n_layer=n_layer, n_head=n_head, seq_len=seq_len,
                                dropout=dropout, bias=bias, norm_layer=norm_layer, multiple_of=multiple_of,
                                layernorm_eps=layernorm_eps, class_vocab_sizes=dict(class_vocab_sizes),
                                cfg_dropout_prob=cfg_dropout_prob, condition_strategy=condition_strategy)
        self.dedup_conditions = True   # ODE path: one adaLN row per distinct label combination instead of one per cell
        self._packed: PackedDiT | None = None
        self._packed_key = None
        self.initialize_weights()

    # ---- initialisation with the reference's distributions (`nnets.py:458-492`) ----
    def initialize_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
        self.pos_embed.data.copy_(torch.from_numpy(get_1d_sincos_pos_embed(self.n_embed, self.seq_len)).float().unsqueeze(0))
        for emb in self.class_embeddings.values():
            nn.init.normal_(emb.weight, std=0.02)
        nn.init.normal_(self.t_embedder.mlp[0].weight, std=0.02)
        nn.init.normal_(self.t_embedder.mlp[2].weight, std=0.02)
        for block in self.blocks:  # adaLN-zero
            nn.init.zeros_(block.adaln_modulation[-1].weight)
            nn.init.zeros_(block.adaln_modulation[-1].bias)
        for lin in (self.final_layer.adaln_modulation[-1], self.final_layer.linear):
            nn.init.zeros_(lin.weight)
            if lin.bias is not None:
                nn.init.zeros_(lin.bias)

    # ---- packed-weight cache ----
    def packed(self) -> PackedDiT:
        key = weights_key(self, "_wkey_all")
        dev = self.pos_embed.device
        if dev.type != "cuda":
            raise RuntimeError("scldm_b200.DiT runs on CUDA only (no CPU fallback): call .cuda() first")
        if self._packed is None or self._packed_key != key:
            self._packed = PackedDiT({k: v.detach() for k, v in self.state_dict().items()}, self.config, dev)
            self._packed_key = key
        return self._packed

    # ---- label bookkeeping (host side, tiny int tensors) ----
    def _null(self, name: str) -> int:
        return self.class_vocab_sizes[name]

    def _cls_rows(self, labels: dict[str, torch.Tensor], n: int, device) -> torch.Tensor:
        """[n_class, n] embedding rows: given labels where present, the class's null token otherwise
        (sum over *all* tables, reference `nnets.py:403-426, 447-456`)."""
        names = sorted(self.class_vocab_sizes.keys())
        rows = []
        for name in names:
            if name in labels:
                rows.append(labels[name].to(device=device, dtype=torch.int32).reshape(n))
            else:
                rows.append(torch.full((n,), self._null(name), dtype=torch.int32, device=device))
        if not rows:
            return torch.zeros(0, n, dtype=torch.int32, device=device)
        return torch.stack(rows)

    def _active_labels(self, condition: dict[str, torch.Tensor], force_drop_ids: bool = False) -> dict[str, torch.Tensor]:
        """Which labels enter the embedding (`nnets.py:380-456`), with the reference's sequence of RNG calls on the labels' device:
        mutually_exclusive -> ONE class picked with `torch.randint` (even when only one is given), and with `force_drop_ids`
        (training) the picked class's labels are replaced by the null token where `torch.rand(B) < cfg_dropout_prob`;
        joint -> all given classes, all dropped together by one mask whenever the module is in training mode (the reference
        ignores `force_drop_ids` on this branch, `nnets.py:441-447`)."""
        avail = [n for n in sorted(self.class_vocab_sizes.keys()) if n in condition]
        if not avail:
            return {}
        first = condition[avail[0]]
        n, device = first.shape[0], first.device
        if self.condition_strategy == "joint":
            if not self.training:
                return {k: condition[k] for k in avail}
            drop = torch.rand(n, device=device) < self.cfg_dropout_prob
            return {k: torch.where(drop, torch.full_like(condition[k], self._null(k)), condition[k]) for k in avail}
        r = torch.randint(0, len(avail), (), device=device)      # drawn even for one class, as the reference (same RNG consumption)
        pick = int(r.item()) if len(avail) > 1 else 0
        if not self.training and force_drop_ids:
            raise AssertionError("force_drop_ids must be False when not training")   # nnets.py:399-400
        name = avail[pick]
        vals = condition[name]
        if force_drop_ids:
            drop = torch.rand(n, device=device) < self.cfg_dropout_prob
            vals = torch.where(drop, torch.full_like(vals, self._null(name)), vals)
        return {name: vals}

    def forward(self, x: torch.Tensor, t: torch.Tensor, condition: dict[str, torch.Tensor], force_drop_ids: bool | None = None):
        """`DiT.forward` (`nnets.py:273-297`).  In training mode the CFG label dropout of the reference is applied to the labels
        on the host side (the kernels only ever see label rows); no autograd graph is built (forward only)."""
        if force_drop_ids is None:
            force_drop_ids = self.training
        n = x.shape[0]
        trainer = getattr(self, "trainer", None)
        if self.training and torch.is_grad_enabled() and trainer is not None:
            # differentiable path (`scldm_b200.training.DiTTrainer`): activations are kept and `loss.backward()` runs the backward
            # kernels, leaving the gradients in the parameters' `.grad` (views of the trainer's flat buffer)
            from .training import differentiable_forward
            cls = self._cls_rows(self._active_labels(condition or {}, force_drop_ids), n, x.device)
            return differentiable_forward(trainer, x.contiguous().float(), t.float(), cls)
        packed = self.packed()
        cls = self._cls_rows(self._active_labels(condition or {}, force_drop_ids), n, x.device)
        plan = ops.DitPlan(packed, n_u=n, n_g=0, n_f=1, coef=[1.0], cls_idx=cls,
                           slot_mod=torch.arange(n, dtype=torch.int32, device=x.device), slot_mode="identity")
        return ops.dit_forward(plan, x.contiguous().float(), t.float())

    def cfg_layout(self, condition: dict[str, torch.Tensor] | None, cfg_scale: dict[str, float] | None, half: int, device,
                   shared_time: bool) -> dict:
        """Slot / conditioning-row layout of one `forward_with_cfg` evaluation (`nnets.py:336-378`); pure host
        logic (runs on any device, see tests/test_host_logic.py).

        states [0,half) are the unconditional first half (one slot each); states [half,2*half) are guided:
        n_f = 1 + n_cond consecutive slots (unconditional pass, then one pass per guidance term) combined with
        `coef`.  Returns n_f, coef, cls_idx [n_class, n_mod], slot_mod [n_slots] and t_index (which input row's
        time every conditioning row uses; None when `shared_time`: all rows share t and 1 + half*n_cond rows suffice)."""
        passes: list[dict[str, torch.Tensor]] = []
        coef = [1.0]
        if condition is not None and cfg_scale is not None:
            second = {k: v[half:] for k, v in condition.items()}
            if self.condition_strategy == "joint":
                avg = sum(cfg_scale.values()) / len(cfg_scale)
                passes, coef = [second], [1.0 - avg, avg]
            else:
                passes = [{name: second[name]} for name in cfg_scale]
                coef = [1.0 - sum(cfg_scale.values())] + [float(s) for s in cfg_scale.values()]
        n_f = 1 + len(passes)
        null_rows = lambda n: self._cls_rows({}, n, device)  # noqa: E731
        ar = torch.arange(half, dtype=torch.int32, device=device)
        slot_mode = "identity"
        if shared_time and self.dedup_conditions and passes:
            # all rows share t, so the adaLN vectors depend on the label combination only: row 0 = unconditional, then
            # one row per distinct combination (14 for the dentate clusters instead of one per cell and pass)
            cols = torch.cat([self._cls_rows(p, half, device) for p in passes], 1)       # [n_class, n_c*half], pass-major
            uniq, inv = self._label_combinations(cols)
            cls_idx = torch.cat([null_rows(1), uniq.to(torch.int32)], 1)
            slot_u = torch.zeros(half, dtype=torch.int32, device=device)
            slot_g = torch.zeros(half, n_f, dtype=torch.int32, device=device)
            slot_g[:, 1:] = 1 + inv.view(len(passes), half).T.to(torch.int32)
            t_index, slot_mode = None, "table"
        elif shared_time:
            # row 0 = unconditional; row 1 + j*n_c + k = guided cell j, conditional pass k
            n_c = len(passes)
            cls = [null_rows(1)] + ([torch.stack([self._cls_rows(p, half, device) for p in passes], 2).reshape(-1, half * n_c)] if n_c else [])
            cls_idx = torch.cat(cls, 1)
            slot_u = torch.zeros(half, dtype=torch.int32, device=device)
            slot_g = torch.zeros(half, n_f, dtype=torch.int32, device=device)
            for k in range(n_c):
                slot_g[:, 1 + k] = 1 + ar * n_c + k
            t_index, slot_mode = None, "cfg_shared"
        else:
            # rows [0,half): unconditional first half; then per guided cell j: n_f rows (uncond, cond passes)
            per = [null_rows(half)] + [self._cls_rows(p, half, device) for p in passes]  # each [n_class, half]
            cls_g = torch.stack(per, 2).reshape(-1, half * n_f)
            cls_idx = torch.cat([null_rows(half), cls_g], 1)
            slot_u = ar
            slot_g = half + ar[:, None] * n_f + torch.arange(n_f, dtype=torch.int32, device=device)[None, :]
            t_index = torch.cat([ar, (half + ar).repeat_interleave(n_f)]).long()
        slot_mod = torch.cat([slot_u, slot_g.reshape(-1)])
        return dict(n_u=half, n_g=half, n_f=n_f, coef=coef, cls_idx=cls_idx, slot_mod=slot_mod, t_index=t_index, slot_mode=slot_mode)

    def plans_are_sync_free(self) -> bool:
        """True when building a shared-time CFG plan never synchronises with the device (`_label_combinations` enumerates small
        label spaces; larger ones go through `torch.unique`)."""
        total = 1
        for v in self.class_vocab_sizes.values():
            total *= v + 1
        return total <= 256

    def _label_combinations(self, cols: torch.Tensor):
        """Distinct columns of `cols` [n_class, n] (embedding rows per class, null = vocab size) and the column -> combination map.
        Small label spaces (prod (V_c + 1) <= 256: dentate_gyrus 15, tabula_muris 17, hlca 51) are enumerated in
        mixed radix on the device - no `torch.unique`, hence no host synchronisation while a plan is built; larger ones use
        `torch.unique` (one sync)."""
        names = sorted(self.class_vocab_sizes.keys())
        radix = [self.class_vocab_sizes[n] + 1 for n in names]
        total = 1
        for r in radix:
            total *= r
        if cols.shape[0] == 0 or total > 256:
            return torch.unique(cols, dim=1, return_inverse=True)
        key = (tuple(radix), str(cols.device))
        if getattr(self, "_combo_cache_key", None) != key:
            ids = torch.arange(total, device=cols.device)
            table, div = [], 1
            for r in radix:
                table.append((ids // div) % r)
                div *= r
            self._combo_table, self._combo_cache_key = torch.stack(table).to(torch.int32), key
        inv, div = torch.zeros(cols.shape[1], dtype=torch.int64, device=cols.device), 1
        for c, r in enumerate(radix):
            inv = inv + cols[c].long() * div
            div *= r
        return self._combo_table, inv

    def forward_plan(self, condition: dict[str, torch.Tensor] | None, n: int, device):
        """Evaluation plan of a plain conditional `forward(x, t, condition, force_drop_ids=False)` on n cells that all share the
        time (an ODE drift call without guidance, BASELINE configs[0]): one slot per cell, one conditioning row per distinct label
        combination.  Returns None when the reference's forward is not a pure function of its inputs (several mutually-exclusive
        classes: `torch.randint` picks the active class anew at every call, `nnets.py:395`) - callers then loop on the host."""
        avail = [k for k in sorted(self.class_vocab_sizes.keys()) if condition and k in condition]
        if self.condition_strategy != "joint" and len(avail) > 1:
            return None
        cols = self._cls_rows({k: condition[k] for k in avail}, n, device)
        if cols.shape[0] == 0:
            return ops.DitPlan(self.packed(), n_u=n, n_g=0, n_f=1, coef=[1.0], cls_idx=cols[:, :1],
                               slot_mod=torch.zeros(n, dtype=torch.int32, device=device), slot_mode="table")
        uniq, inv = torch.unique(cols, dim=1, return_inverse=True)
        return ops.DitPlan(self.packed(), n_u=n, n_g=0, n_f=1, coef=[1.0], cls_idx=uniq.to(torch.int32), slot_mod=inv.to(torch.int32),
                           slot_mode="table")

    def cfg_plan(self, condition, cfg_scale, half: int, device, shared_time: bool):
        lay = self.cfg_layout(condition, cfg_scale, half, device, shared_time)
        plan = ops.DitPlan(self.packed(), n_u=lay["n_u"], n_g=lay["n_g"], n_f=lay["n_f"], coef=lay["coef"],
                           cls_idx=lay["cls_idx"], slot_mod=lay["slot_mod"], slot_mode=lay["slot_mode"])
        return plan, lay["t_index"]

    def forward_with_cfg(self, x: torch.Tensor, t: torch.Tensor, condition: dict[str, torch.Tensor] | None = None,
                         cfg_scale: dict[str, float] | None = None) -> torch.Tensor:
        """`DiT.forward_with_cfg` (`nnets.py:336-378`): rows [0,B) unconditional, rows [B,2B) guided; the
        unconditional and conditional passes of every guided cell are batched in one launch sequence."""
        half = x.shape[0] // 2
        plan, t_index = self.cfg_plan(condition, cfg_scale, half, x.device, shared_time=False)
        return ops.dit_forward(plan, x.contiguous().float(), t.float()[t_index])
