"""LDM training step on the device (reference: `LatentDiffusion.training_step`, `src/scldm/models.py:634-666`;
`Transport.training_losses`, `transport/transport.py:110-150`; `configure_optimizers`, `models.py:599-610` with
`torch.optim.AdamW` from `experiments/configs/model/ldm_base.yaml:36-40`; `gradient_clip_val: 10` from
`configs/training/default.yaml:15`; `wsd_schedule`, `_utils.py:19-60`; DDP, `experiments/scripts/train_ldm.py:101`).

`DiTTrainer` owns ONE flat fp32 parameter buffer (the DiT's `nn.Parameter`s become views into it, so `state_dict()` is
unchanged), a flat gradient buffer with the same layout, the AdamW moments and the bf16 UMMA-tile copies of the GEMM
weights.  Forward / backward / optimizer all run in hand-written sm_100a kernels through the C-ABI
(`scldm_dit_train_forward`, `scldm_dit_train_backward`, `scldm_adamw_step`); there is no eager fallback.

Data parallelism: one process per GPU.  The flat gradient buffer is ordered by completion time in the backward pass
(final layer, blocks L-1..0, conditioning / input tensors); the backward call records a CUDA event after the last block of
every bucket, and the bucket's NCCL all-reduce is enqueued on a side stream behind that event, so it overlaps the backward
of the earlier blocks (what torch DDP's bucketed all-reduce does for the reference).
"""

from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib
from .pack import BLOCK_K, _swizzle_rows

MAX_LAYERS = 32


def _pack_index_tiles(idx: torch.Tensor, block_rows: int = 256) -> torch.Tensor:
    """The permutation `pack.pack_kmajor_tiles` applies, on an int64 index matrix (padding = -1)."""
    N, K = idx.shape
    nt, ks = -(-N // block_rows), -(-K // BLOCK_K)
    wp = torch.full((nt * block_rows, ks * BLOCK_K), -1, dtype=torch.int64)
    wp[:N, :K] = idx
    tiles = wp.view(nt, block_rows, ks, BLOCK_K).permute(0, 2, 1, 3).contiguous()
    return _swizzle_rows(tiles).reshape(-1)


def wsd_schedule(num_training_steps, final_lr_factor=0.1, num_warmup_steps=1000, init_div_factor=100, fract_decay=0.1,
                 decay_type="cosine"):
    """Warm-up / hold / decay LR factor, restating `scldm._utils.wsd_schedule` (`_utils.py:19-60`)."""
    n_anneal = int(fract_decay * num_training_steps)
    n_hold = num_training_steps - n_anneal
    num_warmup_steps = num_warmup_steps or 0

    def schedule(step):
        if step < num_warmup_steps:
            return (step / num_warmup_steps) + (1 - step / num_warmup_steps) / init_div_factor
        if step < n_hold:
            return 1.0
        if step < num_training_steps:
            if decay_type == "cosine":
                prog = (step - num_warmup_steps) / (num_training_steps - num_warmup_steps)
                return final_lr_factor + (1 - final_lr_factor) * 0.5 * (1 + math.cos(math.pi * prog))
            if decay_type == "sqrt":
                return final_lr_factor + (1 - final_lr_factor) * (1 - math.sqrt((step - n_hold) / n_anneal))
            raise ValueError(f"decay type {decay_type} is not in ['cosine','sqrt']")
        return final_lr_factor

    return schedule


PER_LAYER = [("attn.c_attn.weight", "w_qkv"), ("attn.c_attn.bias", "b_qkv"), ("attn.c_proj.weight", "w_proj"), ("attn.c_proj.bias", "b_proj"),
             ("mlp.w1.weight", "w1"), ("mlp.w2.weight", "w2"), ("mlp.c_proj.weight", "w3"), ("adaln_modulation.1.weight", "w_mod")]


def flat_layout(cfg, shapes: dict[str, tuple]):
    """Order and element offsets of the trainable DiT tensors inside the flat buffers: final layer, blocks L-1..0 (a gradient is
    complete once its block has been differentiated), then the tensors whose gradients only complete at the end (adaLN biases -
    contiguous in block order, because ONE GEMM computes all modulation vectors -, timestep MLP, class tables, input projection).
    Returns (names in order, offsets, layer_end {block: end of the flat prefix that is final after that block}, n_params)."""
    L = cfg.n_layer
    names = ["final_layer.linear.weight", "final_layer.linear.bias", "final_layer.adaln_modulation.1.weight"]
    for l in range(L - 1, -1, -1):
        names += [f"blocks.{l}.{n}" for n, _ in PER_LAYER]
    names += [f"blocks.{l}.adaln_modulation.1.bias" for l in range(L)] + ["final_layer.adaln_modulation.1.bias"]
    names += ["t_embedder.mlp.0.weight", "t_embedder.mlp.0.bias", "t_embedder.mlp.2.weight", "t_embedder.mlp.2.bias"]
    names += [f"class_embeddings.{n}.weight" for n in sorted(cfg.class_vocab_sizes.keys())]
    names += ["input_proj.weight", "input_proj.bias"]
    missing = [n for n in names if n not in shapes]
    extra = [n for n in shapes if n not in set(names)]
    if missing or extra:
        raise NotImplementedError(f"unexpected DiT parameter set (missing {missing}, extra {extra})")
    off, offsets, layer_end = 0, {}, {}
    for name in names:
        offsets[name] = off
        off += (math.prod(shapes[name]) + 3) // 4 * 4
        if name.startswith("blocks.") and name.endswith("adaln_modulation.1.weight"):
            layer_end[int(name.split(".")[1])] = off
    return names, offsets, layer_end, off


def pack_sources(cfg, offsets: dict[str, int], shapes: dict[str, tuple]):
    """For every element of the bf16 tile arena (`pk_qkv | pk_proj | pk_w12 | pk_w3 | pk_mod`, layouts in include/scldm_b200.h) the
    flat-buffer index it is a copy of (-1: zero padding).  Returns (src int64 [arena], {key: arena offset})."""
    L, D, H = cfg.n_layer, cfg.n_embed, cfg.hidden
    T = -(-H // 128)

    def idx_of(name):
        return (offsets[name] + torch.arange(math.prod(shapes[name]), dtype=torch.int64)).view(shapes[name])

    arena, pk_off = [], {}

    def add(key, idx_flat):
        pk_off[key] = sum(a.numel() for a in arena)
        arena.append(idx_flat)

    add("qkv", torch.cat([_pack_index_tiles(idx_of(f"blocks.{l}.attn.c_attn.weight")) for l in range(L)]))
    add("proj", torch.cat([_pack_index_tiles(idx_of(f"blocks.{l}.attn.c_proj.weight")) for l in range(L)]))
    w12 = []
    for l in range(L):
        w1 = torch.full((T * 128, D), -1, dtype=torch.int64)
        w2 = torch.full((T * 128, D), -1, dtype=torch.int64)
        w1[:H], w2[:H] = idx_of(f"blocks.{l}.mlp.w1.weight"), idx_of(f"blocks.{l}.mlp.w2.weight")
        # N tile j = [w1 rows 128j.. | w2 rows 128j..]: a SwiGLU pair shares an accumulator tile (as pack.PackedDiT.w_mlp1, unhalved)
        w12.append(_pack_index_tiles(torch.stack([w1.view(T, 128, D), w2.view(T, 128, D)], 1).reshape(-1, D)))
    add("w12", torch.cat(w12))
    w3 = []
    for l in range(L):
        m = torch.full((D, T * 128), -1, dtype=torch.int64)
        m[:, :H] = idx_of(f"blocks.{l}.mlp.c_proj.weight")
        w3.append(_pack_index_tiles(m))
    add("w3", torch.cat(w3))
    add("mod", _pack_index_tiles(torch.cat([idx_of(f"blocks.{l}.adaln_modulation.1.weight") for l in range(L)]
                                           + [idx_of("final_layer.adaln_modulation.1.weight")], 0)))
    return torch.cat(arena), pk_off


class DiTTrainer:
    """Flat-buffer training state of a `scldm_b200.nnets.DiT` living on a CUDA device."""

    def __init__(self, dit, *, lr=5e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_grad_norm=10.0, lr_lambda=None,
                 process_group=None, n_buckets=3, ema_decay=None, ema_update_every=10, ema_update_after_step=10_000):
        cfg = dit.config
        dev = dit.pos_embed.device
        if dev.type != "cuda":
            raise RuntimeError("DiTTrainer needs the DiT on a CUDA device (there is no CPU training path)")
        if (cfg.n_embed, cfg.seq_len, cfg.n_embed_input, cfg.n_head) != (256, 16, 16, 8) or not cfg.bias:
            raise NotImplementedError(f"sm_100a training kernels cover n_embed=256, seq_len=16, n_embed_input=16, n_head=8, bias=True; got {cfg}")
        if cfg.n_layer > MAX_LAYERS or cfg.dropout != 0.0:
            raise NotImplementedError("n_layer <= 32 and dropout == 0 (ldm_base.yaml:19-21)")
        self.lib = _lib.load()
        self.dit, self.cfg, self.device = dit, cfg, dev
        self.lr, self.betas, self.eps, self.weight_decay, self.max_grad_norm = lr, betas, eps, weight_decay, max_grad_norm
        self.lr_lambda = lr_lambda
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if self._dist() else 1
        self.step_count = 0
        L, D, H = cfg.n_layer, cfg.n_embed, cfg.hidden
        self.T = -(-H // 128)
        self.class_names = sorted(cfg.class_vocab_sizes.keys())

        # ---- flat layout, ordered by gradient completion in the backward pass ----
        sd = dict(dit.named_parameters())
        shapes = {k: tuple(v.shape) for k, v in sd.items() if v.requires_grad}
        order, offsets, self.layer_end, self.n_params = flat_layout(cfg, shapes)
        off = self.n_params
        self.offsets = offsets
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
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
        self.pos = dit.pos_embed.detach().reshape(cfg.seq_len, D).to(torch.float32).contiguous()

        # ---- bf16 UMMA tiles of the GEMM weights: one arena, `pk_dst[i]` = arena position of parameter i ----
        src, self.pk_off = pack_sources(cfg, offsets, shapes)
        pk_dst = torch.full((self.n_params,), -1, dtype=torch.int32)
        pos = torch.nonzero(src >= 0).reshape(-1)
        pk_dst[src[pos]] = pos.to(torch.int32)
        self.pk_dst = pk_dst.to(dev)
        self.pk = torch.zeros(src.numel(), dtype=torch.bfloat16, device=dev)

        # ---- C struct ----
        s = _lib.DitTrain()
        s.n_layer, s.hidden, s.n_class, s.eps = L, H, len(self.class_names), float(cfg.layernorm_eps)
        lo = s.off
        for l in range(L):
            for n, f in PER_LAYER:
                getattr(lo, f)[l] = offsets[f"blocks.{l}.{n}"]
        lo.b_mod = offsets["blocks.0.adaln_modulation.1.bias"]
        for l in range(L):   # the adaLN biases must be contiguous in block order (one bias vector for the one modulation GEMM)
            assert offsets[f"blocks.{l}.adaln_modulation.1.bias"] == lo.b_mod + l * 6 * D
        assert offsets["final_layer.adaln_modulation.1.bias"] == lo.b_mod + L * 6 * D
        for f, n in (("w_mod_final", "final_layer.adaln_modulation.1.weight"), ("w_out", "final_layer.linear.weight"),
                     ("b_out", "final_layer.linear.bias"), ("temb_w0", "t_embedder.mlp.0.weight"), ("temb_b0", "t_embedder.mlp.0.bias"),
                     ("temb_w2", "t_embedder.mlp.2.weight"), ("temb_b2", "t_embedder.mlp.2.bias"), ("w_in", "input_proj.weight"),
                     ("b_in", "input_proj.bias")):
            setattr(lo, f, offsets[n])
        for i, n in enumerate(self.class_names):
            lo.class_tab[i] = offsets[f"class_embeddings.{n}.weight"]
        lo.n_params = self.n_params
        s.params, s.grads, s.pos = self.flat.data_ptr(), self.grad.data_ptr(), self.pos.data_ptr()
        for key in ("qkv", "proj", "w12", "w3", "mod"):
            setattr(s, "pk_" + key, self.pk.data_ptr() + 2 * self.pk_off[key])
        self.struct = s
        self.repack()

        # ---- gradient buckets (flat prefixes ending after a block) + their events / side stream ----
        n_buckets = max(1, min(n_buckets, L))
        cut_layers = sorted({(L * (n_buckets - 1 - b)) // n_buckets for b in range(n_buckets - 1)}, reverse=True)   # e.g. L=8, 3 buckets -> after blocks 5, 2
        self.bucket_layers = [l for l in cut_layers if 0 < l < L]
        self.bucket_bounds = [0] + [self.layer_end[l] for l in self.bucket_layers] + [self.n_params]
        self.events = [torch.cuda.Event() for _ in self.bucket_layers]
        self.comm_stream = torch.cuda.Stream(device=dev)
        self._ws = None
        self._ws_cells = 0
        self.ema = None
        self.ema_decay, self.ema_update_every, self.ema_update_after_step = ema_decay, ema_update_every, ema_update_after_step
        self._ema_step, self._ema_initted = 0, False
        if ema_decay is not None:
            self.ema = self.flat.clone()
        self.last_allreduce_ms = None

    # ------------------------------------------------------------------------------------------------------------
    def _dist(self) -> bool:
        return torch.distributed.is_available() and torch.distributed.is_initialized()

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def zero_grad(self) -> None:
        self.grad.zero_()

    def repack(self) -> None:
        """Refresh the bf16 GEMM tiles from the fp32 master weights (after construction / `load_state_dict`)."""
        with torch.cuda.device(self.device):
            rc = self.lib.scldm_repack(self.flat.data_ptr(), self.pk_dst.data_ptr(), self.pk.data_ptr(), self.n_params, self._stream())
        _lib.check(rc, "scldm_repack")

    def workspace(self, n_cells: int) -> torch.Tensor:
        if self._ws is None or self._ws_cells != n_cells:
            nbytes = int(self.lib.scldm_dit_train_workspace_bytes(C.byref(self.struct), n_cells))
            self._ws = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
            self._ws_cells = n_cells
        return self._ws

    def cls_rows(self, labels: dict[str, torch.Tensor], n: int) -> torch.Tensor:
        return self.dit._cls_rows(labels, n, self.device).contiguous()

    def forward(self, x: torch.Tensor, t: torch.Tensor, cls_idx: torch.Tensor) -> torch.Tensor:
        """v = DiT(x, t, labels) with activations saved for `backward`.  x [B,16,16] fp32, t [B], cls_idx [n_class,B] int32."""
        B = x.shape[0]
        x = x.contiguous().float()
        t = t.contiguous().float()
        cls_idx = cls_idx.contiguous().to(torch.int32)
        v = torch.empty_like(x)
        ws = self.workspace(B)
        with torch.cuda.device(self.device):
            rc = self.lib.scldm_dit_train_forward(C.byref(self.struct), x.data_ptr(), t.data_ptr(), cls_idx.data_ptr() if cls_idx.numel() else None, B,
                                                  v.data_ptr(), ws.data_ptr(), ws.numel(), self._stream())
        _lib.check(rc, "scldm_dit_train_forward")
        return v

    def backward(self, dv: torch.Tensor, *, need_dx: bool = False, zero_grads: bool = True, with_events: bool = True):
        """Backward of the last `forward`: parameter gradients into the flat buffer (`p.grad` views); returns dLoss/dx or None."""
        B = dv.shape[0]
        dv = dv.contiguous().float()
        dx = torch.empty_like(dv) if need_dx else None
        ws = self.workspace(B)
        n_ev = len(self.events) if with_events else 0
        ev_arr = (C.c_void_p * max(n_ev, 1))(*[e.cuda_event for e in self._fresh_events()][:n_ev]) if n_ev else None
        lay_arr = (C.c_int32 * max(n_ev, 1))(*self.bucket_layers[:n_ev]) if n_ev else None
        with torch.cuda.device(self.device):
            rc = self.lib.scldm_dit_train_backward(C.byref(self.struct), dv.data_ptr(), dx.data_ptr() if need_dx else None, B, int(zero_grads),
                                                   ev_arr, lay_arr, n_ev, ws.data_ptr(), ws.numel(), self._stream())
        _lib.check(rc, "scldm_dit_train_backward")
        return dx

    def _fresh_events(self):
        for e in self.events:
            e.record(torch.cuda.current_stream(self.device))   # forces creation of the underlying cudaEvent_t (lazily created by torch)
        return self.events

    def allreduce_grads(self) -> None:
        """Sum the flat gradient over the data-parallel ranks, bucket by bucket on the side stream; bucket b starts as soon as the
        backward pass has produced it (`backward(with_events=True)`).  The mean is taken by the optimizer's `grad_scale`."""
        if not self._dist() or self.world == 1:
            return
        cur = torch.cuda.current_stream(self.device)
        nb = len(self.bucket_bounds) - 1
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(self.comm_stream):
            for b in range(nb):
                if b < len(self.events):
                    self.comm_stream.wait_event(self.events[b])
                else:
                    self.comm_stream.wait_stream(cur)
                if b == 0:
                    t0.record(self.comm_stream)
                torch.distributed.all_reduce(self.grad[self.bucket_bounds[b]: self.bucket_bounds[b + 1]], group=self.pg)
            t1.record(self.comm_stream)
        cur.wait_stream(self.comm_stream)
        self._ar_events = (t0, t1)

    def optimizer_step(self) -> None:
        """Gradient clipping by global norm + AdamW + refresh of the bf16 tiles (+ EMA), one fused pass over the flat buffers."""
        self.step_count += 1
        lr = self.lr * (self.lr_lambda(self.step_count - 1) if self.lr_lambda is not None else 1.0)
        with torch.cuda.device(self.device):
            rc = self.lib.scldm_adamw_step(self.flat.data_ptr(), self.grad.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), self.n_params,
                                           float(lr), float(self.betas[0]), float(self.betas[1]), float(self.eps), float(self.weight_decay),
                                           self.step_count, float(self.max_grad_norm or 0.0), 1.0 / self.world, self.scratch.data_ptr(),
                                           self.pk_dst.data_ptr(), self.pk.data_ptr(), self._stream())
        _lib.check(rc, "scldm_adamw_step")
        self.dit._packed = None            # the inference-side packed weights are stale now (rebuilt lazily by DiT.packed())
        if self.ema is not None:
            self._ema_update()

    def _ema_update(self) -> None:
        """`ema_pytorch.EMA.update` (0.7.7, restated - the package is not installable here): copy until `update_after_step`, then
        every `update_every` steps lerp with decay = clamp(1 - (1 + epoch)^(-2/3), 0, beta), epoch = step - update_after_step - 1."""
        step = self._ema_step
        self._ema_step += 1
        if step % self.ema_update_every != 0:
            return
        if step <= self.ema_update_after_step or not self._ema_initted:
            self.ema.copy_(self.flat)
            self._ema_initted = True
            return
        epoch = max(step - self.ema_update_after_step - 1, 0)
        decay = 0.0 if epoch <= 0 else min(max(1 - (1 + epoch) ** (-2.0 / 3.0), 0.0), self.ema_decay)
        with torch.cuda.device(self.device):
            rc = self.lib.scldm_ema_update(self.ema.data_ptr(), self.flat.data_ptr(), self.n_params, float(decay), self._stream())
        _lib.check(rc, "scldm_ema_update")

    def ema_state_dict(self) -> dict:
        """EMA weights under the DiT's `state_dict` names (what `ema_model` holds in a reference checkpoint)."""
        out = {k: v.detach().clone() for k, v in self.dit.state_dict().items()}
        if self.ema is not None:
            for name, off in self.offsets.items():
                out[name] = self.ema[off: off + out[name].numel()].view(out[name].shape).clone()
        return out

    # ------------------------------------------------------------------------------------------------------------
    def fm_step(self, z: torch.Tensor, condition: dict[str, torch.Tensor] | None, transport, *, t=None, x0=None) -> torch.Tensor:
        """One flow-matching training step on latents z [B,16,16] (`models.py:656-666`): loss = mean_cells mean_flat((v - u)^2),
        backward, DDP all-reduce, clip, AdamW.  Returns the (local) mean loss as a 0-d device tensor (no host sync)."""
        B = z.shape[0]
        if t is None or x0 is None:
            ts, x0s, _ = transport.sample(z)
            t = ts if t is None else t
            x0 = x0s if x0 is None else x0
        te = t.view(-1, 1, 1)
        xt = te * z + (1 - te) * x0
        ut = z - x0
        labels = self.dit._active_labels(condition or {}, force_drop_ids=self.dit.training)
        v = self.forward(xt, t, self.cls_rows(labels, B))
        diff = v - ut
        loss = (diff * diff).flatten(1).mean(1).mean()
        dv = diff * (2.0 / (B * diff[0].numel()))
        self.backward(dv)
        self.allreduce_grads()
        self.optimizer_step()
        return loss


class _DiTTrainFunction(torch.autograd.Function):
    """autograd bridge: `DiT.forward` in training mode -> `DiTTrainer.forward`; `.backward()` -> `DiTTrainer.backward`.
    Parameter gradients land in the flat buffer (the `p.grad` views), as `loss.backward()` would leave them."""

    @staticmethod
    def forward(ctx, x, t, cls_idx, trainer):
        ctx.trainer = trainer
        ctx.need_dx = x.requires_grad
        return trainer.forward(x.detach(), t.detach(), cls_idx)

    @staticmethod
    def backward(ctx, dv):
        dx = ctx.trainer.backward(dv, need_dx=ctx.need_dx, zero_grads=False, with_events=False)
        return dx, None, None, None


def differentiable_forward(trainer: DiTTrainer, x, t, cls_idx):
    # a dummy requires-grad input makes autograd call backward even when x itself needs no gradient
    if not x.requires_grad:
        x = x.detach().requires_grad_(True)
    return _DiTTrainFunction.apply(x, t, cls_idx, trainer)
