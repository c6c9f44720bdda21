"""`LatentDiffusion.sample` drop-in (`scldm/models.py:766-819`) without the Lightning shell.

The reference class is a LightningModule whose generation entry point is `sample(condition,
guidance_weight, batch_size, genes, timesteps)`; training/optimizer/EMA/logging glue is out of scope
(SURVEY.md §2a row 6).  This class keeps the attribute names (`vae_model`, `diffusion_model`,
`transport`, `transport_sampler`) and the `sample` signature/return (`(counts (2B,G), z (2B,M,L))`,
rows [0,B) unconditional, rows [B,2B) guided) and runs every stage on the GPU:

  size factors  table lookup + Philox normal (replaces the per-cell `.item()` loop, models.py:585-596)
  noise         Philox N(0,1) keyed by global cell index
  ODE + CFG     one C-ABI call per chunk of cells (all steps, both CFG branches batched)
  decode        fused latent blocks -> MCAB -> NB head -> Gamma-Poisson draw
"""

from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .nnets import DiT
from .transport import Sampler, Transport
from .transport.transport import FusedCFGModel
from .vae import TransformerVAE

STREAM_NOISE = 0x5A01
STREAM_SIZE_FACTOR = 0x5A02


class LatentDiffusion(nn.Module):
    def __init__(self, vae_model: TransformerVAE, diffusion_model: DiT, transport: Transport,
                 mu_size_factor: dict | None = None, sd_size_factor: dict | None = None,
                 size_factor_condition_key: str | None = None, sampling_method: str = "dopri5", num_steps: int = 50,
                 seed: int = 0, cell_chunk: int = 1184, joint_idx_2_classes: dict | None = None, joint_key: str | None = None,
                 joint_components: list | None = None, **_unused_training_kwargs):
        super().__init__()
        self.vae_model = vae_model
        self.diffusion_model = diffusion_model
        self.transport = transport
        self.transport_sampler = Sampler(transport)
        self.mu_size_factor, self.sd_size_factor = mu_size_factor, sd_size_factor
        self.size_factor_condition_key = size_factor_condition_key
        # joint statistics (`encoder.py:113-134`): "{i}_{j}" -> class index of the joint table `joint_key`
        self.joint_idx_2_classes, self.joint_key, self.joint_components = joint_idx_2_classes, joint_key, joint_components
        # `sample` in the reference always calls `sample_ode()` with its defaults (`models.py:793`, `transport.py:324-332`):
        # adaptive dopri5, 50 output points, atol = rtol = 1e-5.  That is the default here too; the fixed-grid solvers
        # ("euler" = BASELINE's 49-step configuration, "heun2", "midpoint") run the whole loop in one C-ABI call.
        self.sampling_method, self.num_steps = sampling_method, num_steps
        self.atol, self.rtol = 1e-5, 1e-5
        self.seed = seed
        self.cell_chunk = cell_chunk
        self.decode_piece = 296  # cells per decode launch when results stream to the host (`sample(host_out=...)`)
        self.cells_generated = 0  # global cell counter -> RNG offsets independent of batching / sharding
        self._sf_tables: dict = {}

    @property
    def device(self):
        return self.diffusion_model.pos_embed.device

    # ---- size factors (independent path of `_sample_log_size_factors`, models.py:552-597) ----
    def _size_factor_key(self, condition) -> str | None:
        if condition is None or self.mu_size_factor is None or self.sd_size_factor is None:
            return None
        k = self.size_factor_condition_key
        if k and k in condition and k in self.mu_size_factor and k in self.sd_size_factor:
            return k
        inter = sorted(set(condition) & set(self.mu_size_factor) & set(self.sd_size_factor))
        return inter[0] if inter else None

    def _use_joint(self, condition) -> bool:
        """joint branch of `_sample_log_size_factors` (`models.py:498-550`)."""
        return (condition is not None and self.mu_size_factor is not None and self.sd_size_factor is not None
                and getattr(self.diffusion_model, "condition_strategy", None) == "joint" and self.joint_idx_2_classes is not None
                and self.joint_key is not None and self.joint_key in self.mu_size_factor and self.joint_key in self.sd_size_factor)

    def _sample_joint_log_size_factors(self, condition, batch_size: int, cell_offset: int) -> torch.Tensor:
        comps = [k for k in self.joint_components if k in condition] if self.joint_components is not None else list(condition.keys())
        if "joint" not in self._sf_tables:
            # dense (component index tuple) -> (mu, sd) tables on the device; missing keys / stats give exactly 0
            sizes = [self.diffusion_model.class_vocab_sizes[k] + 1 for k in comps]
            n = 1
            for v in sizes:
                n *= v
            mu, sd = torch.zeros(n), torch.zeros(n)
            mu_vec, sd_vec = self.mu_size_factor[self.joint_key], self.sd_size_factor[self.joint_key]
            for key, cls_idx in self.joint_idx_2_classes.items():
                idx = [int(v) for v in key.split("_")]
                if len(idx) != len(sizes) or any(i >= s for i, s in zip(idx, sizes)):
                    continue
                m, s_ = mu_vec.get(cls_idx), sd_vec.get(cls_idx)
                if m is None or s_ is None:
                    continue
                flat = 0
                for i, sz in zip(idx, sizes):
                    flat = flat * sz + i
                mu[flat], sd[flat] = float(m), float(s_)
            self._sf_tables["joint"] = (mu.to(self.device), sd.to(self.device), sizes, comps)
        mu, sd, sizes, comps = self._sf_tables["joint"]
        flat = torch.zeros(batch_size, dtype=torch.long, device=self.device)
        for k, sz in zip(comps, sizes):
            flat = flat * sz + condition[k].to(self.device).long()
        eps = ops.randn_cells(batch_size, 1, self.seed, cell_offset, STREAM_SIZE_FACTOR, self.device).reshape(-1)
        return mu[flat] + sd[flat] * eps

    def _sample_log_size_factors(self, condition, batch_size: int, cell_offset: int) -> torch.Tensor:
        if self._use_joint(condition):
            return self._sample_joint_log_size_factors(condition, batch_size, cell_offset)
        key = self._size_factor_key(condition)
        if key is None:
            return torch.zeros(batch_size, device=self.device)
        if key not in self._sf_tables:
            vocab = self.diffusion_model.class_vocab_sizes.get(key, max(self.mu_size_factor[key]) + 1)
            mu = torch.zeros(vocab + 1)
            sd = torch.zeros(vocab + 1)  # classes without statistics draw exactly 0, as the reference
            for c, v in self.mu_size_factor[key].items():
                if c in self.sd_size_factor[key]:
                    mu[c], sd[c] = float(v), float(self.sd_size_factor[key][c])
            self._sf_tables[key] = (mu.to(self.device), sd.to(self.device))
        mu, sd = self._sf_tables[key]
        eps = ops.randn_cells(batch_size, 1, self.seed, cell_offset, STREAM_SIZE_FACTOR, self.device).reshape(-1)
        lab = condition[key].to(self.device).long()
        return mu[lab] + sd[lab] * eps

    @torch.no_grad()
    def sample(self, condition: dict[str, torch.Tensor] | None, guidance_weight: dict[str, float] | None, batch_size: int,
               genes: torch.Tensor, timesteps: int = 50, *, z0: torch.Tensor | None = None,
               log_size_factors: torch.Tensor | None = None, return_mu: bool = False, cell_offset: int | None = None,
               host_out: tuple[torch.Tensor, torch.Tensor] | None = None):
        """Generation (`models.py:766-819`).  `timesteps` is accepted and ignored exactly as in the reference
        (`models.py:773,793`); the grid is `self.num_steps` points.  `z0` / `log_size_factors` may be injected
        (parity tests share them with the oracle); otherwise they are drawn on device.

        `host_out=(counts_host (2B,G), z_host (2B,M,L))`, pinned: every chunk's rows are copied to the host on a side
        stream as soon as they are decoded, so the device-to-host transfer (the reference's `.cpu()` in
        `predict_step`, `models.py:742`) overlaps the ODE of the next chunk; the copies are complete when this call's
        stream is synchronised (the side stream is joined before returning)."""
        if len(genes) != batch_size:
            raise ValueError(f"genes batch dimension ({genes.shape[0]}) must match batch_size ({batch_size})")
        if condition is not None:
            for key, values in condition.items():
                if len(values) != batch_size:
                    raise ValueError(f"Condition '{key}' length ({len(values)}) must match batch size ({batch_size})")
        if guidance_weight is not None and condition is not None:
            assert set(guidance_weight.keys()) == set(condition.keys()), (
                f"Guidance weight keys {set(guidance_weight.keys())} must match condition keys {set(condition.keys())}")
        dev = self.device
        dit = self.diffusion_model
        offset = self.cells_generated if cell_offset is None else cell_offset  # global index of cell 0 (RNG key)
        self.cells_generated = offset + batch_size
        lsf = log_size_factors.to(dev).float() if log_size_factors is not None else \
            self._sample_log_size_factors(condition, batch_size, offset)
        if z0 is None:
            z0 = ops.randn_cells(batch_size, dit.seq_len * dit.config.n_embed_input, self.seed, offset, STREAM_NOISE, dev)
            z0 = z0.view(batch_size, dit.seq_len, dit.config.n_embed_input)
        z0 = z0.to(dev).float()
        sample_fn = self.transport_sampler.sample_ode(sampling_method=self.sampling_method, num_steps=self.num_steps,
                                                      atol=self.atol, rtol=self.rtol)
        model_fn = FusedCFGModel(dit, guidance_weight)
        cond = {k: v.to(dev) for k, v in (condition or {}).items()}
        from .vae import shared_gene_vector

        gvec = shared_gene_vector(genes).to(device=dev, dtype=torch.int64).contiguous()   # one row; differing rows are rejected loudly
        lib = torch.exp(lsf)
        G = gvec.numel()
        counts = torch.empty(2 * batch_size, G, dtype=torch.float32, device=dev)
        mu_out = torch.empty(2 * batch_size, G, dtype=torch.float32, device=dev) if return_mu else None
        z_out = torch.empty(2 * batch_size, dit.seq_len, dit.config.n_embed_input, dtype=torch.float32, device=dev)
        copy_stream = None
        if host_out is not None:
            counts_host, z_host = host_out
            if tuple(counts_host.shape) != (2 * batch_size, G) or tuple(z_host.shape) != tuple(z_out.shape):
                raise ValueError("host_out shapes must be (2B, G) and (2B, seq_len, n_embed_input)")
            if not (counts_host.is_pinned() and z_host.is_pinned()):
                raise ValueError("host_out buffers must be pinned host memory")
            if not hasattr(self, "_copy_stream"):
                self._copy_stream = torch.cuda.Stream(device=dev)
            copy_stream = self._copy_stream
            copy_stream.wait_stream(torch.cuda.current_stream(dev))   # earlier readers of the host buffers are ordered before us
        # cells are independent: run the ODE + decode chunk by chunk.  A chunk's evaluation plan is built right before its solve is
        # queued, i.e. (from the second chunk on) while the GPU is still busy with the previous chunk - unless deduplicating the
        # label combinations has to synchronise with the device (large label spaces, `torch.unique`): then all plans are built
        # first, because a synchronisation between chunks would leave the GPU idle while the host prepares the next launches.
        chunks = [(c0, min(c0 + self.cell_chunk, batch_size)) for c0 in range(0, batch_size, self.cell_chunk)]
        fixed = self.sampling_method.lower() in Sampler.FIXED

        def make(c0, c1):
            cc = {k: torch.cat([v[c0:c1], v[c0:c1]]) for k, v in cond.items()}
            plan = dit.cfg_plan(cc if cond else None, guidance_weight, c1 - c0, dev, shared_time=True)[0] if fixed else None
            return cc, plan

        ahead = None if dit.plans_are_sync_free() else [make(c0, c1) for c0, c1 in chunks]
        for i, (c0, c1) in enumerate(chunks):
            cc, plan = ahead[i] if ahead is not None else make(c0, c1)
            zc = z0[c0:c1]
            zf = sample_fn(torch.cat([zc, zc]), model_fn, condition=cc, **({"_plan": plan} if plan is not None else {}))[-1]
            n = c1 - c0
            libc = torch.cat([lib[c0:c1], lib[c0:c1]])
            # decode in pieces; with host_out every piece is copied out behind its decode, so only the last piece's transfer
            # is left when the final chunk finishes
            piece = self.decode_piece if copy_stream is not None else n
            for half in (0, 1):
                for p0 in range(0, n, piece):
                    p1 = min(p0 + piece, n)
                    rows = slice(half * batch_size + c0 + p0, half * batch_size + c0 + p1)
                    zh = zf[half * n + p0:half * n + p1]
                    self.vae_model.decode_counts(zh, gvec, libc[half * n + p0:half * n + p1], seed=self.seed,
                                                 cell_offset=offset + c0 + p0 + half * (1 << 40), want_mu=return_mu,
                                                 out_counts=counts[rows], out_mu=mu_out[rows] if return_mu else None)
                    z_out[rows] = zh
                    if copy_stream is not None:
                        ev = torch.cuda.Event()
                        ev.record(torch.cuda.current_stream(dev))
                        copy_stream.wait_event(ev)
                        with torch.cuda.stream(copy_stream):
                            counts_host[rows].copy_(counts[rows], non_blocking=True)
                            z_host[rows].copy_(z_out[rows], non_blocking=True)
        if copy_stream is not None:
            torch.cuda.current_stream(dev).wait_stream(copy_stream)
        if return_mu:
            return counts, z_out, mu_out
        return counts, z_out

    def sample_csr(self, condition, guidance_weight, batch_size: int, genes: torch.Tensor, timesteps: int = 50, **kw):
        """`sample` with the count matrix sparsified on the device: returns `((indptr, indices, data), z)` where the triple is
        what `scipy.sparse.csr_matrix(counts)` holds (the reference builds it on the host from the dense copy in
        `process_generation_output`, `_utils.py:186-200`).  Use `csr_to_scipy` for the host object."""
        counts, z = self.sample(condition, guidance_weight, batch_size, genes, timesteps, **kw)
        return ops.counts_to_csr(counts), z


def csr_to_scipy(csr: tuple[torch.Tensor, torch.Tensor, torch.Tensor], n_genes: int):
    """Host `scipy.sparse.csr_matrix` from the device arrays of `LatentDiffusion.sample_csr` (the D2H copy moves
    8 bytes per non-zero instead of 4 bytes per matrix entry)."""
    from scipy import sparse

    indptr, indices, data = (t.cpu().numpy() for t in csr)
    return sparse.csr_matrix((data, indices, indptr), shape=(len(indptr) - 1, n_genes))
