"""Thin torch-tensor wrappers over the C-ABI calls.  PyTorch supplies device memory and the
current stream; all arithmetic happens in libscldm_b200.so.  No CPU path exists."""

from __future__ import annotations

import ctypes as C
import functools

import torch

from . import _lib
from .pack import PackedDiT, PackedVAEDecoder, PackedVAEEncoder


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"scldm_b200: `{name}` must be a CUDA tensor (there is no CPU fallback)")
    if not t.is_contiguous():
        raise RuntimeError(f"scldm_b200: `{name}` must be contiguous")


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _cuda_device_of(a):
    if isinstance(a, torch.Tensor):
        return a.device if a.is_cuda else None
    d = getattr(a, "device", None)                       # PackedDiT / PackedVAE*
    if d is None and hasattr(a, "packed"):
        d = a.packed.device                              # DitPlan
    if d is None and isinstance(a, (str, torch.device)):
        try:
            d = torch.device(a)
        except (RuntimeError, TypeError):
            return None
    return d if isinstance(d, torch.device) and d.type == "cuda" else None


def _on_arg_device(fn):
    """Run the wrapped op with the CUDA device of its tensors / packed weights current: kernel launches, `cudaFuncSetAttribute` and
    the SM count all belong to the CURRENT device, which need not be the tensors' device in a multi-GPU process.  Mixing devices in
    one call is an error, not undefined behaviour."""
    @functools.wraps(fn)
    def wrapper(*args, **kw):
        devs = [d for d in (_cuda_device_of(a) for a in list(args) + list(kw.values())) if d is not None]
        devs = [d if d.index is not None else torch.device("cuda", torch.cuda.current_device()) for d in devs]
        if not devs:
            return fn(*args, **kw)
        if any(d != devs[0] for d in devs):
            raise RuntimeError(f"scldm_b200.{fn.__name__}: arguments live on different CUDA devices {sorted({str(d) for d in devs})}")
        with torch.cuda.device(devs[0]):
            return fn(*args, **kw)
    return wrapper


class DitPlan:
    """Host-side description of one batched DiT evaluation (see `scldm_dit_plan` in the header).

    n_u states are evaluated once, n_g states n_f times; `cls_idx` gives, for every conditioning row
    and class table, the embedding row (null token = vocab size, reference `nnets.py:319,403-426`);
    `slot_mod` maps every cell-forward slot to its conditioning row.
    """

    SLOT_MODES = {"table": 0, "identity": 1, "cfg_shared": 2}

    def __init__(self, packed: PackedDiT, n_u: int, n_g: int, n_f: int, coef, cls_idx: torch.Tensor, slot_mod: torch.Tensor,
                 slot_mode: str = "table"):
        lib = _lib.load()
        dev = packed.device
        s = _lib.DitPlan()
        s.n_u, s.n_g, s.n_f = int(n_u), int(n_g), int(n_f)
        coef = list(coef)
        if len(coef) > _lib.MAX_COMBINE:
            raise NotImplementedError(f"at most {_lib.MAX_COMBINE} evaluations per guided state")
        for i in range(_lib.MAX_COMBINE):
            s.coef[i] = float(coef[i]) if i < len(coef) else 0.0
        n_class = len(packed.class_names)
        n_mod = int(cls_idx.shape[1]) if n_class > 0 else int(slot_mod.max().item()) + 1
        s.n_mod = n_mod
        self.n_mod = n_mod
        mod_pad = lib.scldm_dit_mod_pad(C.byref(s))
        slots_pad = lib.scldm_dit_slots_pad(C.byref(s))
        n_slots = n_u + n_g * (n_f if n_g > 0 else 1)
        assert slot_mod.numel() == n_slots, (slot_mod.numel(), n_slots)
        ci = torch.zeros(max(n_class, 1), mod_pad, dtype=torch.int32, device=dev)
        if n_class > 0:
            assert cls_idx.shape == (n_class, n_mod)
            ci[:, :n_mod] = cls_idx.to(device=dev, dtype=torch.int32)
            # padded conditioning rows use the null token of every class (valid table rows)
            for c, name in enumerate(packed.class_names):
                ci[c, n_mod:] = packed.cfg.class_vocab_sizes[name] if packed.cfg.cfg_dropout_prob > 0 else 0
        sm = torch.zeros(slots_pad, dtype=torch.int32, device=dev)
        sm[:n_slots] = slot_mod.to(device=dev, dtype=torch.int32)
        self.cls_idx, self.slot_mod = ci.contiguous(), sm.contiguous()
        s.cls_idx, s.slot_mod = self.cls_idx.data_ptr(), self.slot_mod.data_ptr()
        s.slot_mode = self.SLOT_MODES[slot_mode]   # the table is always filled; modes 1/2 promise it has that closed form
        self.struct = s
        self.n_states = n_u + n_g
        self.n_slots, self.slots_pad, self.mod_pad = n_slots, slots_pad, mod_pad
        self.packed = packed

    def workspace_bytes(self, n_evals: int) -> int:
        return int(_lib.load().scldm_dit_workspace_bytes(C.byref(self.packed.struct), C.byref(self.struct), n_evals))


_ws_cache: dict = {}


def _workspace(device, nbytes: int, tag: str) -> torch.Tensor:
    # one scratch buffer per (device, purpose, stream): calls on different streams must not share scratch memory
    key = (str(device), tag, torch.cuda.current_stream(device).cuda_stream)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.zeros(nbytes, dtype=torch.uint8, device=device)  # zero-filled: padded slots stay finite
        _ws_cache[key] = buf
    return buf


@_on_arg_device
def dit_forward(plan: DitPlan, x: torch.Tensor, t_mod: torch.Tensor, workspace: torch.Tensor | None = None) -> torch.Tensor:
    """v = combine(DiT(x, t, cond)) for the states of `plan`.  x [n_states,16,16] fp32, t_mod [n_mod] fp32."""
    lib = _lib.load()
    _require_cuda(x, "x")
    assert x.dtype == torch.float32 and x.shape == (plan.n_states, 16, 16), x.shape
    tm = torch.zeros(plan.mod_pad, dtype=torch.float32, device=x.device)
    tm[: plan.n_mod] = t_mod.to(torch.float32)
    ws = workspace if workspace is not None else _workspace(x.device, plan.workspace_bytes(0), "dit")
    v = torch.empty_like(x)
    rc = lib.scldm_dit_forward(C.byref(plan.packed.struct), C.byref(plan.struct), x.data_ptr(), tm.data_ptr(), v.data_ptr(),
                               ws.data_ptr(), ws.numel(), _stream_ptr(x.device))
    _lib.check(rc, "scldm_dit_forward")
    return v


@_on_arg_device
def dit_forward_shared_t(plan: DitPlan, x: torch.Tensor, t: float, workspace: torch.Tensor | None = None) -> torch.Tensor:
    """`dit_forward` when every conditioning row shares the scalar time `t` (an ODE drift evaluation): one timestep-embedding
    row, no per-row time vector."""
    lib = _lib.load()
    _require_cuda(x, "x")
    assert x.dtype == torch.float32 and x.shape == (plan.n_states, 16, 16), x.shape
    ws = workspace if workspace is not None else _workspace(x.device, plan.workspace_bytes(0), "dit")
    v = torch.empty_like(x)
    rc = lib.scldm_dit_forward_shared_t(C.byref(plan.packed.struct), C.byref(plan.struct), x.data_ptr(), float(t), v.data_ptr(),
                                        ws.data_ptr(), ws.numel(), _stream_ptr(x.device))
    _lib.check(rc, "scldm_dit_forward_shared_t")
    return v


@_on_arg_device
def dit_sample_ode(plan: DitPlan, x: torch.Tensor, t_grid: torch.Tensor, method: str = "euler",
                   workspace: torch.Tensor | None = None) -> torch.Tensor:
    """Integrates x (in place) over the fixed time grid; returns x."""
    lib = _lib.load()
    _require_cuda(x, "x")
    assert x.dtype == torch.float32 and x.shape == (plan.n_states, 16, 16), x.shape
    if method not in _lib.ODE_METHODS:
        raise NotImplementedError(f"ODE method '{method}': fixed-grid {sorted(_lib.ODE_METHODS)} are implemented on device")
    grid = [float(v) for v in t_grid.detach().cpu().to(torch.float32).tolist()]
    n_grid = len(grid)
    stages = 1 if method == "euler" else 2
    arr = (C.c_float * n_grid)(*grid)
    ws = workspace if workspace is not None else _workspace(x.device, plan.workspace_bytes((n_grid - 1) * stages), "dit")
    rc = lib.scldm_dit_sample_ode(C.byref(plan.packed.struct), C.byref(plan.struct), x.data_ptr(), arr, n_grid,
                                  _lib.ODE_METHODS[method], ws.data_ptr(), ws.numel(), _stream_ptr(x.device))
    _lib.check(rc, "scldm_dit_sample_ode")
    return x


def dit_workspace_views(plan: DitPlan, ws: torch.Tensor, n_evals: int = 0) -> dict:
    """Test helper: typed views of the intermediates a call left in the workspace."""
    lib = _lib.load()
    offs = (C.c_size_t * 6)()
    n = lib.scldm_dit_workspace_layout(C.byref(plan.packed.struct), C.byref(plan.struct), n_evals, offs, 6)
    assert n == 6
    base = (-ws.data_ptr()) % 1024
    rows = plan.slots_pad * 16
    w = plan.packed
    names = ["X", "mod", "cls", "temb", "acc", "tvals"]
    o = {k: base + int(offs[i]) for i, k in enumerate(names)}

    def view(off, nbytes, dtype, shape):
        return ws[off: off + nbytes].view(dtype).view(*shape)

    # dit_stack_kernel keeps the residual stream tile-blocked: [tile][col / 4][row % 128][4] (csrc/dit_kernels.cuh: x_index)
    X = view(o["X"], rows * 256 * 4, torch.float32, (rows // 128, 64, 128, 4)).permute(0, 2, 1, 3).reshape(rows, 256)
    return {
        "X": X,
        "mod": view(o["mod"], plan.mod_pad * w.mod_stride * 4, torch.float32, (plan.mod_pad, w.mod_stride)),
        "cls": view(o["cls"], plan.mod_pad * 256 * 4, torch.float32, (plan.mod_pad, 256)),
        "temb": view(o["temb"], plan.mod_pad * 256 * 4, torch.float32, (plan.mod_pad, 256)),
    }


@_on_arg_device
def vae_qside(packed: PackedVAEDecoder):
    """Cell-invariant MCAB query projections for the whole vocabulary, fp32 and bf16 (cached on `packed`)."""
    if packed.qp is None:
        lib = _lib.load()
        qp = torch.empty(packed.emb.shape[0], 32, dtype=torch.float32, device=packed.device)
        qpb = torch.empty(packed.emb.shape[0], 32, dtype=torch.bfloat16, device=packed.device)
        rc = lib.scldm_vae_qside(C.byref(packed.struct), qp.data_ptr(), qpb.data_ptr(), _stream_ptr(packed.device))
        _lib.check(rc, "scldm_vae_qside")
        packed.qp, packed.qp_bf16 = qp, qpb
    return packed.qp, packed.qp_bf16


@_on_arg_device
def vae_decode(packed: PackedVAEDecoder, z: torch.Tensor, genes: torch.Tensor, lib_size: torch.Tensor, want_mu=True,
               want_counts=False, seed: int = 0, cell_offset: int = 0, out_mu: torch.Tensor | None = None,
               out_counts: torch.Tensor | None = None, precision: str = "bf16"):
    """z [cells,16,16] fp32, genes [G] int64 (shared by all cells), lib_size [cells] fp32 ->
    (mu [cells,G] | None, theta [G] (shared table) or [cells,G] (unshared-theta head), counts [cells,G] | None)."""
    lib = _lib.load()
    _require_cuda(z, "z")
    n_cells, G = z.shape[0], genes.numel()
    assert z.dtype == torch.float32 and z.shape[1:] == (16, 16)
    assert genes.dtype == torch.int64 and genes.is_cuda and genes.dim() == 1
    assert int(lib_size.numel()) == n_cells
    lib_size = lib_size.reshape(-1).to(torch.float32).contiguous()
    qp, qpb = vae_qside(packed)
    for o in (out_mu, out_counts):
        if o is not None:
            assert o.is_cuda and o.is_contiguous() and o.dtype == torch.float32 and o.shape == (n_cells, G)
    mu = out_mu if out_mu is not None else (torch.empty(n_cells, G, dtype=torch.float32, device=z.device) if want_mu else None)
    counts = out_counts if out_counts is not None else (
        torch.empty(n_cells, G, dtype=torch.float32, device=z.device) if want_counts else None)
    theta = torch.empty(G if getattr(packed, "shared_theta", True) else (n_cells, G), dtype=torch.float32, device=z.device)
    nbytes = int(lib.scldm_vae_decode_workspace_bytes(n_cells, G))
    ws = _workspace(z.device, nbytes, "vae")
    rc = lib.scldm_vae_decode(C.byref(packed.struct), qp.data_ptr(), qpb.data_ptr(), z.data_ptr(), n_cells, genes.data_ptr(), G,
                              lib_size.data_ptr(), mu.data_ptr() if mu is not None else None, theta.data_ptr(),
                              counts.data_ptr() if counts is not None else None, seed & (2**64 - 1), cell_offset,
                              _lib.DECODE_PRECISION[precision], ws.data_ptr(), ws.numel(), _stream_ptr(z.device))
    _lib.check(rc, "scldm_vae_decode")
    return mu, theta, counts


@_on_arg_device
def vae_encode(packed: PackedVAEEncoder, genes_subset: torch.Tensor, counts_subset: torch.Tensor) -> torch.Tensor:
    """genes_subset [cells,S] int64, counts_subset [cells,S] float -> z [cells,16,16] fp32."""
    lib = _lib.load()
    _require_cuda(genes_subset, "genes_subset")
    _require_cuda(counts_subset, "counts_subset")
    assert genes_subset.dtype == torch.int64 and genes_subset.dim() == 2 and genes_subset.shape == counts_subset.shape
    counts = counts_subset.to(torch.float32).contiguous()
    n_cells, S = genes_subset.shape
    z = torch.empty(n_cells, 16, 16, dtype=torch.float32, device=genes_subset.device)
    rc = lib.scldm_vae_encode(C.byref(packed.struct), genes_subset.data_ptr(), counts.data_ptr(), n_cells, S, z.data_ptr(),
                              _stream_ptr(genes_subset.device))
    _lib.check(rc, "scldm_vae_encode")
    return z


@_on_arg_device
def vae256_decode(packed, z: torch.Tensor, genes: torch.Tensor, lib_size: torch.Tensor, want_mu=True, want_counts=False, seed: int = 0,
                  cell_offset: int = 0, out_mu: torch.Tensor | None = None, out_counts: torch.Tensor | None = None, max_rows: int = 1 << 21):
    """`vae_decode` for the n_embed = 256 VAE (`pack256.PackedVAE256Decoder`).  The cells are processed in chunks of at most
    `max_rows` gene rows (cells x padded G) so that the (rows, 256) intermediates of the tensor-core MCAB stay a few GB."""
    lib = _lib.load()
    _require_cuda(z, "z")
    n_cells, G = z.shape[0], genes.numel()
    assert z.dtype == torch.float32 and z.shape[1:] == (16, 16) and genes.dtype == torch.int64 and genes.is_cuda and genes.dim() == 1
    assert int(lib_size.numel()) == n_cells
    lib_size = lib_size.reshape(-1).to(torch.float32).contiguous()
    dev = z.device
    if packed.qp_bf16 is None:      # cell-invariant query side, once per vocabulary
        qp = torch.empty(packed.emb.shape[0], 256, dtype=torch.bfloat16, device=dev)
        ws = _workspace(dev, int(lib.scldm_vae256_qside_workspace_bytes(packed.emb.shape[0])), "vae256q")
        _lib.check(lib.scldm_vae256_qside(C.byref(packed.struct), qp.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(dev)), "scldm_vae256_qside")
        packed.qp_bf16 = qp
    mu = out_mu if out_mu is not None else (torch.empty(n_cells, G, dtype=torch.float32, device=dev) if want_mu else None)
    counts = out_counts if out_counts is not None else (torch.empty(n_cells, G, dtype=torch.float32, device=dev) if want_counts else None)
    theta = torch.empty(G, dtype=torch.float32, device=dev)
    g_pad = -(-G // 128) * 128
    per = max(8, (max_rows // g_pad) // 8 * 8)
    ws = _workspace(dev, int(lib.scldm_vae256_decode_workspace_bytes(min(per, n_cells), G)), "vae256d")
    for c0 in range(0, n_cells, per):
        c1 = min(c0 + per, n_cells)
        rc = lib.scldm_vae256_decode(C.byref(packed.struct), packed.qp_bf16.data_ptr(), z[c0:c1].data_ptr(), c1 - c0, genes.data_ptr(), G,
                                     lib_size[c0:c1].data_ptr(), mu[c0:c1].data_ptr() if mu is not None else None, theta.data_ptr(),
                                     counts[c0:c1].data_ptr() if counts is not None else None, seed & (2**64 - 1), cell_offset + c0,
                                     ws.data_ptr(), ws.numel(), _stream_ptr(dev))
        _lib.check(rc, "scldm_vae256_decode")
    return mu, theta, counts


@_on_arg_device
def vae256_encode(packed, genes_subset: torch.Tensor, counts_subset: torch.Tensor, max_rows: int = 1 << 20) -> torch.Tensor:
    """`vae_encode` for the n_embed = 256 VAE (`pack256.PackedVAE256Encoder`), in chunks of at most `max_rows` token rows."""
    lib = _lib.load()
    _require_cuda(genes_subset, "genes_subset")
    _require_cuda(counts_subset, "counts_subset")
    assert genes_subset.dtype == torch.int64 and genes_subset.dim() == 2 and genes_subset.shape == counts_subset.shape
    counts = counts_subset.to(torch.float32).contiguous()
    genes_subset = genes_subset.contiguous()
    n_cells, S = genes_subset.shape
    dev = genes_subset.device
    z = torch.empty(n_cells, 16, 16, dtype=torch.float32, device=dev)
    s_pad = -(-S // 128) * 128
    per = max(8, (max_rows // s_pad) // 8 * 8)
    ws = _workspace(dev, int(lib.scldm_vae256_encode_workspace_bytes(min(per, n_cells), S)), "vae256e")
    for c0 in range(0, n_cells, per):
        c1 = min(c0 + per, n_cells)
        rc = lib.scldm_vae256_encode(C.byref(packed.struct), genes_subset[c0:c1].data_ptr(), counts[c0:c1].data_ptr(), c1 - c0, S, z[c0:c1].data_ptr(),
                                     ws.data_ptr(), ws.numel(), _stream_ptr(dev))
        _lib.check(rc, "scldm_vae256_encode")
    return z


@_on_arg_device
def randn_cells(n_cells: int, per_cell: int, seed: int, cell_offset: int, stream_id: int, device) -> torch.Tensor:
    lib = _lib.load()
    out = torch.empty(n_cells, per_cell, dtype=torch.float32, device=device)
    rc = lib.scldm_randn_cells(out.data_ptr(), n_cells, per_cell, seed & (2**64 - 1), cell_offset, stream_id, _stream_ptr(device))
    _lib.check(rc, "scldm_randn_cells")
    return out


@_on_arg_device
def counts_to_csr(counts: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Dense (rows, G) fp32 counts -> the arrays of `scipy.sparse.csr_matrix(counts)` built on the device:
    `indptr` int64 [rows+1], `indices` int32 [nnz] (ascending within a row), `data` fp32 [nnz].  The reference builds
    them on the host from the dense copy (`_utils.py:186-200`); one 8-byte D2H read (nnz) sizes the outputs here."""
    if counts.device.type != "cuda" or counts.dtype != torch.float32 or counts.dim() != 2:
        raise RuntimeError("counts_to_csr expects a 2-D float32 CUDA tensor (no CPU fallback)")
    lib = _lib.load()
    counts = counts.contiguous()
    rows, G = counts.shape
    dev = counts.device
    row_nnz = torch.empty(max(rows, 1), dtype=torch.int32, device=dev)
    indptr = torch.empty(rows + 1, dtype=torch.int64, device=dev)
    _lib.check(lib.scldm_csr_count(counts.data_ptr(), rows, G, row_nnz.data_ptr(), indptr.data_ptr(), _stream_ptr(dev)), "scldm_csr_count")
    nnz = int(indptr[-1].item())
    indices = torch.empty(nnz, dtype=torch.int32, device=dev)
    data = torch.empty(nnz, dtype=torch.float32, device=dev)
    _lib.check(lib.scldm_csr_fill(counts.data_ptr(), rows, G, indptr.data_ptr(), indices.data_ptr(), data.data_ptr(), _stream_ptr(dev)),
               "scldm_csr_fill")
    return indptr, indices, data


@_on_arg_device
def nb_nll(counts: torch.Tensor, mu: torch.Tensor, theta: torch.Tensor) -> torch.Tensor:
    """Per-cell NB reconstruction loss `(-log_nb_positive(counts, mu, theta)).sum(dim=1)` (`distributions.py:6-42`,
    `models.py:233-247`): counts / mu (N, G) fp32, theta (N, G) or a shared (G,) row -> (N,) fp32."""
    for name, t in (("counts", counts), ("mu", mu), ("theta", theta)):
        if t.device.type != "cuda" or t.dtype != torch.float32:
            raise RuntimeError(f"nb_nll: {name} must be a float32 CUDA tensor (no CPU fallback)")
    if counts.dim() != 2 or mu.shape != counts.shape or theta.shape[-1] != counts.shape[1] or theta.dim() not in (1, 2):
        raise ValueError(f"nb_nll: shape mismatch counts {tuple(counts.shape)} mu {tuple(mu.shape)} theta {tuple(theta.shape)}")
    if theta.dim() == 2 and theta.shape[0] not in (1, counts.shape[0]):
        raise ValueError("nb_nll: theta must have one row or one row per cell")
    if theta.dim() == 2 and (theta.shape[0] == 1 or theta.stride(0) == 0):
        theta = theta[0]   # one shared row (the expanded view `decode` returns for shared_theta)
    counts, mu, theta = counts.contiguous(), mu.contiguous(), theta.contiguous()
    rows, G = counts.shape
    stride = G if theta.dim() == 2 else 0
    out = torch.empty(rows, dtype=torch.float32, device=counts.device)
    _lib.check(_lib.load().scldm_nb_nll(counts.data_ptr(), mu.data_ptr(), theta.data_ptr(), stride, rows, G, out.data_ptr(),
                                        _stream_ptr(counts.device)), "scldm_nb_nll")
    return out


@_on_arg_device
def tokenize_expressed(counts: torch.Tensor, gene_ids: torch.Tensor, genes_seq_len: int, mask_idx: int = 0) -> dict[str, torch.Tensor]:
    """`tokenize_cells(..., sample_genes="expressed")` (`datamodule.py:708-731`) on the device: dense (N, G) counts and the
    (G,) gene-token row -> `genes_subset` int64 / `counts_subset` fp32 (N, genes_seq_len) with the expressed genes packed
    left, plus `library_size` (N, 1).  Raises ValueError like the reference when a cell expresses more genes than fit."""
    if counts.device.type != "cuda" or counts.dtype != torch.float32 or counts.dim() != 2:
        raise RuntimeError("tokenize_expressed expects a 2-D float32 CUDA tensor (no CPU fallback)")
    lib = _lib.load()
    counts = counts.contiguous()
    rows, G = counts.shape
    dev = counts.device
    gene_ids = gene_ids.to(device=dev, dtype=torch.int64).contiguous()
    if gene_ids.numel() != G:
        raise ValueError("gene_ids must hold one token per column of counts")
    genes_out = torch.empty(rows, genes_seq_len, dtype=torch.int64, device=dev)
    counts_out = torch.empty(rows, genes_seq_len, dtype=torch.float32, device=dev)
    library = torch.empty(rows, 1, dtype=torch.float32, device=dev)
    overflow = torch.zeros(1, dtype=torch.int32, device=dev)
    rc = lib.scldm_tokenize_expressed(counts.data_ptr(), rows, G, gene_ids.data_ptr(), genes_seq_len, mask_idx, genes_out.data_ptr(),
                                      counts_out.data_ptr(), library.data_ptr(), overflow.data_ptr(), _stream_ptr(dev))
    _lib.check(rc, "scldm_tokenize_expressed")
    if int(overflow.item()) > 0:
        raise ValueError("genes_seq_len is smaller than number of expressed genes")
    return {"genes_subset": genes_out, "counts_subset": counts_out, "library_size": library}


@_on_arg_device
def sde_drift(v: torch.Tensor, x: torch.Tensor, t: float, form: int, norm: float) -> torch.Tensor:
    """v + D(t) (t v - x) / (1 - t): SDE drift of the Linear path with a velocity model (`transport.py:231-233`)."""
    _require_cuda(v, "v")
    _require_cuda(x, "x")
    out = torch.empty_like(v)
    _lib.check(_lib.load().scldm_sde_drift(v.data_ptr(), x.data_ptr(), float(t), int(form), float(norm), out.data_ptr(), v.numel(), _stream_ptr(v.device)), "scldm_sde_drift")
    return out


@_on_arg_device
def sde_kick(x: torch.Tensor, noise: torch.Tensor | None, t: float, dt: float, form: int, norm: float, seed: int, cell_offset: int, per_cell: int, step: int,
             x_for_diffusion: torch.Tensor | None = None) -> torch.Tensor:
    """x + sqrt(2 D(t)) sqrt(dt) w with w = `noise` or Philox normals keyed by (seed, global cell, element, step)."""
    _require_cuda(x, "x")
    out = torch.empty_like(x)
    if noise is not None:
        noise = noise.contiguous().float()
    rc = _lib.load().scldm_sde_kick(x.data_ptr(), noise.data_ptr() if noise is not None else None, float(t), float(dt), int(form), float(norm), seed & (2**64 - 1),
                                    int(cell_offset), int(per_cell), int(step), out.data_ptr(), x.numel(), _stream_ptr(x.device))
    _lib.check(rc, "scldm_sde_kick")
    return out


@_on_arg_device
def axpy2(a: torch.Tensor, c1: float, d1: torch.Tensor, c2: float = 0.0, d2: torch.Tensor | None = None) -> torch.Tensor:
    """a + c1 d1 (+ c2 d2)."""
    _require_cuda(a, "a")
    out = torch.empty_like(a)
    rc = _lib.load().scldm_axpy2(a.data_ptr(), float(c1), d1.contiguous().data_ptr(), float(c2), d2.contiguous().data_ptr() if d2 is not None else None, out.data_ptr(),
                                 a.numel(), _stream_ptr(a.device))
    _lib.check(rc, "scldm_axpy2")
    return out


def prof_enable(on: bool, device=None) -> None:
    """Bracket every library launch with CUDA events on the current stream (bench.py roofline timing)."""
    dev = device if device is not None else torch.cuda.current_device()
    _lib.load().scldm_prof_enable(1 if on else 0, _stream_ptr(dev))


def prof_summary() -> dict[str, tuple[int, float]]:
    """{kernel class: (launches, total ms)} since the last call (synchronises the device)."""
    buf = C.create_string_buffer(1 << 16)
    n = _lib.load().scldm_prof_summary(buf, len(buf))
    out = {}
    for line in buf.raw[:n].decode().splitlines():
        name, cnt, ms = line.rsplit(" ", 2)
        out[name] = (int(cnt), float(ms))
    return out


def set_option(name: str, value: int) -> None:
    """Runtime switch of the library (`scldm_set_option`): "solve", "pdl", "mod_batch", "dec_cpb", "dec_occ"."""
    _lib.check(_lib.load().scldm_set_option(name.encode(), int(value)), "scldm_set_option")


def get_option(name: str) -> int:
    return int(_lib.load().scldm_get_option(name.encode()))


def launch_count() -> int:
    return int(_lib.load().scldm_launch_count())
