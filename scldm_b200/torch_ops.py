"""`torch.library` registration of the hot-path entry points, so that the ops are first-class PyTorch operators
(`torch.ops.scldm_b200.*`): visible to `torch.compile` / `torch.export` tracing (with shape-propagating fake kernels) and to
profilers, instead of opaque ctypes calls.  The device kernels are the same C-ABI calls `scldm_b200.ops` makes.

Packed weights and evaluation plans are host objects, not tensors: an op receives an integer handle from `register(obj)`.

    h = torch_ops.register(vae.packed_encoder())
    z = torch.ops.scldm_b200.vae_encode(genes_subset, counts_subset, h)
"""

from __future__ import annotations

import itertools

import torch

from . import ops

_registry: dict[int, object] = {}
_next = itertools.count(1)


def register(obj) -> int:
    """Keep `obj` (PackedDiT plan / packed VAE weights) alive and return the handle the operators take."""
    h = next(_next)
    _registry[h] = obj
    return h


def release(handle: int) -> None:
    _registry.pop(handle, None)


def _get(handle: int):
    try:
        return _registry[handle]
    except KeyError:
        raise RuntimeError(f"scldm_b200: unknown handle {handle} (torch_ops.register it first)") from None


@torch.library.custom_op("scldm_b200::dit_forward", mutates_args=())
def dit_forward(x: torch.Tensor, t_mod: torch.Tensor, plan: int) -> torch.Tensor:
    """DiT.forward / forward_with_cfg for the states of a registered `ops.DitPlan` (nnets.py:273-378)."""
    return ops.dit_forward(_get(plan), x, t_mod)


@dit_forward.register_fake
def _(x, t_mod, plan):
    return torch.empty_like(x)


@torch.library.custom_op("scldm_b200::dit_sample_ode", mutates_args=())
def dit_sample_ode(x: torch.Tensor, t_grid: torch.Tensor, method: str, plan: int) -> torch.Tensor:
    """Fixed-grid ODE solve (transport.py:324-369 + integrators.py:100-112): returns x(t_grid[-1]); t_grid is a CPU tensor."""
    return ops.dit_sample_ode(_get(plan), x.clone(), t_grid, method)


@dit_sample_ode.register_fake
def _(x, t_grid, method, plan):
    return torch.empty_like(x)


@torch.library.custom_op("scldm_b200::vae_encode", mutates_args=())
def vae_encode(genes_subset: torch.Tensor, counts_subset: torch.Tensor, packed: int) -> torch.Tensor:
    """TransformerVAE.encode (vae.py:58-69) -> z (cells, 16, 16)."""
    p = _get(packed)
    fn = ops.vae256_encode if type(p).__name__ == "PackedVAE256Encoder" else ops.vae_encode
    return fn(p, genes_subset, counts_subset)


@vae_encode.register_fake
def _(genes_subset, counts_subset, packed):
    return counts_subset.new_empty((genes_subset.shape[0], 16, 16), dtype=torch.float32)


@torch.library.custom_op("scldm_b200::vae_decode", mutates_args=())
def vae_decode(z: torch.Tensor, genes: torch.Tensor, library_size: torch.Tensor, seed: int, cell_offset: int, packed: int) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """TransformerVAE.decode + NegativeBinomial.sample (vae.py:71-87, models.py:819) -> (mu (cells, G), theta (G,), counts (cells, G))."""
    p = _get(packed)
    fn = ops.vae256_decode if type(p).__name__ == "PackedVAE256Decoder" else ops.vae_decode
    mu, theta, counts = fn(p, z, genes, library_size, want_mu=True, want_counts=True, seed=seed, cell_offset=cell_offset)
    return mu, theta, counts


@vae_decode.register_fake
def _(z, genes, library_size, seed, cell_offset, packed):
    n, g = z.shape[0], genes.numel()
    return z.new_empty((n, g)), z.new_empty((g,)), z.new_empty((n, g))
