"""Drop-in for `scldm.evaluations` (`src/scldm/evaluations.py:10-108`): the MMD kernels used by the generation-evaluation hooks
(`models.py:40-48`) and the Sinkhorn Wasserstein distance, computed on the device.

One kernel pass (`scldm_pair_stats`) yields, for every pair of rows, sum x*y, sum |x - y|, sum |x + y| and sum min(x, y); the four
MMD kernel matrices are O(1) functions of those - the reference materialises a (Bx, By, D) broadcast per kernel instead.
`wasserstein(method="sinkhorn")` restates POT's `sinkhorn2` / `sinkhorn_knopp` (third-party, not installable here: flagged as a
restatement; stopping rule err < 1e-9 checked every 10 iterations, as POT).  `method="emd"` (exact LP, `ot.emd2`) is a host solver in
the reference and is not provided."""

from __future__ import annotations

import ctypes as C
import math

import torch
from torch import nn

from . import _lib


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def pair_stats(x: torch.Tensor, y: torch.Tensor) -> dict[str, torch.Tensor]:
    """x (Bx, D), y (By, D) float32 CUDA -> {"dot", "l1", "abs_sum", "min_sum"}: (Bx, By) each."""
    if not (x.is_cuda and y.is_cuda) or x.dim() != 2 or y.dim() != 2 or x.shape[1] != y.shape[1]:
        raise RuntimeError("pair_stats expects two 2-D CUDA tensors with the same number of columns (no CPU fallback)")
    x, y = x.float().contiguous(), y.float().contiguous()
    out = torch.empty(4, x.shape[0], y.shape[0], dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.load().scldm_pair_stats(x.data_ptr(), x.shape[0], y.data_ptr(), y.shape[0], x.shape[1], out.data_ptr(), _stream(x.device))
    _lib.check(rc, "scldm_pair_stats")
    return {"dot": out[0], "l1": out[1], "abs_sum": out[2], "min_sum": out[3]}


class RBFKernel(nn.Module):
    """exp(-scale ||x - y||^2) (`evaluations.py:10-22`)."""

    def __init__(self, scale: float = 1.0):
        super().__init__()
        self.scale = scale

    def forward(self, x, y):
        s = pair_stats(x, y)
        sq = (x.float() ** 2).sum(1, keepdim=True) - 2 * s["dot"] + (y.float() ** 2).sum(1, keepdim=True).T
        return torch.exp(-self.scale * sq)


class BrayCurtisKernel(nn.Module):
    """1 - sum|x - y| / (sum|x + y| + 1e-8) (`evaluations.py:25-39`)."""

    def forward(self, x, y):
        s = pair_stats(x, y)
        return 1 - s["l1"] / (s["abs_sum"] + 1e-8)


class TanimotoKernel(nn.Module):
    """sum xy / (sum(x + y - xy) + 1e-8) (`evaluations.py:42-55`)."""

    def forward(self, x, y):
        s = pair_stats(x, y)
        den = x.float().sum(1, keepdim=True) + y.float().sum(1, keepdim=True).T - s["dot"] + 1e-8
        return s["dot"] / den


class RuzickaKernel(nn.Module):
    """sum min(x, y) / (sum max(x, y) + 1e-8) (`evaluations.py:58-69`); max = x + y - min."""

    def forward(self, x, y):
        s = pair_stats(x, y)
        mx = x.float().sum(1, keepdim=True) + y.float().sum(1, keepdim=True).T - s["min_sum"]
        return s["min_sum"] / (mx + 1e-8)


class MMDLoss(nn.Module):
    """`MMDLoss` (`evaluations.py:72-82`): mean k(x,x) + mean k(y,y) - 2 mean k(x,y)."""

    def __init__(self, kernel):
        super().__init__()
        self.kernel = kernel

    def forward(self, x, y):
        return self.kernel(x, x).mean() + self.kernel(y, y).mean() - 2 * self.kernel(x, y).mean()


def wasserstein(x0: torch.Tensor, x1: torch.Tensor, method: str = "sinkhorn", reg: float = 0.05, power: int = 2, numItermax: int = 1000,
                stopThr: float = 1e-9) -> float:
    """`wasserstein` (`evaluations.py:85-108`) with the entropic solver: uniform marginals, cost = cdist(x0, x1)^power,
    Sinkhorn-Knopp scaling on the device; returns sqrt(cost) for power 2."""
    assert power == 1 or power == 2
    if method != "sinkhorn":
        raise NotImplementedError("only method='sinkhorn' runs on the device (ot.emd2 is an exact host LP solver)")
    s = pair_stats(x0, x1)
    sq = ((x0.float() ** 2).sum(1, keepdim=True) - 2 * s["dot"] + (x1.float() ** 2).sum(1, keepdim=True).T).clamp_min(0)
    M = (sq if power == 2 else sq.sqrt()).contiguous()
    n, m = M.shape
    dev = M.device
    a = torch.full((n,), 1.0 / n, device=dev)
    b = torch.full((m,), 1.0 / m, device=dev)
    K = torch.exp(-M / reg).contiguous()
    u, v = torch.full((n,), 1.0 / n, device=dev), torch.full((m,), 1.0 / m, device=dev)
    res = torch.zeros(2, device=dev)
    lib = _lib.load()
    it = 0
    while True:
        step = min(10, numItermax - it)
        with torch.cuda.device(dev):
            rc = lib.scldm_sinkhorn(K.data_ptr(), M.data_ptr(), a.data_ptr(), b.data_ptr(), n, m, u.data_ptr(), v.data_ptr(), step, res.data_ptr(), _stream(dev))
        _lib.check(rc, "scldm_sinkhorn")
        it += step
        cost, err = (float(t) for t in res.tolist())
        if not math.isfinite(err) or err < stopThr or it >= numItermax:
            break
    return math.sqrt(cost) if power == 2 else cost
