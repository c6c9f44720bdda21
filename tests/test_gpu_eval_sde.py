"""GPU tests of the rows either side of the path that SURVEY.md 8f ranks 4th: the SDE sampler (`Sampler.sample_sde`) against
trajectories minted from the unmodified reference with the reference's own Brownian increments, and the generation-evaluation
metrics (`scldm_b200.evaluations`) against the reference formulas restated in plain torch (the reference module itself imports POT,
which is not installable here)."""

import os

import numpy as np
import pytest
import torch

from oracle.make_golden import WEIGHT_SEED, golden_cases
from scldm_b200 import synthetic

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize("method", ["Euler", "Heun"])
def test_sample_sde_vs_reference_golden(golden_dir, method):
    from scldm_b200.nnets import DiT
    from scldm_b200.transport import Sampler, create_transport
    from scldm_b200.transport.transport import FusedCFGModel

    g = dict(np.load(os.path.join(golden_dir, "sde_me1.npz")))
    cfg = golden_cases()["dit_me1"]["cfg"]
    dit = DiT(**cfg.kwargs())
    dit.load_state_dict(synthetic.dit_state_dict(cfg, WEIGHT_SEED))
    dit = dit.cuda().eval()
    z0, lab = torch.from_numpy(g["z0"]).cuda(), torch.from_numpy(g["label"]).cuda()
    fn = Sampler(create_transport("Linear", "velocity")).sample_sde(sampling_method=method, diffusion_form="sigma", num_steps=10, last_step="Mean", last_step_size=0.04)
    noise = [torch.from_numpy(n) for n in g[f"{method}.noise"]]
    xs = fn(torch.cat([z0, z0]), FusedCFGModel(dit, {"clusters": 2.0}), noise=noise, condition={"clusters": torch.cat([lab, lab])})
    ref = g[f"{method}.states"]
    assert len(xs) == 10 and tuple(xs[0].shape) == ref.shape[1:]
    errs = [rel_l2(x, r) for x, r in zip(xs, ref)]
    print(method, "per-state rel-L2", [f"{e:.1e}" for e in errs])
    assert max(errs) < 5e-3, errs       # bf16 tensor-core DiT, fp32 state / noise / update
    # own Philox increments: deterministic under a seed, different across seeds, finite
    a = fn(torch.cat([z0, z0]), FusedCFGModel(dit, {"clusters": 2.0}), condition={"clusters": torch.cat([lab, lab])})
    b = fn(torch.cat([z0, z0]), FusedCFGModel(dit, {"clusters": 2.0}), condition={"clusters": torch.cat([lab, lab])})
    assert torch.equal(a[-1], b[-1]) and bool(torch.isfinite(a[-1]).all()) and not torch.equal(a[-1], xs[-1])


def _ref_kernels(x, y):
    xe, ye = x.unsqueeze(1), y.unsqueeze(0)
    sq = (x**2).sum(1, keepdim=True) - 2 * x @ y.T + (y**2).sum(1, keepdim=True).T
    return {
        "rbf": torch.exp(-1.0 * sq),
        "braycurtis": 1 - (xe - ye).abs().sum(2) / ((xe + ye).abs().sum(2) + 1e-8),
        "tanimoto": (xe * ye).sum(2) / ((xe + ye - xe * ye).sum(2) + 1e-8),
        "ruzicka": torch.min(xe, ye).sum(2) / (torch.max(xe, ye).sum(2) + 1e-8),
    }


def test_mmd_kernels_match_reference_formulas():
    from scldm_b200 import evaluations as E

    gen = torch.Generator().manual_seed(3)
    x = torch.poisson(torch.rand(70, 333, generator=gen) * 3, generator=gen).double()
    y = torch.poisson(torch.rand(45, 333, generator=gen) * 2, generator=gen).double()
    ref = _ref_kernels(x / 10, y / 10)      # scaled so that the RBF kernel is not all zeros
    xs, ys = (x / 10).float().cuda(), (y / 10).float().cuda()
    mine = {"rbf": E.RBFKernel()(xs, ys), "braycurtis": E.BrayCurtisKernel()(xs, ys), "tanimoto": E.TanimotoKernel()(xs, ys), "ruzicka": E.RuzickaKernel()(xs, ys)}
    for k in ref:
        assert mine[k].shape == (70, 45) and rel_l2(mine[k], ref[k]) < 1e-5, k
    for name, kern in (("braycurtis", E.BrayCurtisKernel()), ("ruzicka", E.RuzickaKernel()), ("rbf", E.RBFKernel())):
        kk = {n: _ref_kernels(a, b)[name] for n, (a, b) in {"xx": (x / 10, x / 10), "yy": (y / 10, y / 10), "xy": (x / 10, y / 10)}.items()}
        want = float(kk["xx"].mean() + kk["yy"].mean() - 2 * kk["xy"].mean())
        got = float(E.MMDLoss(kern)(xs, ys))
        assert abs(got - want) < 1e-5 * max(1.0, abs(want)), (name, got, want)


def test_sinkhorn_wasserstein_matches_a_float64_restatement():
    from scldm_b200 import evaluations as E

    gen = torch.Generator().manual_seed(4)
    x0 = torch.randn(60, 8, generator=gen) * 0.3
    x1 = torch.randn(50, 8, generator=gen) * 0.3 + 0.2
    for power in (1, 2):
        M = torch.cdist(x0.double(), x1.double())
        if power == 2:
            M = M**2
        K = torch.exp(-M / 0.05)
        a, b = torch.full((60,), 1 / 60, dtype=torch.float64), torch.full((50,), 1 / 50, dtype=torch.float64)
        u, v = torch.ones(60, dtype=torch.float64) / 60, torch.ones(50, dtype=torch.float64) / 50
        for _ in range(2000):      # POT sinkhorn_knopp iterations
            v = b / (K.T @ u)
            u = a / (K @ v)
        want = float((u[:, None] * K * v[None, :] * M).sum())
        want = want**0.5 if power == 2 else want
        got = E.wasserstein(x0.cuda(), x1.cuda(), method="sinkhorn", power=power, numItermax=2000)
        assert abs(got - want) < 2e-3 * want, (power, got, want)
