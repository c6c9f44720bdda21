"""dopri5 restatement (scldm_b200/transport/adaptive.py) on ODEs with known solutions (CPU)."""

import math

import pytest

import torch

from scldm_b200.transport.adaptive import dopri5


def test_linear_decay_and_oscillator():
    y0 = torch.tensor([[1.0, 2.0], [0.5, -1.0]], dtype=torch.float64)
    ts = torch.linspace(0, 1, 11).tolist()
    traj, nfe = dopri5(lambda t, y: -3.0 * y, y0, ts, rtol=1e-7, atol=1e-9)
    for i, t in enumerate(ts):
        assert torch.allclose(traj[i], y0 * math.exp(-3 * t), rtol=1e-5, atol=1e-8)
    assert traj.shape == (11, 2, 2) and 20 < nfe < 400

    A = torch.tensor([[0.0, 1.0], [-1.0, 0.0]], dtype=torch.float64)
    traj, _ = dopri5(lambda t, y: y @ A.T, torch.tensor([[1.0, 0.0]], dtype=torch.float64), [0.0, 1.0, 2.5], rtol=1e-8, atol=1e-10)
    assert abs(float(traj[-1, 0, 0]) - math.cos(2.5)) < 1e-6 and abs(float(traj[-1, 0, 1]) + math.sin(2.5)) < 1e-6


def test_time_dependent_rhs_and_tolerance_scaling():
    f = lambda t, y: torch.full_like(y, 1.0) * math.cos(5 * t)  # noqa: E731
    y0 = torch.zeros(3, dtype=torch.float64)
    loose, n_loose = dopri5(f, y0, [0.0, 1.0], rtol=1e-3, atol=1e-3)
    tight, n_tight = dopri5(f, y0, [0.0, 1.0], rtol=1e-8, atol=1e-8)
    exact = math.sin(5.0) / 5
    assert abs(float(tight[-1, 0]) - exact) < 1e-7 and abs(float(loose[-1, 0]) - exact) < 2e-2
    assert n_tight > n_loose


@pytest.mark.parametrize("method", ["euler", "heun2", "midpoint"])
def test_fixed_grid_host_loop_matches_oracle_stepper(method):
    """`Sampler.sample_ode` with an opaque callable integrates on the host with the fixed-grid formulas; on CPU tensors (no
    kernels involved) it must reproduce the oracle's restatement of torchdiffeq's steppers, which the golden trajectories pin."""
    import torch

    from oracle import scldm_oracle as O
    from scldm_b200.transport import Sampler, create_transport

    torch.manual_seed(0)
    x0 = torch.randn(5, 16, 16)
    A = 0.3 * torch.randn(16, 16)

    def f(x, t):                       # smooth, time-dependent, couples channels
        return -x @ A + torch.sin(3.0 * t).view(-1, 1, 1) * torch.ones_like(x)

    fn = Sampler(create_transport("Linear", "velocity")).sample_ode(sampling_method=method, num_steps=12)
    ours = fn(x0, lambda x, t, **kw: f(x, t))[-1]
    ref = O.sample_ode(x0, f, num_steps=12, method=method)[-1]
    assert torch.allclose(ours, ref, rtol=1e-6, atol=1e-6)
