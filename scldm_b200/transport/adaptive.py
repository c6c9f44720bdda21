"""Adaptive Dormand-Prince 5(4) integrator with torchdiffeq's step controller.

RESTATEMENT of a third-party algorithm: the reference's default `sample_ode()` solver is torchdiffeq's `dopri5`
(`transport.py:327`, `integrators.py:111`), and torchdiffeq is neither vendored nor installable here.  What follows is
the published DP5(4) tableau plus torchdiffeq's conventions as recalled: mixed tolerance `atol + rtol*max(|y0|,|y1|)`,
RMS error norm over the *whole batch tensor* (one step size for all cells), `safety=0.9, ifactor=10, dfactor=0.2`,
Hairer's initial-step heuristic, FSAL, and the quartic interpolant used to read the solution at requested times.
Only function evaluations run on the GPU (the CUDA DiT); the controller is host logic, as in torchdiffeq.
"""

from __future__ import annotations

import math

import torch

# Dormand-Prince tableau
_ALPHA = [1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0]
_BETA = [
    [1 / 5],
    [3 / 40, 9 / 40],
    [44 / 45, -56 / 15, 32 / 9],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
    [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84],
]
_C_SOL = [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0]
_C_ERR = [35 / 384 - 1951 / 21600, 0, 500 / 1113 - 22642 / 50085, 125 / 192 - 451 / 720, -2187 / 6784 - -12231 / 42400,
          11 / 84 - 649 / 6300, -1.0 / 60.0]
_C_MID = [6025192743 / 30085553152 / 2, 0, 51252292925 / 65400821598 / 2, -2691868925 / 45128329728 / 2,
          187940372067 / 1594534317056 / 2, -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2]


def _rms(x: torch.Tensor) -> float:
    return float(x.pow(2).mean().sqrt())


def _initial_step(f, t0: float, y0, f0, rtol: float, atol: float, order: int = 4) -> float:
    scale = atol + y0.abs() * rtol
    d0, d1 = _rms(y0 / scale), _rms(f0 / scale)
    h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
    y1 = y0 + h0 * f0
    f1 = f(t0 + h0, y1)
    d2 = _rms((f1 - f0) / scale) / h0
    h1 = max(1e-6, h0 * 1e-3) if (d1 <= 1e-15 and d2 <= 1e-15) else (0.01 / max(d1, d2)) ** (1.0 / (order + 1))
    return min(100 * h0, h1)


def _interp(y0, y1, y_mid, f0, f1, dt, x):
    a = 2 * dt * (f1 - f0) - 8 * (y1 + y0) + 16 * y_mid
    b = dt * (5 * f0 - 3 * f1) + 18 * y0 + 14 * y1 - 32 * y_mid
    c = dt * (f1 - 4 * f0) - 11 * y0 - 5 * y1 + 16 * y_mid
    d = dt * f0
    return (((a * x + b) * x + c) * x + d) * x + y0


def dopri5(f, y0: torch.Tensor, t_eval, rtol: float = 1e-5, atol: float = 1e-5, max_steps: int = 100000):
    """Integrates dy/dt = f(t, y) from t_eval[0] and returns (states at every t_eval, number of function evaluations)."""
    ts = [float(v) for v in t_eval]
    t = ts[0]
    y = y0
    f0 = f(t, y)
    nfe = 1
    dt = _initial_step(f, t, y, f0, rtol, atol)
    nfe += 1
    out = [y0]
    nxt = 1
    prev = None  # (t0, t1, y0, y1, y_mid, f0, f1) of the last accepted step
    steps = 0
    while nxt < len(ts):
        # read every requested time already covered by the last accepted step
        if prev is not None and ts[nxt] <= prev[1]:
            t0s, t1s, ya, yb, ym, fa, fb = prev
            out.append(_interp(ya, yb, ym, fa, fb, t1s - t0s, (ts[nxt] - t0s) / (t1s - t0s)))
            nxt += 1
            continue
        steps += 1
        if steps > max_steps:
            raise RuntimeError("dopri5: max_steps exceeded")
        k = [f0]
        for a_i, b_i in zip(_ALPHA, _BETA):
            yi = y
            for bj, kj in zip(b_i, k):
                if bj != 0:
                    yi = yi + (dt * bj) * kj
            k.append(f(t + a_i * dt, yi))
        nfe += 6
        y1 = y
        err = torch.zeros_like(y)
        ymid = y
        for cs, ce, cm, kj in zip(_C_SOL, _C_ERR, _C_MID, k):
            if cs != 0:
                y1 = y1 + (dt * cs) * kj
            if ce != 0:
                err = err + (dt * ce) * kj
            if cm != 0:
                ymid = ymid + (dt * cm) * kj
        tol = atol + rtol * torch.maximum(y.abs(), y1.abs())
        ratio = _rms(err / tol)
        if not math.isfinite(ratio):
            # a NaN / Inf drift would be "rejected" for ever while min(10, max(nan, 0.2)) keeps GROWING the step: fail loudly instead
            # (torchdiffeq asserts on the same condition)
            raise RuntimeError(f"dopri5: non-finite error estimate at t = {t} (dt = {dt}): the model returned NaN or Inf")
        if t + dt == t:
            raise RuntimeError(f"dopri5: step size underflow at t = {t} (dt = {dt})")
        accept = ratio <= 1.0
        if accept:
            prev = (t, t + dt, y, y1, ymid, f0, k[-1])
            t, y, f0 = t + dt, y1, k[-1]   # FSAL: k7 = f(t1, y1)
        # step-size controller
        if ratio == 0:
            factor = 10.0
        else:
            dfac = 1.0 if ratio < 1 else 0.2
            factor = min(10.0, max(0.9 / ratio ** 0.2, dfac))
        dt = dt * factor
    return torch.stack(out), nfe
