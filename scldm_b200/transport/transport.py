from __future__ import annotations

import enum

import torch

from .. import ops


class ModelType(enum.Enum):
    NOISE = enum.auto()
    SCORE = enum.auto()
    VELOCITY = enum.auto()


class PathType(enum.Enum):
    LINEAR = enum.auto()
    GVP = enum.auto()
    VP = enum.auto()


class WeightType(enum.Enum):
    NONE = enum.auto()
    VELOCITY = enum.auto()
    LIKELIHOOD = enum.auto()


class Transport:
    """State of the flow (`transport.py:35-57`).  Only LINEAR + VELOCITY is on the generation path."""

    def __init__(self, *, model_type, path_type, loss_type, train_eps, sample_eps):
        if path_type is not PathType.LINEAR or model_type is not ModelType.VELOCITY:
            raise NotImplementedError("scldm_b200 covers the Linear path with velocity prediction (ldm_base.yaml:29-35)")
        self.model_type, self.path_type, self.loss_type = model_type, path_type, loss_type
        self.train_eps, self.sample_eps = train_eps, sample_eps

    def check_interval(self, train_eps, sample_eps, *, diffusion_form="SBDM", sde=False, reverse=False, eval=False,
                       last_step_size=0.0):
        """`Transport.check_interval` (`transport.py:69-95`) for the velocity/Linear ODE case: t0=0, t1=1."""
        t0, t1 = 0, 1
        if sde:   # ICPlan branch of the reference (`transport.py:85-89`): a first semi-implicit step for SBDM, the last step held back
            eps = train_eps if not eval else sample_eps
            t0 = eps if diffusion_form == "SBDM" else 0
            t1 = 1 - eps if last_step_size == 0 else 1 - last_step_size
        if reverse:
            t0, t1 = 1 - t0, 1 - t1
        return t0, t1

    def sample(self, x1: torch.Tensor):
        """`Transport.sample` (`transport.py:97-108`): x0 ~ N(0,I), t ~ U(t0,t1) drawn on the CPU then moved (as the reference)."""
        x0 = torch.randn_like(x1)
        t0, t1 = self.check_interval(self.train_eps, self.sample_eps)
        t = torch.rand((x1.shape[0],)) * (t1 - t0) + t0
        return t.to(x1), x0, x1

    def training_losses(self, model, x1: torch.Tensor, model_kwargs: dict | None = None, *, t: torch.Tensor | None = None,
                        x0: torch.Tensor | None = None) -> dict:
        """`Transport.training_losses` (`transport.py:110-150`) for the Linear path + velocity prediction:
        x_t = t x1 + (1-t) x0 (`path.py:129-151`), loss = mean_flat((model(x_t, t) - (x1 - x0))^2).
        Forward only (validation / `shared_step`, `models.py:673-704`); `t` / `x0` may be injected for parity tests."""
        model_kwargs = model_kwargs or {}
        if t is None or x0 is None:
            ts, x0s, _ = self.sample(x1)
            t = ts if t is None else t
            x0 = x0s if x0 is None else x0
        te = t.view(-1, *([1] * (x1.dim() - 1)))
        xt = te * x1 + (1 - te) * x0
        ut = x1 - x0
        out = model(xt, t, **model_kwargs)
        assert out.shape == xt.shape
        return {"pred": out, "loss": mean_flat((out - ut) ** 2)}


def mean_flat(x: torch.Tensor) -> torch.Tensor:
    """mean over all non-batch dims (`transport/utils.py`)."""
    return x.flatten(1).mean(1)


def create_transport(path_type="Linear", prediction="velocity", loss_weight=None, train_eps=None, sample_eps=None) -> Transport:
    """`create_transport` (`transport/__init__.py:6-68`); velocity & Linear forces both eps to 0 (`:55-57`)."""
    model_type = {"noise": ModelType.NOISE, "score": ModelType.SCORE}.get(prediction, ModelType.VELOCITY)
    loss_type = {"velocity": WeightType.VELOCITY, "likelihood": WeightType.LIKELIHOOD}.get(loss_weight, WeightType.NONE)
    ptype = {"Linear": PathType.LINEAR, "GVP": PathType.GVP, "VP": PathType.VP}[path_type]
    return Transport(model_type=model_type, path_type=ptype, loss_type=loss_type, train_eps=0, sample_eps=0)


class FusedCFGModel:
    """Callable handed to the sampler in place of the reference's
    `lambda x, t, **kw: dit.forward_with_cfg(x, t, **kw, cfg_scale=w)` (`models.py:809-811`).
    Being a recognisable object (not an opaque lambda) lets `Sampler.sample_ode` run the whole
    time loop inside one C-ABI call instead of a Python loop of model calls."""

    def __init__(self, dit, cfg_scale: dict[str, float] | None):
        self.dit, self.cfg_scale = dit, cfg_scale

    def __call__(self, x, t, condition=None):
        return self.dit.forward_with_cfg(x, t, condition=condition, cfg_scale=self.cfg_scale)


class FusedForwardModel:
    """Callable for sampling WITHOUT guidance: `lambda x, t, **kw: dit.forward(x, t, **kw, force_drop_ids=False)` as a recognisable
    object, so that fixed-grid solvers run the whole loop in one C-ABI call (BASELINE configs[0]: plain conditional sampling)."""

    def __init__(self, dit):
        self.dit = dit

    def __call__(self, x, t, condition=None):
        return self.dit.forward(x, t, condition, force_drop_ids=False)


class Sampler:
    """`Sampler` (`transport.py:206-225, 324-369`): `sample_ode(...)` returns fn(x, model, **model_kwargs)."""

    FIXED = ("euler", "heun2", "midpoint")
    ADAPTIVE = ("dopri5",)

    def __init__(self, transport: Transport):
        self.transport = transport

    def sample_ode(self, *, sampling_method="dopri5", num_steps=50, atol=1e-5, rtol=1e-5, reverse=False, return_trajectory=False):
        """`return_trajectory=True` returns all `num_steps` states `(T, ...)` like the reference (`integrators.py:100-112` stacks
        every grid point); the default keeps only the end points `(2, ...)`, because `LatentDiffusion.sample` reads `[-1]` alone
        (`models.py:812`) and the whole solve then runs in ONE C-ABI call."""
        if reverse:
            raise NotImplementedError("reverse-time ODE is outside the generation path")
        method = sampling_method.lower()
        if method not in self.FIXED + self.ADAPTIVE:
            raise NotImplementedError(f"sampling_method='{sampling_method}': implemented: {self.FIXED + self.ADAPTIVE}")
        t0, t1 = self.transport.check_interval(self.transport.train_eps, self.transport.sample_eps, sde=False, eval=True)
        grid = torch.linspace(t0, t1, num_steps)  # `ode.__init__`, integrators.py:95

        def solve_fixed(plan, x0):
            if not return_trajectory:
                return torch.stack([x0, ops.dit_sample_ode(plan, x0.clone(), grid, method)])
            # one C-ABI call per grid interval (same kernels and fp32 step sizes; each call starts from the fp32 input projection)
            states, xk = [x0], x0.clone()
            for k in range(num_steps - 1):
                xk = ops.dit_sample_ode(plan, xk, grid[k:k + 2], method)
                states.append(xk.clone())
            return torch.stack(states)

        def sample(x, model, **model_kwargs):
            """Returns the trajectory end points stacked as (2, ...): [x(t0), x(t1)], or all `num_steps` states with
            `return_trajectory=True` (the reference returns all of them but `LatentDiffusion.sample` only reads `[-1]`,
            `models.py:812`)."""
            x0 = x.contiguous().float()
            if method in self.ADAPTIVE:
                return self._sample_adaptive(x0, model, grid, atol, rtol, **model_kwargs)
            if isinstance(model, FusedCFGModel):
                half = x0.shape[0] // 2
                plan = model_kwargs.get("_plan")   # prebuilt by LatentDiffusion.sample (plan building synchronises with the device)
                if plan is None:
                    plan, _ = model.dit.cfg_plan(model_kwargs.get("condition"), model.cfg_scale, half, x0.device, shared_time=True)
                return solve_fixed(plan, x0)
            if isinstance(model, FusedForwardModel):
                plan = model.dit.forward_plan(model_kwargs.get("condition"), x0.shape[0], x0.device)
                if plan is not None:
                    return solve_fixed(plan, x0)
            # generic callable: host-driven loop with the same fixed-grid formulas (slow path, still CUDA model calls)
            xk = x0
            states = [x0]
            for k in range(num_steps - 1):
                ta, tb = grid[k].item(), grid[k + 1].item()
                dt = torch.tensor(tb, dtype=torch.float32) - torch.tensor(ta, dtype=torch.float32)
                dt = float(dt)
                tv = lambda v: torch.full((xk.shape[0],), v, dtype=torch.float32, device=xk.device)  # noqa: E731
                k1 = model(xk, tv(ta), **model_kwargs)
                if method == "euler":
                    xk = xk + dt * k1
                elif method == "heun2":
                    xk = xk + dt * 0.5 * (k1 + model(xk + dt * k1, tv(tb), **model_kwargs))
                else:
                    xk = xk + dt * model(xk + 0.5 * dt * k1, tv(ta + 0.5 * dt), **model_kwargs)
                states.append(xk)
            return torch.stack(states if return_trajectory else [x0, xk])

        return sample

    DIFFUSION_FORMS = {"constant": 0, "SBDM": 1, "sigma": 2, "linear": 3, "decreasing": 4, "inccreasing-decreasing": 5}

    def sample_sde(self, *, sampling_method="Euler", diffusion_form="SBDM", diffusion_norm=1.0, last_step="Mean", last_step_size=0.04, num_steps=250,
                   seed: int = 0):
        """`Sampler.sample_sde` (`transport.py:269-322`) for the Linear path with a velocity model: Euler-Maruyama / Heun steps of
        `integrators.sde` (`integrators.py:7-75`) on the uniform grid linspace(t0, t1, num_steps), then the last step ("Mean", "Euler",
        "Tweedie" or None).  Returns fn(init, model, **model_kwargs) -> list of `num_steps` states like the reference.  Every model
        evaluation is one batched launch sequence on the device; drift / noise / update run in the library's own kernels.  The
        Brownian increments come from Philox streams keyed by (seed, cell, step) - pass `noise=[...]` (num_steps - 1 tensors) to inject
        them.  NOTE (reference behaviour): with the eps = 0 that `create_transport` forces for Linear + velocity, the default
        diffusion_form="SBDM" = (1-t)/t is infinite at t0 = 0, in the reference as here; use "sigma", "linear", "constant", ..."""
        if sampling_method not in ("Euler", "Heun"):
            raise NotImplementedError("Sampler type not implemented.")
        if diffusion_form not in self.DIFFUSION_FORMS:
            raise NotImplementedError(f"Diffusion form {diffusion_form} not implemented")
        if last_step not in (None, "Mean", "Tweedie", "Euler"):
            raise NotImplementedError()
        if last_step is None:
            last_step_size = 0.0
        form = self.DIFFUSION_FORMS[diffusion_form]
        t0, t1 = self.transport.check_interval(self.transport.train_eps, self.transport.sample_eps, diffusion_form=diffusion_form, sde=True, eval=True,
                                               reverse=False, last_step_size=last_step_size)
        assert t0 < t1, "SDE sampler has to be in forward time"
        grid = torch.linspace(t0, t1, num_steps)
        dt = float(grid[1] - grid[0])

        def _sample(init, model, noise=None, cell_offset: int = 0, **model_kwargs):
            x = init.contiguous().float()
            per_cell = x[0].numel()
            if isinstance(model, FusedCFGModel):
                plan, _ = model.dit.cfg_plan(model_kwargs.get("condition"), model.cfg_scale, x.shape[0] // 2, x.device, shared_time=True)
                velocity = lambda xx, t: ops.dit_forward_shared_t(plan, xx.contiguous(), float(t))  # noqa: E731
            else:
                velocity = lambda xx, t: model(xx, torch.full((xx.shape[0],), float(t), dtype=torch.float32, device=xx.device), **model_kwargs)  # noqa: E731
            drift = lambda xx, t: ops.sde_drift(velocity(xx, t), xx, float(t), form, diffusion_norm)  # noqa: E731
            xs = []
            for k in range(num_steps - 1):
                t = float(grid[k])
                w = None if noise is None else noise[k].to(x)
                if sampling_method == "Euler":      # x' = x + drift dt + sqrt(2 D) dW     (integrators.py:27-35)
                    x = ops.sde_kick(ops.axpy2(x, dt, drift(x, t)), w, t, dt, form, diffusion_norm, seed, cell_offset, per_cell, k, x_for_diffusion=x)
                else:                               # Heun (integrators.py:37-46)
                    xhat = ops.sde_kick(x, w, t, dt, form, diffusion_norm, seed, cell_offset, per_cell, k)
                    k1 = drift(xhat, t)
                    k2 = drift(ops.axpy2(xhat, dt, k1), t + dt)
                    x = ops.axpy2(xhat, 0.5 * dt, k1, 0.5 * dt, k2)
                xs.append(x)
            if last_step == "Mean":
                x = ops.axpy2(xs[-1], last_step_size, drift(xs[-1], t1))
            elif last_step == "Euler":
                x = ops.axpy2(xs[-1], last_step_size, velocity(xs[-1], t1))
            elif last_step == "Tweedie":            # x / alpha + sigma^2 / alpha * score, alpha = t1, sigma = 1 - t1
                v = velocity(xs[-1], t1)
                score = (t1 * v - xs[-1]) / (1 - t1)
                x = xs[-1] / t1 + (1 - t1) ** 2 / t1 * score
            else:
                x = xs[-1]
            xs.append(x)
            assert len(xs) == num_steps, "Samples does not match the number of steps"
            return xs

        return _sample

    def _sample_adaptive(self, x0, model, grid, atol, rtol, **model_kwargs):
        """dopri5 (the reference's default, `transport.py:327`): host-side step controller (as in torchdiffeq), function
        evaluations on the GPU.  With a `FusedCFGModel` every evaluation is ONE batched launch sequence (shared-time CFG
        plan built once); returns all `num_steps` requested states like the reference."""
        from .adaptive import dopri5

        if isinstance(model, FusedCFGModel):
            half = x0.shape[0] // 2
            plan, _ = model.dit.cfg_plan(model_kwargs.get("condition"), model.cfg_scale, half, x0.device, shared_time=True)

            def f(t, y):
                return ops.dit_forward_shared_t(plan, y.contiguous(), float(t))
        else:
            def f(t, y):
                tv = torch.full((y.shape[0],), float(t), dtype=torch.float32, device=y.device)
                return model(y, tv, **model_kwargs)

        traj, nfe = dopri5(f, x0, grid.tolist(), rtol=rtol, atol=atol)
        self.last_nfe = nfe
        return traj
