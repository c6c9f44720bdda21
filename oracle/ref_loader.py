"""ORACLE — TEST INFRASTRUCTURE ONLY.

Loads the *unmodified* reference modules from `/root/reference/src/scldm` in-process so that
golden vectors can be minted from the reference itself.  On the GPU box `/root/reference` does not
exist; the byte-identical copies staged by `oracle/build_ref.py` under `oracle/_ref/` (git-ignored,
shipped with the snapshot) are loaded instead - as the checker in tests and as the thing timed by
`bench.py --impl reference`, never on the product path.

`scldm/__init__.py` cannot be imported (it pulls `scvi` and package metadata), so a namespace
stub is registered instead and the submodules are imported directly.  Two third-party
packages the reference needs are absent from this image and are shimmed with ~10-line
restatements of their published behaviour:
  * `scvi.distributions.NegativeBinomial(mu, theta)`  -> Gamma-Poisson `.sample()`
  * `torchdiffeq.odeint(fn, x, t, method=..)`         -> fixed-grid euler / midpoint / heun2
Everything else that runs (`scldm.layers`, `scldm.nnets`, `scldm.stochastic_layers`,
`scldm.vae`, `scldm.transport`) is the reference's own code.
"""

from __future__ import annotations

import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "scldm")   # oracle/build_ref.py (travels to the GPU box)
REFERENCE_SRC = os.environ.get("SCLDM_REFERENCE_SRC") or ("/root/reference/src/scldm" if os.path.isdir("/root/reference/src/scldm") else _STAGED)


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_SRC, "nnets.py"))


def _install_shims() -> None:
    import torch

    if "scvi" not in sys.modules:
        scvi = types.ModuleType("scvi")
        dist = types.ModuleType("scvi.distributions")

        class NegativeBinomial(torch.distributions.Distribution):
            arg_constraints: dict = {}

            def __init__(self, mu, theta):
                self.mu, self.theta = mu, theta
                super().__init__(validate_args=False)

            @torch.no_grad()
            def sample(self, sample_shape=torch.Size()):
                rate = self.theta / self.mu
                gamma = torch.distributions.Gamma(concentration=self.theta.expand_as(self.mu), rate=rate).sample()
                return torch.poisson(torch.clamp(gamma, max=1e8))

        dist.NegativeBinomial = NegativeBinomial
        scvi.distributions = dist
        sys.modules["scvi"] = scvi
        sys.modules["scvi.distributions"] = dist

    if "torchdiffeq" not in sys.modules:
        tde = types.ModuleType("torchdiffeq")

        def odeint(fn, x, t, method="euler", atol=None, rtol=None):
            out = [x]
            for k in range(len(t) - 1):
                t0, t1 = t[k], t[k + 1]
                dt = t1 - t0
                if method == "euler":
                    x = x + dt * fn(t0, x)
                elif method == "midpoint":
                    x = x + dt * fn(t0 + 0.5 * dt, x + 0.5 * dt * fn(t0, x))
                elif method == "heun2":
                    k1 = fn(t0, x)
                    x = x + dt * 0.5 * (k1 + fn(t1, x + dt * k1))
                else:
                    raise NotImplementedError(f"shim has no '{method}' (adaptive solvers are not restated)")
                out.append(x)
            return torch.stack(out)

        tde.odeint = odeint
        sys.modules["torchdiffeq"] = tde


def load_reference():
    """Returns a namespace with the reference's hot-path modules, imported unmodified."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_SRC}")
    _install_shims()
    if "scldm" not in sys.modules:
        pkg = types.ModuleType("scldm")
        pkg.__path__ = [REFERENCE_SRC]
        sys.modules["scldm"] = pkg
    import scldm.layers  # noqa: F401
    import scldm.nnets  # noqa: F401
    import scldm.stochastic_layers  # noqa: F401
    import scldm.transport  # noqa: F401
    import scldm.vae  # noqa: F401

    ns = types.SimpleNamespace()
    ns.layers = sys.modules["scldm.layers"]
    ns.nnets = sys.modules["scldm.nnets"]
    ns.stochastic_layers = sys.modules["scldm.stochastic_layers"]
    ns.vae = sys.modules["scldm.vae"]
    ns.transport = sys.modules["scldm.transport"]
    return ns


def build_reference_dit(cfg, state_dict):
    ref = load_reference()
    m = ref.nnets.DiT(**cfg.kwargs())
    m.load_state_dict(state_dict, strict=True)
    return m.eval()


def build_reference_vae(cfg, state_dict):
    ref = load_reference()
    vae = ref.vae.TransformerVAE(
        encoder=ref.nnets.Encoder(**cfg.encoder_kwargs()),
        decoder=ref.nnets.Decoder(**cfg.decoder_kwargs()),
        decoder_head=ref.stochastic_layers.NegativeBinomialTransformerLayer(**cfg.head_kwargs()),
        input_layer=ref.layers.InputTransformerVAE(**cfg.input_kwargs()),
    )
    vae.load_state_dict(state_dict, strict=True)
    return vae.eval()


def reference_sample_fn(dit_cfg, vae_cfg, dit_sd, vae_sd, device="cpu", num_steps=50, method="euler"):
    """`LatentDiffusion.sample` (`models.py:766-819`) composed from the reference's OWN modules (its Lightning class is not
    importable): `Sampler.sample_ode(method, num_steps)` driving `DiT.forward_with_cfg`, then `TransformerVAE.decode(...).sample()`.
    Returns step(z0, labels, guidance_weight, genes, log_size_factors) -> (counts (2B,G), z (2B,16,16))."""
    import torch

    ref = load_reference()
    dit = build_reference_dit(dit_cfg, dit_sd).to(device)
    vae = build_reference_vae(vae_cfg, vae_sd).to(device)
    transport = ref.transport.create_transport(path_type="Linear", prediction="velocity", loss_weight="velocity", train_eps=1e-5, sample_eps=1e-5)
    fn = ref.transport.Sampler(transport).sample_ode(sampling_method=method, num_steps=num_steps)

    @torch.no_grad()
    def step(z0, labels, guidance_weight, genes, log_size_factors):
        model_fn = lambda x, t, **kw: dit.forward_with_cfg(x, t, **kw, cfg_scale=guidance_weight)  # noqa: E731
        z = fn(torch.cat([z0, z0]), model_fn, condition={k: torch.cat([v, v]) for k, v in labels.items()})[-1]
        lib = torch.exp(log_size_factors).view(-1, 1)
        nb = vae.decode(z, torch.cat([genes, genes]), torch.cat([lib, lib]))
        return nb.sample(), z

    return step
