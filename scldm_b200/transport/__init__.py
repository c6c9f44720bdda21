"""Drop-in for `scldm.transport` restricted to what `LatentDiffusion` uses: Linear interpolant,
velocity prediction, ODE sampling (`transport/__init__.py:6-68`, `transport.py:324-369`)."""

from .transport import ModelType, PathType, Sampler, Transport, WeightType, create_transport

__all__ = ["ModelType", "PathType", "Sampler", "Transport", "WeightType", "create_transport"]
