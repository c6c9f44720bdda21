"""Drop-in for `scldm.stochastic_layers.NegativeBinomialTransformerLayer` (weights only; the head is
fused into the MCAB decode kernel) and a minimal NB distribution object standing in for
`scvi.distributions.NegativeBinomial` (scvi is a third-party dependency absent from this image)."""

from __future__ import annotations

import torch
import torch.nn as nn


class NegativeBinomialTransformerLayer(nn.Module):
    """weights of the NB head (`stochastic_layers.py:76-100`): params Linear(E->1|2), theta Embedding(G+1,1)."""

    def __init__(self, *, n_genes: int, shared_theta: bool = False, n_embed: int | None = None, norm_layer: str = "layernorm",
                 layernorm_eps: float = 1e-8, eps_: float = 1e-6, t: float = 1.0):
        super().__init__()
        self.shared_theta = shared_theta
        if shared_theta:
            self.theta = nn.Embedding(n_genes + 1, 1)
            nn.init.ones_(self.theta.weight)
            self.params = nn.Linear(n_embed, 1, bias=True)
        else:
            self.theta = None
            self.params = nn.Linear(n_embed, 2, bias=True)
        self.eps_ = eps_
        if t != 1.0:
            # the reference computes softmax(logit / t) (`stochastic_layers.py:86,115`); the fused head has no temperature input and
            # every shipped config uses the default - refuse rather than return wrong means
            raise NotImplementedError("NegativeBinomialTransformerLayer: softmax temperature t != 1.0 is not implemented in the fused NB head")
        self.t = t

    def forward(self, counts, genes, library_size):
        raise RuntimeError("the NB head is fused into the MCAB decode kernel; use TransformerVAE.decode")


class NegativeBinomial:
    """NB(mu, theta) in scvi's mean/inverse-dispersion parameterisation.  `.sample()` draws
    Poisson(Gamma(theta, rate=theta/mu)) on the GPU with the library's Philox sampler."""

    def __init__(self, mu: torch.Tensor, theta: torch.Tensor, _sampler=None):
        self.mu = mu
        self.theta = theta
        self._sampler = _sampler

    @property
    def mean(self):
        return self.mu

    @property
    def variance(self):
        return self.mu + self.mu**2 / self.theta

    def sample(self, sample_shape=torch.Size()):
        """One draw per (cell, gene); every call uses fresh Philox counters (as scvi's sampler uses fresh randomness)."""
        if self._sampler is None:
            raise RuntimeError("this NegativeBinomial was not produced by TransformerVAE.decode")
        if len(tuple(sample_shape)) != 0:
            raise NotImplementedError("NegativeBinomial.sample: only sample_shape=() (one draw per entry) is implemented")
        return self._sampler()
