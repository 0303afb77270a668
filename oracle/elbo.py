"""ORACLE - test infrastructure, not product code.

CPU restatement of ``mininf.nn.EvidenceLowerBoundLoss`` (mininf/nn.py:212-228) generalised to S
particles (S-particle ELBO := mean of S reference evaluations) and to externally
supplied reparameterisation noise, so the CUDA engine can be compared on identical draws.

``torch.distributions`` on the CPU is the numeric engine, exactly as in the reference; gradients
come from torch autograd. For a given noise tensor the reparameterised values and their
backward functions are the ones torch's own ``rsample`` uses:

  Normal  loc + eps * scale                                       TORCH normal.py:82-85
  Gamma   g / rate, backward through ``_standard_gamma_grad``     TORCH gamma.py:79-87
  Beta    first Dirichlet component, backward ``_dirichlet_grad`` TORCH dirichlet.py:16-35

Parity pinned by tests/test_oracle_golden.py (fixtures from the reference itself).
"""
from __future__ import annotations

from typing import Any, Callable, Dict, Optional

import torch
from torch import distributions as td

from . import handlers


class _GammaGiven(torch.autograd.Function):
    """A standard-gamma draw ``g`` as a function of its concentration (implicit gradient)."""

    @staticmethod
    def forward(ctx, concentration, g):  # type: ignore[override]
        ctx.save_for_backward(concentration, g)
        return g.clone()

    @staticmethod
    def backward(ctx, grad):  # type: ignore[override]
        concentration, g = ctx.saved_tensors
        return grad * torch._standard_gamma_grad(concentration, g), None


class _DirichletGiven(torch.autograd.Function):
    """A Dirichlet draw ``x`` as a function of its concentration (TORCH dirichlet.py:16-35)."""

    @staticmethod
    def forward(ctx, concentration, x):  # type: ignore[override]
        ctx.save_for_backward(concentration, x)
        return x.clone()

    @staticmethod
    def backward(ctx, grad):  # type: ignore[override]
        concentration, x = ctx.saved_tensors
        total = concentration.sum(-1, True).expand_as(concentration)
        g = torch._dirichlet_grad(x, concentration, total)
        return g * (grad - (x * grad).sum(-1, True)), None


def rsample_given(dist: td.Distribution, noise: torch.Tensor) -> torch.Tensor:
    """Reparameterised value of ``dist`` for externally drawn noise (see module docstring)."""
    if isinstance(dist, td.Normal):
        return dist.loc + noise * dist.scale
    if isinstance(dist, td.Gamma):
        shape = dist.batch_shape
        g = _GammaGiven.apply(dist.concentration.expand(shape), noise.expand(shape))
        value = g / dist.rate.expand(shape)
        value.detach().clamp_(min=torch.finfo(value.dtype).tiny)
        return value
    if isinstance(dist, td.Beta):
        conc = torch.stack([dist.concentration1, dist.concentration0], -1)
        x = torch.stack([noise, 1 - noise], -1)
        return _DirichletGiven.apply(conc, x).select(-1, 0)
    raise NotImplementedError(type(dist))


def draw_noise(dist: td.Distribution, n_particles: int) -> torch.Tensor:
    """Noise in the engine's convention, [S, *batch_shape]: eps / standard gamma / the Beta draw."""
    shape = (n_particles,) + tuple(dist.batch_shape)
    if isinstance(dist, td.Normal):
        return torch.randn(shape, dtype=dist.loc.dtype)
    if isinstance(dist, td.Gamma):
        return torch._standard_gamma(dist.concentration.detach().expand(shape))
    if isinstance(dist, td.Beta):
        return td.Beta(dist.concentration1.detach(), dist.concentration0.detach()).sample((n_particles,))
    raise NotImplementedError(type(dist))


def neg_elbo(model: Callable[[Any], Any], data: Dict[str, Any],
             approximation: Dict[str, td.Distribution], noise: Optional[Dict[str, torch.Tensor]],
             n_particles: int = 1, validate: bool = True) -> torch.Tensor:
    """``-(mean_s log p(data, z_s) + H[q])`` with z_s the s-th reparameterised draw.

    ``model(m)`` is a model function taking the API namespace; ``noise[name][s]`` is the noise of
    particle s (``None`` draws with torch's global CPU generator, particle-major, factors in
    dictionary order - the reference's stream, mininf/nn.py:145).
    """
    entropy = sum(dist.entropy().sum() for dist in approximation.values())
    total = 0.0
    for s in range(n_particles):
        if noise is None:
            draws = {name: dist.rsample() for name, dist in approximation.items()}
        else:
            draws = {name: rsample_given(dist, noise[name][s]) for name, dist in approximation.items()}
        conditioned = handlers.condition(handlers.condition(lambda: model(handlers), **data), **draws)
        scored = handlers.evaluate(conditioned, {}, validate=validate)
        total = total + sum(scored.values())
    return -(total / n_particles + entropy)
