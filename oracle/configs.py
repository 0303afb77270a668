"""ORACLE - test infrastructure, not product code.

The BASELINE.json configurations as model functions ``model(m)`` over an API namespace ``m``
(the reference ``mininf``, the product ``mininf_b200`` or ``oracle.handlers``), with their
synthetic-data recipes (256 row chunks, chunk c seeded ``seed0 + c``, so the
global data set is identical at every world size) and approximation families.

Model sources in the reference:
  coin         README.md:40-44, approximation README.md:59
  regression   tests/test_mininf.py:7-12 / examples/minibatch.md:24-33 (with ``no_log_prob`` X)
  logistic     examples/minibatch.md:24-33 with a Bernoulli(logits) likelihood under ``batch``
  missing      examples/missing-observations.md:131 (masked conditioning), Poisson + Normal sites
  features     examples/regression-with-feature-uncertainty.md:28-38 widened to p features
"""
from __future__ import annotations

import dataclasses
from typing import Any, Callable, Dict, List, Tuple

import torch
from torch.distributions import Bernoulli, Beta, Gamma, Normal, Poisson

N_CHUNKS = 256


def chunk_bounds(n: int, chunk: int) -> Tuple[int, int]:
    """Rows [lo, hi) of chunk ``chunk`` when ``n`` rows are split into N_CHUNKS pieces."""
    per = -(-n // N_CHUNKS)
    return min(chunk * per, n), min((chunk + 1) * per, n)


def rank_rows(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Rows owned by ``rank``: chunks [rank*256/W, (rank+1)*256/W)."""
    lo, _ = chunk_bounds(n, rank * N_CHUNKS // world)
    _, hi = chunk_bounds(n, (rank + 1) * N_CHUNKS // world - 1)
    return lo, hi


def _chunked(n: int, seed0: int, device: Any, fill: Callable[[int, torch.Generator], Dict[str, torch.Tensor]],
             rows: Tuple[int, int] | None = None) -> Dict[str, torch.Tensor]:
    """Generate per-row tensors chunk by chunk with one generator seed per chunk."""
    lo_all, hi_all = rows if rows is not None else (0, n)
    parts: Dict[str, List[torch.Tensor]] = {}
    generator = torch.Generator(device=device)
    for chunk in range(N_CHUNKS):
        lo, hi = chunk_bounds(n, chunk)
        if hi <= lo_all or lo >= hi_all or hi == lo:
            continue
        generator.manual_seed(seed0 + chunk)
        for key, tensor in fill(hi - lo, generator).items():
            parts.setdefault(key, []).append(tensor)
    return {key: torch.cat(tensors) for key, tensors in parts.items()}


@dataclasses.dataclass
class Config:
    name: str
    model: Callable[[Any], Any]                       # model(m)
    data: Dict[str, Any]                              # conditioned values
    families: Dict[str, Tuple[type, Dict[str, torch.Tensor]]]   # latent -> (cls, constrained init)
    extra: Dict[str, Any] = dataclasses.field(default_factory=dict)

    def approximation(self, device: Any = "cpu", dtype: torch.dtype = torch.float32,
                      requires_grad: bool = True) -> Tuple[Dict[str, Any], Dict[str, torch.Tensor]]:
        """Distributions built from fresh leaf tensors (the constrained parameters)."""
        dists, leaves = {}, {}
        for name, (cls, params) in self.families.items():
            kwargs = {}
            for key, value in params.items():
                leaf = value.detach().to(device=device, dtype=dtype).clone().requires_grad_(requires_grad)
                leaves[f"{name}.{key}"] = leaf
                kwargs[key] = leaf
            dists[name] = cls(**kwargs)
        return dists, leaves


# ---------------------------------------------------------------------------------------------
# C1: biased coin
# ---------------------------------------------------------------------------------------------
def coin(device: Any = "cpu") -> Config:
    n = 10

    def model(m: Any) -> None:
        theta = m.sample("theta", Beta(2, 2))
        m.sample("x", Bernoulli(theta), sample_shape=[n])

    # README.md:50-51 with torch.manual_seed(0) draws k = 9 heads
    x = torch.tensor([1., 1, 1, 1, 1, 1, 1, 1, 1, 0], device=device)
    families = {"theta": (Beta, {"concentration1": torch.tensor(2.0), "concentration0": torch.tensor(2.0)})}
    return Config("coin", model, {"x": x}, families, {"n": n})


# ---------------------------------------------------------------------------------------------
# C2: Bayesian linear regression (optionally with a latent noise scale, C2b)
# ---------------------------------------------------------------------------------------------
def regression_data(n: int, p: int, seed0: int = 2000, device: Any = "cpu",
                    rows: Tuple[int, int] | None = None) -> Dict[str, torch.Tensor]:
    g = torch.Generator(device=device)
    g.manual_seed(seed0 - 1)
    theta_true = torch.randn(p, generator=g, device=device) / p ** 0.5

    def fill(count: int, generator: torch.Generator) -> Dict[str, torch.Tensor]:
        X = torch.randn(count, p, generator=generator, device=device)
        y = X @ theta_true + torch.randn(count, generator=generator, device=device)
        return {"X": X, "y": y}

    out = _chunked(n, seed0, device, fill, rows)
    out["theta_true"] = theta_true
    return out


def regression(n: int, p: int = 64, sigma_latent: bool = False, device: Any = "cpu",
               seed0: int = 2000, rows: Tuple[int, int] | None = None,
               gen_device: Any = None) -> Config:
    """``gen_device`` is where the random numbers are drawn (CPU and CUDA generators differ);
    tests draw on the CPU and move the tensors so the oracle and the engine see identical data."""
    data = {k: v.to(device) for k, v in regression_data(n, p, seed0, gen_device or device, rows).items()}
    n_local = data["X"].shape[0]

    def model(m: Any) -> None:
        theta = m.sample("theta", Normal(0, 1), p)
        sigma = m.sample("sigma", Gamma(2, 2)) if sigma_latent else 1.0
        with m.no_log_prob():
            X = m.sample("X", Normal(0, 1), (n_local, p))
        m.sample("y", Normal(X @ theta, sigma))

    families: Dict[str, Any] = {"theta": (Normal, {"loc": torch.zeros(p), "scale": 0.1 * torch.ones(p)})}
    if sigma_latent:
        families["sigma"] = (Gamma, {"concentration": torch.tensor(2.0), "rate": torch.tensor(2.0)})
    return Config("regression", model, {"X": data["X"], "y": data["y"]}, families,
                  {"theta_true": data["theta_true"], "n": n, "p": p})


# ---------------------------------------------------------------------------------------------
# C3: minibatch logistic regression
# ---------------------------------------------------------------------------------------------
def logistic(n_declared: int, batch_rows: int, p: int = 256, device: Any = "cpu",
             seed0: int = 3000, batch_id: int = 0, gen_device: Any = None,
             intercept: bool = False) -> Config:
    target, device = device, gen_device or device
    g = torch.Generator(device=device)
    g.manual_seed(seed0 - 1)
    theta_true = torch.randn(p, generator=g, device=device) / p ** 0.5

    def fill(count: int, generator: torch.Generator) -> Dict[str, torch.Tensor]:
        X = torch.randn(count, p, generator=generator, device=device)
        y = torch.bernoulli(torch.sigmoid(X @ theta_true + (0.4 if intercept else 0.0)), generator=generator)
        return {"X": X, "y": y}

    data = {k: v.to(target) for k, v in _chunked(batch_rows, seed0 + 1000 * batch_id, device, fill).items()}
    theta_true = theta_true.to(target)

    def model(m: Any) -> None:
        theta = m.sample("theta", Normal(0, 1), p)
        with m.batch(n_declared):
            with m.no_log_prob():
                X = m.sample("X", Normal(0, 1), (n_declared, p))
            m.sample("y", Bernoulli(logits=X @ theta))

    def model_with_intercept(m: Any) -> None:
        alpha = m.sample("alpha", Normal(0, 2))
        theta = m.sample("theta", Normal(0, 1), p)
        with m.batch(n_declared):
            with m.no_log_prob():
                X = m.sample("X", Normal(0, 1), (n_declared, p))
            m.sample("y", Bernoulli(logits=alpha + X @ theta))

    families = {"theta": (Normal, {"loc": torch.zeros(p), "scale": 0.1 * torch.ones(p)})}
    if intercept:
        families["alpha"] = (Normal, {"loc": torch.tensor(0.1), "scale": torch.tensor(0.2)})
        model = model_with_intercept
    return Config("logistic", model, {"X": data["X"], "y": data["y"]}, families,
                  {"theta_true": theta_true, "n_declared": n_declared, "p": p})


# ---------------------------------------------------------------------------------------------
# C5: missing observations (masked Poisson and Normal sites sharing a covariate)
# ---------------------------------------------------------------------------------------------
def missing(n: int, device: Any = "cpu", seed0: int = 5000, missing_fraction: float = 0.3,
            rows: Tuple[int, int] | None = None, gen_device: Any = None) -> Config:
    target, device = device, gen_device or device
    truth = {"a": 0.3, "b": 0.5, "c": -0.2, "d": 0.8, "sigma": 0.7}

    def fill(count: int, generator: torch.Generator) -> Dict[str, torch.Tensor]:
        x = torch.randn(count, generator=generator, device=device)
        counts = torch.poisson(torch.exp(truth["a"] + truth["b"] * x), generator=generator)
        w = truth["c"] + truth["d"] * x + truth["sigma"] * torch.randn(count, generator=generator, device=device)
        m_counts = torch.rand(count, generator=generator, device=device) > missing_fraction
        m_w = torch.rand(count, generator=generator, device=device) > missing_fraction
        # finite fill in the holes (a NaN fill gives NaN gradients in the reference)
        return {"x": x, "counts": torch.where(m_counts, counts, 0.0), "w": torch.where(m_w, w, 0.0),
                "m_counts": m_counts, "m_w": m_w}

    raw = {k: v.to(target) for k, v in _chunked(n, seed0, device, fill, rows).items()}
    x = raw["x"]

    def model(m: Any) -> None:
        a = m.sample("a", Normal(0, 1))
        b = m.sample("b", Normal(0, 1))
        c = m.sample("c", Normal(0, 1))
        d = m.sample("d", Normal(0, 1))
        sigma = m.sample("sigma", Gamma(2, 2))
        m.sample("counts", Poisson((a + b * x).exp()))
        m.sample("w", Normal(c + d * x, sigma))

    data = {"counts": torch.masked.as_masked_tensor(raw["counts"], raw["m_counts"]),
            "w": torch.masked.as_masked_tensor(raw["w"], raw["m_w"])}
    families: Dict[str, Any] = {k: (Normal, {"loc": torch.tensor(0.1), "scale": torch.tensor(0.2)})
                                for k in "abcd"}
    families["sigma"] = (Gamma, {"concentration": torch.tensor(2.0), "rate": torch.tensor(2.0)})
    return Config("missing", model, data, families, {"raw": raw, "n": n})


# ---------------------------------------------------------------------------------------------
# C4: regression with feature uncertainty (per-observation latent features)
# ---------------------------------------------------------------------------------------------
def feature_uncertainty(n: int, p: int = 32, device: Any = "cpu", seed0: int = 4000,
                        rows: Tuple[int, int] | None = None, gen_device: Any = None) -> Config:
    target, device = device, gen_device or device
    noise_scale = 0.5
    g = torch.Generator(device=device)
    g.manual_seed(seed0 - 1)
    slope_true = torch.randn(p, generator=g, device=device) / p ** 0.5

    def fill(count: int, generator: torch.Generator) -> Dict[str, torch.Tensor]:
        z = torch.randn(count, p, generator=generator, device=device)
        x = z + noise_scale * torch.randn(count, p, generator=generator, device=device)
        y = torch.poisson(torch.exp(0.5 + z @ slope_true), generator=generator)
        return {"x": x, "y": y}

    data = {k: v.to(target) for k, v in _chunked(n, seed0, device, fill, rows).items()}
    n_local = data["x"].shape[0]

    def model(m: Any) -> None:
        population_scale = m.sample("population_scale", Gamma(2, 2))
        z = m.sample("z", Normal(0, population_scale), (n_local, p))
        m.sample("x", Normal(z, noise_scale))
        intercept = m.sample("intercept", Normal(0, 1))
        slope = m.sample("slope", Normal(0, 1), p)
        m.sample("y", Poisson((intercept + z @ slope).exp()))

    families: Dict[str, Any] = {
        "population_scale": (Gamma, {"concentration": torch.tensor(2.0), "rate": torch.tensor(2.0)}),
        "z": (Normal, {"loc": data["x"].detach().cpu().clone(), "scale": torch.ones(n_local, p)}),
        "intercept": (Normal, {"loc": torch.tensor(0.1), "scale": torch.tensor(0.2)}),
        "slope": (Normal, {"loc": torch.zeros(p), "scale": 0.2 * torch.ones(p)}),
    }
    return Config("features", model, data, families, {"n": n, "p": p, "slope_true": slope_true.to(target)})


# ---------------------------------------------------------------------------------------------
# Beyond BASELINE.json: link forms that REDUCE to the affine link (SURVEY.md §8 a10 widened) and
# the feature-uncertainty example exactly as the reference writes it
# ---------------------------------------------------------------------------------------------
def affine_links(n: int, device: Any = "cpu", seed0: int = 6000, gen_device: Any = None) -> Config:
    """Normal, Bernoulli and Poisson sites over one covariate whose links are written with a
    difference, a division, a scaled latent and a sigmoid: ``a - b*x/2``, ``sigmoid(1 - b*x)``
    (as Bernoulli probs), ``exp((0.5 + a)*x/2)``. Plain torch arithmetic in the reference
    (autograd differentiates whatever links latents to parameters, mininf/nn.py:223-225)."""
    target, device = device, gen_device or device
    truth = {"a": 0.3, "b": 0.5, "sigma": 0.7}
    generator = torch.Generator(device=device)
    generator.manual_seed(seed0)
    x = torch.randn(n, generator=generator, device=device)
    w = truth["a"] - truth["b"] * x / 2 + truth["sigma"] * torch.randn(n, generator=generator, device=device)
    k = torch.bernoulli(torch.sigmoid(1.0 - truth["b"] * x), generator=generator)
    counts = torch.poisson(torch.exp((0.5 + truth["a"]) * x / 2), generator=generator)
    x = x.to(target)

    def model(m: Any) -> None:
        a = m.sample("a", Normal(0, 1))
        b = m.sample("b", Normal(0, 1))
        sigma = m.sample("sigma", Gamma(2, 2))
        m.sample("w", Normal(a - b * x / 2, sigma))
        m.sample("k", Bernoulli(probs=torch.sigmoid(1.0 - b * x)))
        m.sample("counts", Poisson(torch.exp((0.5 + a) * x / 2)))

    data = {"w": w.to(target), "k": k.to(target), "counts": counts.to(target)}
    families: Dict[str, Any] = {
        "a": (Normal, {"loc": torch.tensor(0.1), "scale": torch.tensor(0.2)}),
        "b": (Normal, {"loc": torch.tensor(0.3), "scale": torch.tensor(0.15)}),
        "sigma": (Gamma, {"concentration": torch.tensor(2.0), "rate": torch.tensor(2.5)}),
    }
    return Config("affine_links", model, data, families, {"n": n, "x": x})


def feature_example(n: int = 30, device: Any = "cpu", seed0: int = 7000, gen_device: Any = None) -> Config:
    """examples/regression-with-feature-uncertainty.md:28-38 as written: ONE latent feature per
    observation, ``Poisson((intercept + z * slope).exp())``, and - as the example's text says - the
    observation noise scale conditioned on as a known value."""
    target, device = device, gen_device or device
    generator = torch.Generator(device=device)
    generator.manual_seed(seed0)
    z_true = torch.randn(n, generator=generator, device=device)
    x = z_true + 0.3 * torch.randn(n, generator=generator, device=device)
    y = torch.poisson(torch.exp(0.2 + 0.7 * z_true), generator=generator)

    def model(m: Any) -> None:
        population_scale = m.sample("population_scale", Gamma(2, 2))
        z = m.sample("z", Normal(0, population_scale), n)
        noise_scale = m.sample("noise_scale", Gamma(2, 2))
        m.sample("x", Normal(z, noise_scale))
        intercept = m.sample("intercept", Normal(0, 1))
        slope = m.sample("slope", Normal(0, 1))
        m.sample("y", Poisson((intercept + z * slope).exp()))

    data = {"x": x.to(target), "y": y.to(target), "noise_scale": torch.tensor(0.3, device=target)}
    families: Dict[str, Any] = {
        "population_scale": (Gamma, {"concentration": torch.tensor(2.0), "rate": torch.tensor(2.0)}),
        "z": (Normal, {"loc": x.detach().cpu().clone(), "scale": 0.3 * torch.ones(n)}),
        "intercept": (Normal, {"loc": torch.tensor(0.1), "scale": torch.tensor(0.2)}),
        "slope": (Normal, {"loc": torch.tensor(0.5), "scale": torch.tensor(0.2)}),
    }
    return Config("feature_example", model, data, families, {"n": n})


def several_covariates(n: int, device: Any = "cpu", seed0: int = 8000, gen_device: Any = None) -> Config:
    """``a + b1*x1 - b2*x2/3`` under a Normal and ``exp(0.1*(b1*x3 + b2*x2))`` under a Poisson:
    regression written with one named scalar coefficient per covariate instead of ``X @ theta``."""
    target, device = device, gen_device or device
    generator = torch.Generator(device=device)
    generator.manual_seed(seed0)
    x1, x2 = (torch.randn(n, generator=generator, device=device) for _ in range(2))
    x3 = torch.rand(n, generator=generator, device=device) + 0.5
    y = 0.2 + 0.5 * x1 - 0.3 * x2 / 3 + 0.8 * torch.randn(n, generator=generator, device=device)
    k = torch.poisson(torch.exp(0.1 * (0.5 * x3 + 0.3 * x2)), generator=generator)
    x1, x2, x3 = x1.to(target), x2.to(target), x3.to(target)

    def model(m: Any) -> None:
        a = m.sample("a", Normal(0, 1))
        b1 = m.sample("b1", Normal(0, 1))
        b2 = m.sample("b2", Normal(0, 1))
        s = m.sample("s", Gamma(2, 2))
        m.sample("y", Normal(a + b1 * x1 - b2 * x2 / 3, s))
        m.sample("k", Poisson(torch.exp(0.1 * (b1 * x3 + b2 * x2))))

    data = {"y": y.to(target), "k": k.to(target)}
    families: Dict[str, Any] = {
        "a": (Normal, {"loc": torch.tensor(0.1), "scale": torch.tensor(0.2)}),
        "b1": (Normal, {"loc": torch.tensor(0.4), "scale": torch.tensor(0.1)}),
        "b2": (Normal, {"loc": torch.tensor(0.2), "scale": torch.tensor(0.15)}),
        "s": (Gamma, {"concentration": torch.tensor(2.0), "rate": torch.tensor(2.5)}),
    }
    return Config("several_covariates", model, data, families, {"n": n})
