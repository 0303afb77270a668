"""Links the tracer reduces to the affine form (`a - b*x/2`, `(c + a)*x/2`, `Bernoulli(probs=
sigmoid(c - b*x))`) through the CUDA engine against a float64 evaluation of the
reference algorithm (oracle/elbo.py). The lowering itself is covered without a GPU in
tests/test_trace_lowering.py; this file checks that the kernels score what the reference scores.
Needs a B200."""
import numpy as np
import pytest
import torch
from torch.distributions import Bernoulli, Gamma, Normal, Poisson

import mininf_b200 as mininf
from oracle import elbo

DEV = "cuda:0"
TRUTH = {"a": 0.3, "b": 0.5, "sigma": 0.7}


def make_data(n, seed):
    generator = torch.Generator().manual_seed(seed)
    x = torch.randn(n, generator=generator)
    w = TRUTH["a"] - TRUTH["b"] * x / 2 + TRUTH["sigma"] * torch.randn(n, generator=generator)
    k = torch.bernoulli(torch.sigmoid(1.0 - TRUTH["b"] * x), generator=generator)
    counts = torch.poisson(torch.exp((0.5 + TRUTH["a"]) * x / 2), generator=generator)
    return x, {"w": w, "k": k, "counts": counts}


def make_model(x):
    def model(m):
        a = m.sample("a", Normal(0, 1))
        b = m.sample("b", Normal(0, 1))
        sigma = m.sample("sigma", Gamma(2, 2))
        m.sample("w", Normal(a - b * x / 2, sigma))
        m.sample("k", Bernoulli(probs=torch.sigmoid(1.0 - b * x)))
        m.sample("counts", Poisson(torch.exp((0.5 + a) * x / 2)))
    return model


def make_approximation(device, dtype):
    leaves = {"a.loc": 0.1, "a.scale": 0.2, "b.loc": 0.3, "b.scale": 0.15, "sigma.concentration": 2.0, "sigma.rate": 2.5}
    leaves = {key: torch.tensor(value, device=device, dtype=dtype, requires_grad=True) for key, value in leaves.items()}
    approx = {"a": Normal(leaves["a.loc"], leaves["a.scale"]), "b": Normal(leaves["b.loc"], leaves["b.scale"]),
              "sigma": Gamma(leaves["sigma.concentration"], leaves["sigma.rate"])}
    return approx, leaves


def oracle_eval(x, data, noise, S):
    approx, leaves = make_approximation("cpu", torch.float64)
    expected = elbo.neg_elbo(make_model(x.double()), {k: v.double() for k, v in data.items()}, approx,
                             {k: v.double() for k, v in noise.items()}, S)
    expected.backward()
    return float(expected), {k: float(v.grad) for k, v in leaves.items()}


@pytest.mark.gpu
@pytest.mark.parametrize("n,S", [(100, 3), (5000, 8)])
def test_widened_links_against_the_float64_oracle(n, S):
    """Small-site kernel (n = 100) and fused site sweep (n = 5000): 1e-5 relative on the loss,
    the gradient tolerances of the other site tests."""
    torch.manual_seed(n)
    x, data = make_data(n, 17 + n)
    approx32, _ = make_approximation("cpu", torch.float32)
    noise = {name: elbo.draw_noise(dist, S) for name, dist in approx32.items()}
    expected, grads = oracle_eval(x, data, noise, S)

    approx, leaves = make_approximation(DEV, torch.float32)
    module = mininf.nn.EvidenceLowerBoundLoss(S, check="sync")
    model = make_model(x.to(DEV))
    conditioned = mininf.condition(lambda: model(mininf), **{k: v.to(DEV) for k, v in data.items()})
    loss = module(conditioned, approx, _noise={k: v.to(DEV) for k, v in noise.items()})
    loss.backward()
    assert (len(module.last_plan.sweep_groups) == 1) == (n >= 2048)
    assert abs(float(loss) - expected) <= 1e-5 * abs(expected)
    for key, leaf in leaves.items():
        np.testing.assert_allclose(float(leaf.grad), grads[key], rtol=2e-4, atol=1e-3, err_msg=key)


@pytest.mark.gpu
def test_the_feature_uncertainty_example_as_written_against_the_float64_oracle():
    """examples/regression-with-feature-uncertainty.md:28-38 literally (one latent feature per row,
    n = 30, `intercept + z * slope`, the noise scale conditioned on): the row-latent kernel with
    p = 1 on a latent far below the size at which latents become row latents by themselves."""
    torch.manual_seed(13)
    n, S = 30, 4

    def model(m):
        population_scale = m.sample("population_scale", Gamma(2, 2))
        z = m.sample("z", Normal(0, population_scale), n)
        noise_scale = m.sample("noise_scale", Gamma(2, 2))
        m.sample("x", Normal(z, noise_scale))
        intercept = m.sample("intercept", Normal(0, 1))
        slope = m.sample("slope", Normal(0, 1))
        m.sample("y", Poisson((intercept + z * slope).exp()))

    z_true = torch.randn(n)
    data = {"x": z_true + 0.3 * torch.randn(n), "y": torch.poisson(torch.exp(0.2 + 0.7 * z_true)),
            "noise_scale": torch.tensor(0.3)}
    z_loc = data["x"].clone()

    def approximation(device, dtype):
        values = {"z.loc": z_loc, "z.scale": 0.3 * torch.ones(n), "intercept.loc": torch.tensor(0.1),
                  "intercept.scale": torch.tensor(0.2), "slope.loc": torch.tensor(0.5), "slope.scale": torch.tensor(0.2),
                  "population_scale.concentration": torch.tensor(2.0), "population_scale.rate": torch.tensor(2.0)}
        leaves = {k: v.to(device=device, dtype=dtype).clone().requires_grad_() for k, v in values.items()}
        approx = {"z": Normal(leaves["z.loc"], leaves["z.scale"]),
                  "intercept": Normal(leaves["intercept.loc"], leaves["intercept.scale"]),
                  "slope": Normal(leaves["slope.loc"], leaves["slope.scale"]),
                  "population_scale": Gamma(leaves["population_scale.concentration"], leaves["population_scale.rate"])}
        return approx, leaves

    approx32, _ = approximation("cpu", torch.float32)
    noise = {name: elbo.draw_noise(dist, S) for name, dist in approx32.items()}
    approx64, leaves64 = approximation("cpu", torch.float64)
    expected = elbo.neg_elbo(model, {k: v.double() for k, v in data.items()}, approx64,
                             {k: v.double() for k, v in noise.items()}, S)
    expected.backward()

    approx, leaves = approximation(DEV, torch.float32)
    module = mininf.nn.EvidenceLowerBoundLoss(S, check="sync")
    conditioned = mininf.condition(lambda: model(mininf), **{k: v.to(DEV) for k, v in data.items()})
    loss = module(conditioned, approx, _noise={k: v.to(DEV) for k, v in noise.items()})
    loss.backward()
    assert list(module.last_plan.row_groups) == ["z"] and module.last_plan.row_groups["z"].p == 1
    assert abs(float(loss) - float(expected)) <= 1e-5 * abs(float(expected))
    for key, leaf in leaves.items():
        np.testing.assert_allclose(leaf.grad.cpu().numpy(), leaves64[key].grad.numpy(), rtol=2e-4, atol=1e-3, err_msg=key)
