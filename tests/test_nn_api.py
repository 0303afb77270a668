"""Interface contract of the variational layer (reference tests/test_nn.py) on CPU: the module
classes behave like the reference's; the engine-backed loss refuses CPU tensors loudly."""
import numpy as np
import pytest
import torch
from torch import distributions

import mininf_b200 as mininf
from mininf_b200.nn import (EvidenceLowerBoundLoss, FactorizedDistribution, LogLikelihoodLoss,
                            ParameterizedDistribution, ParameterizedFactorizedDistribution)


@pytest.mark.parametrize("cls, params, const, grads", [
    (distributions.Normal, {"loc": 0.0, "scale": 1.0}, set(), {"loc", "scale"}),
    (distributions.Normal, {"loc": torch.randn(3), "scale": torch.ones(2, 1)}, {"loc"}, {"scale"}),
    (distributions.LKJCholesky, {"dim": 3, "concentration": 9}, set(), {"concentration"}),
])
def test_parameterized_distribution(cls, params, const, grads):
    module = ParameterizedDistribution(cls, _const=const, **params)
    dist = module()
    assert isinstance(dist, cls)
    log_prob = dist.log_prob(dist.sample())
    assert torch.isfinite(log_prob).all()
    log_prob.sum().backward()
    assert set(module.distribution_parameters) == grads
    assert all(module.distribution_parameters[name].grad is not None for name in grads)


def test_unconstrained_storage_and_hidden_parameters():
    module = ParameterizedDistribution(distributions.Gamma, concentration=3.0, rate=2.0)
    np.testing.assert_allclose(module.distribution_parameters["concentration"].item(), np.log(3.0), rtol=1e-6)
    dist = ParameterizedDistribution(distributions.Normal, loc=0.0, scale=1.0)()
    assert not isinstance(dist.loc, torch.nn.Parameter) and not isinstance(dist.scale, torch.nn.Parameter)


@pytest.mark.parametrize("clone", [False, True])
def test_clone_isolates_inputs(clone):
    loc = torch.randn(3)
    copied = loc.clone()
    module = ParameterizedDistribution(distributions.Normal, loc=loc, scale=1, _clone=clone)
    optimizer = torch.optim.Adam(module.parameters(), 0.1)
    module().rsample().square().sum().backward()
    optimizer.step()
    if clone:
        np.testing.assert_allclose(loc, copied)
    else:
        assert ((loc - copied).abs() > 1e-6).all()


def test_factorized_distributions():
    x = distributions.Normal(0, 1)
    y = distributions.Gamma(2 * torch.ones(5), 2)
    joint = FactorizedDistribution(x=x, y=y)
    assert joint.entropy() == x.entropy() + y.entropy().sum()
    assert joint.rsample([3])["y"].shape == (3, 5) and joint.sample([7])["x"].shape == (7,)
    module = ParameterizedFactorizedDistribution(
        {"a": ParameterizedDistribution(distributions.Normal, loc=0.0, scale=1.0)},
        b=ParameterizedDistribution(distributions.Gamma, concentration=3.0, rate=2.0))
    assert set(module) == {"a", "b"}
    assert isinstance(module(), FactorizedDistribution) and isinstance(module()["b"], distributions.Gamma)


def test_elbo_rejects_non_dictionaries_and_cpu_tensors():
    loss = EvidenceLowerBoundLoss()
    with pytest.raises(TypeError, match="dictionaries of tensors"):
        loss(None, distributions.Normal(0, 1))

    def model():
        mininf.sample("x", distributions.Normal(0, 1), 3)

    approximation = ParameterizedDistribution(distributions.Normal, loc=0.0, scale=torch.ones(3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        loss(model, {"x": approximation()})
    with pytest.raises(ValueError):
        EvidenceLowerBoundLoss(dense_precision="bf16")


def test_log_likelihood_loss_has_no_cpu_fallback():
    """mininf/nn.py:231-257 on the engine: CUDA tensors only (GPU parity: test_engine_gpu.py)."""
    def model():
        mininf.sample("x", distributions.Normal(0, 1), 3)

    estimate = torch.nn.Parameter(torch.ones(3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        LogLikelihoodLoss()(model, {"x": estimate})
