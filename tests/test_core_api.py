"""The host-side mirror keeps the reference's interface contract (what the reference pins in
tests/test_core.py, tests/test_util.py and tests/test_mininf.py), checked here against
``mininf_b200`` as a drop-in for ``mininf``."""
import logging

import numpy as np
import pytest
import torch
from torch.distributions import constraints

import mininf_b200 as mininf
from mininf_b200.util import _normalize_shape, check_constraint, get_masked_data_with_dense_grad


# ---- singleton contexts (reference tests/test_core.py:10-57) -----------------------------------
def test_context_registry_errors():
    with pytest.raises(RuntimeError, match="must define"):
        with mininf.core.SingletonContextMixin():
            pass
    with mininf.State():
        with pytest.raises(RuntimeError, match="is already active."):
            with mininf.State():
                pass
    with mininf.State() as state:
        with pytest.raises(RuntimeError, match="Cannot reactivate"):
            with state:
                pass
    with pytest.raises(RuntimeError, match="no context is active."):
        with mininf.State():
            del mininf.State.INSTANCES["state"]
    with pytest.raises(RuntimeError, match="comprising {'a': <class 'int'>}> is active."):
        with mininf.State():
            other = mininf.State({"a": 3})
            mininf.State.INSTANCES["state"] = other
    assert mininf.State.INSTANCES.pop("state") is other


def test_get_instance():
    with mininf.State() as state:
        assert mininf.State.get_instance() is state
    assert mininf.State.get_instance() is None
    with pytest.raises(KeyError, match="context is active."):
        mininf.State.get_instance(True)

    class Conflict(mininf.core.SingletonContextMixin):
        SINGLETON_KEY = "state"

    with Conflict(), pytest.raises(TypeError, match="is not an instance of."):
        mininf.State.get_instance()


# ---- sampling, conditioning, validation ---------------------------------------------------------
def test_linear_regression_shapes():
    def linear_regression(n, p):
        features = mininf.sample("features", torch.distributions.Normal(0, 1), (n, p))
        coefs = mininf.sample("coefs", torch.distributions.Normal(0, 1), p)
        sigma = mininf.sample("sigma", torch.distributions.Gamma(2, 2))
        mininf.sample("outcomes", torch.distributions.Normal(features @ coefs, sigma))

    assert linear_regression(5, 2) is None
    with mininf.State() as state:
        linear_regression(50, 3)
    expected = {"features": (50, 3), "coefs": (3,), "sigma": (), "outcomes": (50,)}
    assert {key: tuple(val.shape) for key, val in state.items()} == expected
    with mininf.core.LogProbTracer() as log_prob, state:
        linear_regression(50, 3)
    assert {key: tuple(val[0].shape) for key, val in log_prob.items()} == expected


def test_log_prob_tracer_matches_distribution():
    distribution = torch.distributions.Uniform(0, 2)

    def model():
        mininf.sample("x", distribution, (7, 8))

    with mininf.State() as state:
        model()
        with mininf.core.LogProbTracer() as log_prob:
            model()
    np.testing.assert_allclose(log_prob["x"][0], distribution.log_prob(state["x"]))
    assert log_prob.total.ndim == 0
    str(log_prob), str(state)


def test_missing_and_non_tensor_values():
    with mininf.State() as state, mininf.core.LogProbTracer():
        with pytest.raises(ValueError, match="'a' is missing."):
            mininf.sample("a", None)
        state["a"] = "foobar"
        with pytest.raises(TypeError, match="Expected a tensor"):
            mininf.sample("a", None)


def test_condition_semantics():
    def model():
        x = mininf.sample("x", torch.distributions.Uniform(0, 1))
        mininf.sample("y", torch.distributions.Gamma(2, 2), 3)
        return x

    conditioned = mininf.condition(model, x=0.3)
    with mininf.State() as first:
        conditioned()
    with mininf.State() as second:
        conditioned()
    np.testing.assert_allclose(first["x"], 0.3)
    assert (first["y"] - second["y"]).abs().min() > 1e-12
    assert mininf.condition(model, x=0.25)() == 0.25
    subset = {"x": 0.1}
    assert mininf.condition(model, subset)() == 0.1
    assert mininf.condition(model, subset, x=0.7)() == 0.7      # keyword arguments win


@pytest.mark.parametrize("strict", [False, True])
def test_condition_conflict(strict):
    def model():
        return mininf.sample("x", torch.distributions.Normal(0, 1))

    inner = mininf.condition(model, x=0.1, _strict=strict)
    outer = mininf.condition(inner, x=0.7)
    if strict:
        with pytest.raises(ValueError, match="Cannot update"):
            outer()
    else:
        assert outer() == 0.1                                       # first conditioning wins


def test_validation_messages():
    def model():
        mininf.sample("x", torch.distributions.LKJCholesky(2, 4), (5, 7))

    with pytest.raises(TypeError, match="Expected a tensor"):
        mininf.condition(model, x="foo")()
    with mininf.State() as state, mininf.core.SampleTracer(_validate_parameters=False):
        mininf.condition(model, x="foo")()
        assert state["x"] == "foo"
    with pytest.raises(ValueError, match="Expected shape"):
        mininf.condition(model, x=torch.distributions.LKJCholesky(2, 4).sample())()
    with pytest.raises(ValueError, match="Expected shape"):
        mininf.condition(model, x=torch.distributions.LKJCholesky(2, 4).sample((5, 6)))()
    with pytest.raises(ValueError, match="is not in the support"):
        mininf.condition(model, x=torch.randn(5, 7, 2, 2))()


def test_duplicate_site_and_subset():
    def twice():
        mininf.sample("x", torch.distributions.Normal(0, 1))
        mininf.sample("x", torch.distributions.Normal(0, 1))

    with mininf.State():
        twice()
        with pytest.raises(RuntimeError, match="call `sample` twice"), mininf.core.LogProbTracer():
            twice()
    state = mininf.State({"a": torch.randn(3), "b": torch.randn(4), "c": torch.rand(7)})
    subset = state.subset("a", "b")
    assert set(subset) == {"a", "b"} and all(subset[k] is state[k] for k in subset)


# ---- masked data (reference tests/test_core.py:203-245, tests/test_util.py) ---------------------
def test_masked_log_prob_and_gradient():
    distribution = torch.distributions.Gamma(2, 2)

    def model():
        mininf.sample("x", distribution, (7, 8))

    with mininf.State() as state:
        model()
    original = state["x"].clone()
    mask = torch.rand(7, 8) < 0.5
    state["x"] = torch.masked.as_masked_tensor(torch.where(mask, original, -9), mask)
    with state, mininf.core.LogProbTracer() as log_prob:
        model()
    expected = distribution.log_prob(original)
    assert (log_prob["x"][0].get_data()[mask] == expected[mask]).all()
    torch.testing.assert_close(log_prob.total, expected[mask].sum())
    state["x"] = torch.masked.as_masked_tensor(torch.where(mask, -9, original), mask)
    with state, pytest.raises(ValueError, match="is not in the support GreaterThanEq"), \
            mininf.core.LogProbTracer(_validate_parameters=False):
        model()

    x = torch.randn(100, requires_grad=True)
    state = mininf.State(x=torch.masked.as_masked_tensor(x, torch.randn(100) < 0))
    with state, mininf.core.LogProbTracer() as log_prob:
        mininf.sample("x", torch.distributions.Normal(0, 1), [100])
    assert log_prob.total.grad_fn and log_prob.total.isfinite()
    log_prob.total.backward()
    assert x.grad is not None and x.grad.isfinite().all()


def test_masked_helpers():
    data = torch.tensor([1.0, -1.0, 2.0])
    mask = torch.tensor([True, False, True])
    masked = torch.masked.as_masked_tensor(data, mask)
    checked = check_constraint(constraints.positive, masked)
    assert checked.get_data()[mask].all() and not check_constraint(constraints.positive, data).all()
    leaf = torch.randn(5, requires_grad=True)
    m = torch.tensor([True, True, False, True, False])
    dense = get_masked_data_with_dense_grad(torch.masked.as_masked_tensor(leaf, m))
    dense.square().sum().backward()
    assert (leaf.grad[~m] == 0).all() and (leaf.grad[m] != 0).all()


@pytest.mark.parametrize("shape, expected", [(None, ()), (torch.Size([3]), (3,)), (4, (4,)),
                                             (torch.as_tensor(5), (5,)), ((2, 3), (2, 3)), ([7], (7,))])
def test_normalize_shape(shape, expected):
    assert _normalize_shape(shape) == expected and isinstance(_normalize_shape(shape), torch.Size)


# ---- values, batch, no_log_prob ------------------------------------------------------------------
def test_value_sites():
    def model():
        return mininf.value("x")

    with pytest.raises(ValueError, match="No default value given."):
        model()
    assert mininf.condition(model, x=3)() == 3
    with pytest.raises(ValueError, match=r"Expected shape \(\) for parameter"):
        mininf.condition(model, x=torch.randn(3))()
    default = torch.randn(5, 7)
    torch.testing.assert_close(mininf.value("y", value=default), default)
    assert torch.is_tensor(mininf.value("z", 3)) and mininf.value("z", 3.2) == 3.2
    with mininf.State(x=torch.randn(3, 4)), mininf.core.LogProbTracer() as log_prob:
        mininf.value("x", torch.randn(3, 4))
    assert "x" not in log_prob and log_prob.total == 0
    with pytest.raises(ValueError, match="is not in the specified support"):
        mininf.core.Value(-3, support=constraints.nonnegative)
    with pytest.raises(ValueError, match=r"is not in the support of Value\(support=GreaterThanEq"):
        mininf.condition(lambda: mininf.value("x", support=constraints.nonnegative), x=-2)()
    for tracer_type in (mininf.core.SampleTracer, mininf.core.LogProbTracer):
        with tracer_type(), mininf.State(x=torch.arange(3)), pytest.raises(ValueError, match=r"Expected shape \(5,\)"):
            mininf.value("x", shape=5)


def test_batch_scaling(caplog):
    distribution = torch.distributions.Normal(0, 1)

    def model(batch_shape):
        with mininf.batch(batch_shape):
            mininf.sample("x", distribution, (14, 9))

    for shape, declared, factor in [((7, 9), 14, 2), ((14, 1), (14, 9), 9), ((2, 3), (14, 9), 21)]:
        x = distribution.sample(shape)
        with mininf.State(x=x), mininf.core.LogProbTracer() as log_prob:
            model(declared)
        torch.testing.assert_close(log_prob.total, distribution.log_prob(x).sum() * factor)
    x = distribution.sample([15, 9])
    with caplog.at_level(logging.WARNING), mininf.State(x=x), mininf.core.LogProbTracer() as log_prob:
        model((14, 9))
    torch.testing.assert_close(log_prob.total, distribution.log_prob(x).sum() * 14 / 15)
    assert "exceeds expected batch shape" in caplog.messages[0]
    with mininf.State(x=x), pytest.raises(ValueError, match="has more dimensions"), mininf.core.LogProbTracer():
        model([7, 9, 2])
    with mininf.State(x=torch.masked.as_masked_tensor(x, x > 0)), mininf.core.LogProbTracer() as log_prob, \
            pytest.raises(ValueError, match="not supported for masked data"):
        model([7])
        log_prob.total


def test_adaptive_batch_with_index_values():
    def model():
        n = mininf.value("n")
        with mininf.batch(n):
            i = mininf.value("i", shape=n)
            predictor = mininf.sample("x", torch.distributions.Normal(torch.ones(n), 1))[i]
            mininf.sample("y", torch.distributions.Normal(predictor, 1))

    n = 7
    x = torch.randn(n)
    y = torch.randn(n) + x
    i = torch.as_tensor([2, 3, 6])
    with mininf.State(n=n, x=x, y=y[i], i=i), mininf.core.LogProbTracer() as log_prob:
        model()
    torch.testing.assert_close(log_prob.contribution("x"), torch.distributions.Normal(1, 1).log_prob(x).sum())
    torch.testing.assert_close(log_prob.contribution("y"),
                               torch.distributions.Normal(x[i], 1).log_prob(y[i]).sum() * n / i.numel())


def test_no_log_prob_and_broadcast():
    def model():
        x = mininf.sample("x", torch.distributions.Normal(0, 1), (3, 4))
        with mininf.no_log_prob():
            y = mininf.sample("y", torch.distributions.Gamma(2, 2), (4, 5))
        return x @ y

    with mininf.State():
        z = model()
        with mininf.core.LogProbTracer() as log_prob:
            torch.testing.assert_close(model(), z)
        assert "y" not in log_prob

    def other():
        a = mininf.value("a")
        x = mininf.sample("x", torch.distributions.Normal(0, 1))
        mininf.value("y", x + a)

    x = torch.randn(7)
    states = mininf.broadcast_samples(mininf.condition(other, a=1.3), x=x)
    torch.testing.assert_close(states["y"], x + 1.3)


def test_inverse_gamma_is_importable_like_the_reference() -> None:
    # tests/test_distributions.py:5-6 of the reference
    from mininf_b200.distributions import InverseGamma
    assert InverseGamma(torch.rand(3, 1), torch.rand(4)).sample([7]).shape == (7, 3, 4)
