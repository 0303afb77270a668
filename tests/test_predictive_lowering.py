"""Host half of the batched posterior predictive (`mininf_b200/engine/predictive.py::lower_predictive`)
without a GPU: the model of examples/predictive.md:22-38 is traced once and lowered to
``mnf_pred_site_t`` records; deterministic ``value`` sites are then evaluated from the records on
the host for every sample and compared with the reference's per-sample semantics
(mininf/core.py:548-584). The kernel that reads the same records is covered by the `-m gpu` tests."""
import pytest
import torch
from torch.distributions import Bernoulli, Gamma, Normal
from torch.distributions.constraints import nonnegative_integer

import mininf_b200 as mininf
from mininf_b200.engine import abi
from mininf_b200.engine.predictive import lower_predictive

CPU = torch.device("cpu")


def buffer(lower, pointer, count, stride=1):
    for tensor in lower.keepalive:
        if tensor.data_ptr() <= pointer < tensor.data_ptr() + max(tensor.numel() * tensor.element_size(), 1):
            offset = (pointer - tensor.data_ptr()) // tensor.element_size()
            return tensor.reshape(-1)[offset + stride * torch.arange(count)]
    raise AssertionError("a record points outside every tensor the lowering keeps alive")


def link(lower, L, z_row, count):
    i = torch.arange(count)
    a = L.a_const + (z_row[L.a_lat + L.a_stride * i] if L.a_lat >= 0 else torch.zeros(count))
    b = L.b_const + (z_row[L.b_lat + L.b_stride * i] if L.b_lat >= 0 else torch.zeros(count))
    x = buffer(lower, L.x, count, L.x_stride) if L.x else torch.ones(count)
    out = a + b * x
    return out.exp() if L.transform == abi.T_EXP else out


def evaluate_value_sites(sites, lower, z):
    """What the kernel does for PRED_VALUE records, sample by sample, in model order."""
    z = z.clone()
    for b in range(z.shape[0]):
        for site in sites:
            if site.kind != abi.PRED_VALUE:
                continue
            if site.X:
                X = buffer(lower, site.X, site.numel * site.ldx).reshape(site.numel, site.ldx)[:, :site.p]
                out = X @ z[b, site.theta_lat:site.theta_lat + site.p] + site.icpt_const
                if site.icpt_lat >= 0:
                    out = out + z[b, site.icpt_lat]
                out = out.exp() if site.transform == abi.T_EXP else out
            else:
                out = link(lower, site.param[0], z[b], site.numel)
            z[b, site.out_col:site.out_col + site.numel] = out
    return z


def test_predictive_example_lowers_to_a_dense_value_site_and_a_normal_draw():
    def model():
        n = mininf.value("n", 30, support=nonnegative_integer)
        p = mininf.value("p", 3, support=nonnegative_integer)
        x = mininf.sample("x", Normal(0, 1), n)
        X = mininf.value("X", x[:, None] ** torch.arange(p))
        theta = mininf.sample("theta", Normal(0, 1), p)
        prediction = mininf.value("prediction", X @ theta)
        sigma = mininf.sample("sigma", Gamma(2, 2))
        mininf.sample("y", Normal(prediction, sigma))

    torch.manual_seed(0)
    B, nlin = 6, 11
    theta, sigma = torch.randn(B, 3), 0.2 + torch.rand(B)
    lin = torch.linspace(-2.0, 2.0, nlin)
    sites, specs, order, traced, z, lower = lower_predictive(
        mininf.condition(model, n=nlin, x=lin), {"theta": theta, "sigma": sigma}, B, CPU)
    assert set(order) == {"n", "p", "x", "X", "theta", "prediction", "sigma", "y"}
    assert set(specs) == {"theta", "prediction", "sigma", "y"}          # x and X are data here
    prediction, y = sites
    assert (prediction.kind, prediction.p, prediction.numel, prediction.theta_lat) == \
        (abi.PRED_VALUE, 3, nlin, specs["theta"].offset)
    assert (y.kind, y.family, y.numel, y.out_col) == (abi.PRED_DRAW, abi.NORMAL, nlin, specs["y"].offset)
    assert (y.param[0].a_lat, y.param[0].a_stride) == (specs["prediction"].offset, 1)
    assert (y.param[1].a_lat, y.param[1].a_stride) == (specs["sigma"].offset, 0)
    # the given samples sit in their columns; the deterministic site evaluates to X @ theta_b
    torch.testing.assert_close(z[:, specs["theta"].offset:specs["theta"].offset + 3], theta)
    filled = evaluate_value_sites(sites, lower, z)
    X = lin[:, None] ** torch.arange(3)
    lo = specs["prediction"].offset
    torch.testing.assert_close(filled[:, lo:lo + nlin], theta @ X.T)


def test_widened_links_and_the_sigmoid_rewrite_in_the_predictive_lowering():
    x = torch.linspace(-1.0, 1.0, 9)

    def model():
        a = mininf.sample("a", Normal(0, 1))
        b = mininf.sample("b", Normal(0, 1))
        mininf.value("mean", a - b * x / 2)
        mininf.sample("k", Bernoulli(probs=torch.sigmoid(1.0 - b * x)))

    a, b = torch.randn(4), torch.randn(4)
    sites, specs, order, traced, z, lower = lower_predictive(model, {"a": a, "b": b}, 4, CPU)
    mean, k = sites
    filled = evaluate_value_sites(sites, lower, z)
    lo = specs["mean"].offset
    torch.testing.assert_close(filled[:, lo:lo + 9], a[:, None] - b[:, None] * x / 2)
    # Bernoulli(probs=sigmoid(eta)) draws as Bernoulli(logits=eta): the record holds eta = 1 - b x
    assert (k.kind, k.family, k.param[0].transform) == (abi.PRED_DRAW, abi.BERNOULLI_LOGITS, abi.T_ID)
    torch.testing.assert_close(link(lower, k.param[0], z[2], 9), 1.0 - b[2] * x)

    def sigmoid_value():
        a = mininf.sample("a", Normal(0, 1))
        mininf.value("p", torch.sigmoid(a + x))

    with pytest.raises(NotImplementedError, match="sigmoid"):
        lower_predictive(sigmoid_value, {"a": a}, 4, CPU)


def test_several_covariates_in_the_predictive_lowering():
    x1, x2 = torch.linspace(-1.0, 1.0, 9), torch.linspace(0.0, 2.0, 9) ** 2

    def model():
        a = mininf.sample("a", Normal(0, 1))
        b1 = mininf.sample("b1", Normal(0, 1))
        b2 = mininf.sample("b2", Normal(0, 1))
        mean = mininf.value("mean", a + b1 * x1 - b2 * x2 / 3)
        mininf.sample("y", Normal(mean, 0.5))

    a, b1, b2 = torch.randn(5), torch.randn(5), torch.randn(5)
    sites, specs, order, traced, z, lower = lower_predictive(model, {"a": a, "b1": b1, "b2": b2}, 5, CPU)
    mean, y = sites
    assert (mean.kind, mean.p, mean.theta_lat, mean.icpt_lat) == (abi.PRED_VALUE, 2, specs["b1"].offset, specs["a"].offset)
    filled = evaluate_value_sites(sites, lower, z)
    lo = specs["mean"].offset
    torch.testing.assert_close(filled[:, lo:lo + 9], a[:, None] + b1[:, None] * x1 - b2[:, None] * x2 / 3)
    assert (y.kind, y.family, y.param[0].a_lat, y.param[0].a_stride) == (abi.PRED_DRAW, abi.NORMAL, lo, 1)
