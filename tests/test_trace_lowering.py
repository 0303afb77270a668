"""Host logic of the engine without a GPU: tracing a model into a site table and lowering it to
the C-ABI structures (``Plan(..., dry_run=True)`` touches no device)."""
import pytest
import torch
from torch.distributions import Bernoulli, Beta, Gamma, Normal, Poisson

import mininf_b200 as mininf
from mininf_b200.engine import abi
from mininf_b200.engine.plan import LatentSpec, Plan
from mininf_b200.engine.trace import Affine, Dense, LatentRef, LinkTensor, SiteTableTracer
from oracle import configs

CPU = torch.device("cpu")


def trace(model, latents, data):
    draws, specs, offset = {}, [], 0
    for name, (family, value) in latents.items():
        numel = max(value.numel(), 1)
        ref = LatentRef(name, 0) if numel == 1 else LatentRef(name)
        draws[name] = LinkTensor.wrap(value, Affine(a_lat=ref))
        specs.append(LatentSpec(name, family, value.shape, numel, offset))
        offset += numel
    with SiteTableTracer() as tracer:
        mininf.condition(mininf.condition(model, **data), **draws)()
    return tracer.sites, specs


def test_regression_lowers_to_one_dense_site_and_two_priors():
    config = configs.regression(3000, 64, sigma_latent=True)
    sites, specs = trace(lambda: config.model(mininf),
                         {"theta": (abi.NORMAL, torch.randn(64)), "sigma": (abi.GAMMA, torch.tensor(1.3))}, config.data)
    assert [s.name for s in sites] == ["theta", "sigma", "y"]          # X is under no_log_prob
    plan = Plan(sites, specs, 64, CPU, dense_mode="auto", dry_run=True)
    assert plan.D == 65 and len(plan.dense_sites) == 1 and not plan.sweep_groups
    site, mode = plan.dense_sites[0]
    # p = 64, S <= 64, N(0, 1) features: the fp16-operand tcgen05 kernel; "tf32" keeps TF32 operands
    assert mode == abi.DENSE_F16 and site.family == abi.NORMAL and (site.p, site.n_rows) == (64, 3000)
    assert Plan(sites, specs, 64, CPU, dense_mode="tf32", dry_run=True).dense_sites[0][1] == abi.DENSE_TF32
    # a design matrix outside fp16's range falls back to TF32 operands; asking for fp16 raises
    huge = dict(config.data, X=config.data["X"] * 1e6)
    far_sites, _ = trace(lambda: config.model(mininf),
                         {"theta": (abi.NORMAL, torch.randn(64)), "sigma": (abi.GAMMA, torch.tensor(1.3))}, huge)
    assert Plan(far_sites, specs, 64, CPU, dry_run=True).dense_sites[0][1] == abi.DENSE_TF32
    with pytest.raises(NotImplementedError, match="fp16"):
        Plan(far_sites, specs, 64, CPU, dense_mode="f16", dry_run=True)
    assert site.theta_lat == 0 and site.icpt_lat == -1 and site.scale.a_lat == 64 and site.weight == 1.0
    assert plan.small_global[1] == 2 and plan.small_observed is None
    # the closed-form switch selects the Gram-statistics mode of the same site
    closed = Plan(sites, specs, 64, CPU, dense_mode="auto", dry_run=True, closed_form=True)
    assert closed.dense_sites[0][1] == abi.DENSE_TF32_CLOSED_FORM
    # more than 128 particles: only the closed form has a tensor-core kernel
    assert Plan(sites, specs, 200, CPU, dry_run=True).dense_sites[0][1] == abi.DENSE_FP32
    assert Plan(sites, specs, 200, CPU, dry_run=True, closed_form=True).dense_sites[0][1] == abi.DENSE_TF32_CLOSED_FORM


def test_shapes_outside_the_tensor_core_kernel_use_fp32_or_raise():
    # rows of 22 floats are not 16-byte multiples: no TMA, hence no tcgen05 kernel
    config = configs.regression(100, 22)
    sites, specs = trace(lambda: config.model(mininf), {"theta": (abi.NORMAL, torch.randn(22))}, config.data)
    assert Plan(sites, specs, 8, CPU, dry_run=True).dense_sites[0][1] == abi.DENSE_FP32
    with pytest.raises(NotImplementedError, match="p == 64"):
        Plan(sites, specs, 8, CPU, dense_mode="tf32", dry_run=True)
    # more than 128 particles
    config = configs.regression(100, 128)
    sites, specs = trace(lambda: config.model(mininf), {"theta": (abi.NORMAL, torch.randn(128))}, config.data)
    assert Plan(sites, specs, 129, CPU, dry_run=True).dense_sites[0][1] == abi.DENSE_FP32
    # any multiple of four features runs on the wide kernel (TMA zero-fills the last chunk)
    config = configs.regression(100, 24)
    sites, specs = trace(lambda: config.model(mininf), {"theta": (abi.NORMAL, torch.randn(24))}, config.data)
    plan = Plan(sites, specs, 40, CPU, dry_run=True)
    assert plan.dense_sites[0][1] == abi.DENSE_TF32


def test_minibatch_weight_and_logits_family():
    config = configs.logistic(100_000, 500, p=64)
    sites, specs = trace(lambda: config.model(mininf), {"theta": (abi.NORMAL, torch.randn(64))}, config.data)
    site, _ = Plan(sites, specs, 16, CPU, dry_run=True).dense_sites[0]
    assert site.family == abi.BERNOULLI_LOGITS and site.weight == 200.0


def test_masked_sites_are_fused_into_one_sweep():
    config = configs.missing(5000)
    latents = {k: (abi.NORMAL, torch.randn(())) for k in "abcd"}
    latents["sigma"] = (abi.GAMMA, torch.tensor(0.7))
    sites, specs = trace(lambda: config.model(mininf), latents, config.data)
    plan = Plan(sites, specs, 64, CPU, dry_run=True)
    assert [len(group) for group in plan.sweep_groups] == [2] and plan.small_global[1] == 5
    poisson, normal = plan.sweep_groups[0]
    assert poisson.family == abi.POISSON and poisson.mask and poisson.numel == 5000
    assert (poisson.param[0].a_lat, poisson.param[0].b_lat, poisson.param[0].transform) == (0, 1, abi.T_EXP)
    assert (normal.param[0].a_lat, normal.param[0].b_lat, normal.param[0].transform) == (2, 3, abi.T_ID)
    assert normal.param[1].a_lat == 4 and not normal.param[1].x
    # the two sites share the covariate buffer's values
    assert poisson.param[0].x and normal.param[0].x


def test_short_sites_go_to_the_small_kernel():
    config = configs.coin()
    sites, specs = trace(lambda: config.model(mininf), {"theta": (abi.BETA, torch.tensor(0.6))}, config.data)
    plan = Plan(sites, specs, 1, CPU, dry_run=True)
    assert plan.small_observed[1] == 1 and plan.small_global[1] == 1 and not plan.dense_sites


def test_link_algebra():
    a = LinkTensor.wrap(torch.tensor(0.5), Affine(a_lat=LatentRef("a", 0)))
    b = LinkTensor.wrap(torch.tensor(2.0), Affine(a_lat=LatentRef("b", 0)))
    theta = LinkTensor.wrap(torch.randn(4), Affine(a_lat=LatentRef("theta")))
    x = torch.randn(6)
    X = torch.randn(6, 4)
    expr = (0.25 + a + b * x)._expr
    assert (expr.a_const, expr.a_lat.name, expr.b_lat.name, expr.transform) == (0.25, "a", "b", "id")
    assert torch.equal(expr.x, x)
    assert (a + b * x).exp()._expr.transform == "exp" and torch.exp(a + x * b)._expr.transform == "exp"
    dense = (X @ theta + a)._expr
    assert isinstance(dense, Dense) and dense.theta == "theta" and dense.icpt_lat.name == "a"
    assert (theta[2] + 1.0)._expr.a_lat == LatentRef("theta", 2)
    assert (1 * theta)._expr.is_pure_latent
    # real values ride along, so torch.distributions validates as usual
    torch.testing.assert_close((a + b * x).unwrap(), 0.5 + 2.0 * x)
    # outside the closed set: opaque
    for opaque in (torch.sin(a), a * b, x - a, theta @ X.T @ X[:, 0], 2.0 / a, (a + b * x) * 2.0, (X @ theta) * 2.0):
        assert opaque._expr is None


def test_unsupported_models_raise_instead_of_falling_back():
    x = torch.randn(50)
    y = torch.randn(50)

    def sine():
        a = mininf.sample("a", Normal(0, 1))
        mininf.sample("y", Normal(torch.sin(a * x), 1.0))

    sites, specs = trace(sine, {"a": (abi.NORMAL, torch.tensor(0.1))}, {"y": y})
    with pytest.raises(NotImplementedError, match="not a supported link"):
        Plan(sites, specs, 2, CPU, dry_run=True)

    def student():
        a = mininf.sample("a", Normal(0, 1))
        mininf.sample("y", torch.distributions.StudentT(3.0, a + 0 * x, 1.0))

    sites, specs = trace(student, {"a": (abi.NORMAL, torch.tensor(0.1))}, {"y": y})
    with pytest.raises(NotImplementedError, match="StudentT has no CUDA log-density"):
        Plan(sites, specs, 2, CPU, dry_run=True)


def test_tracer_keeps_reference_errors():
    def model():
        mininf.sample("a", Normal(0, 1))
        mininf.sample("y", Normal(0, 1), 5)

    a = LinkTensor.wrap(torch.tensor(0.1), Affine(a_lat=LatentRef("a", 0)))
    with SiteTableTracer(), pytest.raises(ValueError, match="'y' is missing"):
        mininf.condition(model, a=a)()
    with SiteTableTracer(), pytest.raises(ValueError, match="Expected shape"):
        mininf.condition(model, a=a, y=torch.randn(4))()
    with SiteTableTracer(), pytest.raises(ValueError, match="is not in the support"):
        mininf.condition(lambda: mininf.sample("c", Poisson(3.0), 3), c=torch.tensor([1.0, 2.5, 0.0]))()
    masked = torch.masked.as_masked_tensor(torch.randn(5), torch.rand(5) > 0.5)

    def batched():
        with mininf.batch(10):
            mininf.sample("y", Normal(0, 1), 10)

    with SiteTableTracer(), pytest.raises(ValueError, match="not supported for masked data"):
        mininf.condition(batched, y=masked)()


# -------------------------------------------------------------------------------------------------
# rebinding a plan to the next minibatch (Plan.bind_sources / Plan.rebind)
# -------------------------------------------------------------------------------------------------
def _leaves(data):
    from mininf_b200.nn import _leaves as leaves
    return leaves(list(data.values()))


def test_plan_rebinds_dense_site_to_a_new_batch():
    first = configs.logistic(100_000, 512, p=64)
    second = configs.logistic(100_000, 512, p=64, batch_id=1)
    sites, specs = trace(lambda: first.model(mininf), {"theta": (abi.NORMAL, torch.randn(64))}, first.data)
    plan = Plan(sites, specs, 16, CPU, dry_run=True)
    plan.bind_sources(_leaves(first.data))
    assert plan.rebindable
    site, mode = plan.dense_sites[0]
    assert mode in (abi.DENSE_F16, abi.DENSE_TF32) and site.X == first.data["X"].data_ptr() and site.y == first.data["y"].data_ptr()
    assert plan.rebind(_leaves(second.data))
    assert site.X == second.data["X"].data_ptr() and site.y == second.data["y"].data_ptr()
    # a batch whose rows are not 16-byte aligned cannot feed the TMA path: the caller re-lowers
    shifted = {"X": torch.randn(512 * 64 + 1)[1:].view(512, 64), "y": second.data["y"]}
    assert shifted["X"].data_ptr() % 16 != 0
    assert not plan.rebind(_leaves(shifted))
    assert site.X == second.data["X"].data_ptr()            # untouched by the refused rebind
    # different sizes never rebind
    assert not plan.rebind(_leaves(configs.logistic(100_000, 256, p=64).data))


def test_plan_rebinds_masked_sweep_groups_and_covariates():
    n = 5000

    def model():
        a = mininf.sample("a", Normal(0, 1))
        b = mininf.sample("b", Normal(0, 1))
        with mininf.no_log_prob():
            x = mininf.sample("x", Normal(0, 1), n)
        mininf.sample("counts", Poisson((a + b * x).exp()))

    def batch(seed):
        raw = configs.missing(n, seed0=seed).extra["raw"]
        return {"x": raw["x"], "counts": torch.masked.as_masked_tensor(raw["counts"], raw["m_counts"])}

    first, second = batch(5000), batch(7000)
    latents = {k: (abi.NORMAL, torch.randn(())) for k in "ab"}
    sites, specs = trace(model, latents, first)
    plan = Plan(sites, specs, 8, CPU, dry_run=True)
    plan.bind_sources(_leaves(first))
    assert plan.rebindable and len(plan.sweep_groups) == 1
    site = plan.sweep_groups[0][0]
    assert (site.value, site.mask, site.param[0].x) == (first["counts"].get_data().data_ptr(),
                                                       first["counts"].get_mask().data_ptr(), first["x"].data_ptr())
    assert plan.rebind(_leaves(second))
    assert (site.value, site.mask, site.param[0].x) == (second["counts"].get_data().data_ptr(),
                                                       second["counts"].get_mask().data_ptr(), second["x"].data_ptr())
    # a covariate captured by the model's closure is not a conditioned tensor: such plans re-lower
    config = configs.missing(n)
    latents = {k: (abi.NORMAL, torch.randn(())) for k in "abcd"}
    latents["sigma"] = (abi.GAMMA, torch.tensor(0.7))
    sites, specs = trace(lambda: config.model(mininf), latents, config.data)
    plan = Plan(sites, specs, 8, CPU, dry_run=True)
    plan.bind_sources(_leaves(config.data))
    assert not plan.rebindable


def test_plans_with_converted_copies_are_not_rebindable():
    config = configs.logistic(100_000, 300, p=64)
    data = {"X": config.data["X"], "y": config.data["y"].double()}       # float64 -> the plan owns a float32 copy
    sites, specs = trace(lambda: config.model(mininf), {"theta": (abi.NORMAL, torch.randn(64))}, data)
    plan = Plan(sites, specs, 4, CPU, dry_run=True)
    plan.bind_sources(_leaves(data))
    assert not plan.rebindable and not plan.rebind(_leaves(data | {"y": data["y"].clone()}))
    assert plan.rebind(_leaves(data)) is False


def test_tuple_unpacking_a_latent_keeps_the_link_to_it():
    """`a, b = sample(...)` goes through Tensor.__iter__ -> unbind: the pieces must stay scalar
    references to the latent (a stripped tensor would freeze the trace-time draw into the plan and
    the gradient with respect to 'ab' would silently be zero)."""
    x, y = torch.randn(4000), torch.randn(4000)

    def model():
        a, b = mininf.sample("ab", Normal(0, 1), [2])
        mininf.sample("y", Normal(a + b * x, 1.0))

    sites, specs = trace(model, {"ab": (abi.NORMAL, torch.randn(2))}, {"y": y})
    plan = Plan(sites, specs, 4, CPU, dry_run=True)
    (site,) = plan.sweep_groups[0]
    assert (site.param[0].a_lat, site.param[0].b_lat) == (0, 1)


@pytest.mark.parametrize("how", ["split", "inplace", "item", "max"])
def test_untracked_uses_of_a_latent_raise_instead_of_freezing_the_draw(how):
    x, y = torch.randn(300), torch.randn(300)

    def model():
        ab = mininf.sample("ab", Normal(0, 1), [2])
        if how == "split":
            a, b = ab.split(1)
            loc = a + b * x
        elif how == "inplace":
            loc = ab[0] + ab[1] * x
            loc.add_(1.0)
        elif how == "item":
            loc = ab[0].item() + x
        else:
            loc = ab.max(dim=0).values + x
        mininf.sample("y", Normal(loc, 1.0))

    with pytest.raises(NotImplementedError):
        sites, specs = trace(model, {"ab": (abi.NORMAL, torch.randn(2))}, {"y": y})
        Plan(sites, specs, 4, CPU, dry_run=True)


def _evaluate(expr, latents, shape):
    """The link form ``T(a_const + a_lat + (b_const + b_lat) * x)`` evaluated on the host."""
    def value(ref):
        if ref is None:
            return torch.zeros(())
        return latents[ref.name].reshape(-1)[ref.index] if ref.is_scalar else latents[ref.name]
    out = expr.a_const + value(expr.a_lat) + torch.zeros(shape)
    if expr.x is not None:
        out = out + (expr.b_const + value(expr.b_lat)) * expr.x
    return {"id": lambda t: t, "exp": torch.exp, "sigmoid": torch.sigmoid}[expr.transform](out)


def test_widened_link_algebra_evaluates_to_the_traced_values():
    # negation, scaling, division and differences reduce to the one affine form the kernels read
    # (SURVEY 8 a10); every expression must reproduce the real values that ride along
    torch.manual_seed(3)
    latents = {"a": torch.tensor(0.5), "b": torch.tensor(-2.0), "theta": torch.randn(6)}
    a = LinkTensor.wrap(latents["a"], Affine(a_lat=LatentRef("a", 0)))
    b = LinkTensor.wrap(latents["b"], Affine(a_lat=LatentRef("b", 0)))
    theta = LinkTensor.wrap(latents["theta"], Affine(a_lat=LatentRef("theta")))
    x, w = torch.randn(6), torch.rand(6) + 0.5
    built = {
        "-a": -a, "2 * theta": 2.0 * theta, "theta * -0.5": theta * -0.5, "a - b * x": a - b * x,
        "1 - x * b": 1.0 - x * b, "a / 2": a / 2, "theta / w": theta / w, "(1 + a) * x": (1 + a) * x,
        "(b * x) * w": (b * x) * w, "(b * x) / w": (b * x) / w, "a - b": a - b, "3 - a": 3 - a,
        "rsub": torch.rsub(a, 2.0), "a * x - 1": a * x - 1, "-(b * x)": -(b * x), "theta - x": theta - x,
        "a - x": a - x, "0.5 * x + a": 0.5 * x + a, "exp(a - b * x)": torch.exp(a - b * x),
        "neg of a data product": torch.neg(b * x) + 2.0, "x / 4 + theta": x / 4 + theta,
        "sigmoid": torch.sigmoid(a + b * x),
    }
    for name, tensor in built.items():
        expr = tensor._expr
        assert isinstance(expr, Affine), name
        torch.testing.assert_close(_evaluate(expr, latents, tensor.shape), tensor.unwrap(), msg=name)
    # a coefficient on the intercept latent next to a slope term, a latent denominator: opaque
    for name, tensor in {"x - a": x - a, "2 * (a + b * x)": 2 * (a + b * x), "x / a": x / a,
                         "floor division": torch.div(a, 2.0, rounding_mode="floor"),
                         "sigmoid of exp": torch.sigmoid(torch.exp(a))}.items():
        assert tensor._expr is None, name


def test_widened_links_lower_to_the_site_table():
    x = torch.randn(5000)
    y = torch.randn(5000)

    def model():
        a = mininf.sample("a", Normal(0, 1))
        b = mininf.sample("b", Normal(0, 1))
        mininf.sample("y", Normal(a - b * x / 2, 1.0))

    sites, specs = trace(model, {"a": (abi.NORMAL, torch.tensor(0.1)), "b": (abi.NORMAL, torch.tensor(0.2))}, {"y": y})
    plan = Plan(sites, specs, 8, CPU, dry_run=True)
    (site,) = plan.sweep_groups[0]
    link = site.param[0]
    assert (link.a_lat, link.b_lat, link.a_const, link.b_const, link.transform) == (0, 1, 0.0, 0.0, abi.T_ID)
    covariate = next(t for t in plan.keepalive if t.data_ptr() == link.x)
    torch.testing.assert_close(covariate, -x / 2)
    # a covariate built at trace time is not one of the conditioned tensors: a new batch retraces
    plan.bind_sources([y])
    assert not plan.rebindable


def test_bernoulli_probs_of_a_sigmoid_is_lowered_as_logits():
    config = configs.logistic(2000, 2000, p=64)
    X, y = config.data["X"], config.data["y"]

    def dense():
        theta = mininf.sample("theta", Normal(0, 1), [64])
        mininf.sample("y", Bernoulli(probs=torch.sigmoid(X @ theta)))

    sites, specs = trace(dense, {"theta": (abi.NORMAL, 0.05 * torch.randn(64))}, {"y": y})
    site, _ = Plan(sites, specs, 16, CPU, dry_run=True).dense_sites[0]
    assert site.family == abi.BERNOULLI_LOGITS and site.n_rows == 2000

    def scalar():
        a = mininf.sample("a", Normal(0, 1))
        mininf.sample("y", Bernoulli(probs=torch.sigmoid(a + 0.5 * X[:, 0])))

    sites, specs = trace(scalar, {"a": (abi.NORMAL, torch.tensor(0.1))}, {"y": y})
    table = Plan(sites, specs, 16, CPU, dry_run=True)._small_observed_host
    assert len(table) == 1 and table[0].family == abi.BERNOULLI_LOGITS
    assert (table[0].param[0].a_lat, table[0].param[0].b_const, table[0].param[0].transform) == (0, 1.0, abi.T_ID)   # 0.5 * X[:, 0] is data

    # a sigmoid anywhere else is not a link the kernels evaluate
    def elsewhere():
        a = mininf.sample("a", Normal(0, 1))
        mininf.sample("y", Normal(torch.sigmoid(a + 0.5 * X[:, 0]), 1.0))

    sites, specs = trace(elsewhere, {"a": (abi.NORMAL, torch.tensor(0.1))}, {"y": y.float()})
    with pytest.raises(NotImplementedError, match="sigmoid"):
        Plan(sites, specs, 16, CPU, dry_run=True)


def test_one_element_conditioned_tensors_make_the_plan_retrace():
    # a scalar tensor handed to `condition` is folded into a link constant; a plan that kept it
    # across batches would silently score the old value (the reference re-reads it every step)
    x = torch.randn(3000)
    y = torch.randn(3000)
    noise_scale = torch.tensor(0.5)

    def model():
        a = mininf.sample("a", Normal(0, 1))
        scale = mininf.value("noise_scale")
        mininf.sample("y", Normal(a + x, scale))

    sites, specs = trace(model, {"a": (abi.NORMAL, torch.tensor(0.1))}, {"y": y, "noise_scale": noise_scale})
    plan = Plan(sites, specs, 4, CPU, dry_run=True)
    (site,) = plan.sweep_groups[0]
    assert site.param[1].a_const == 0.5 and not site.param[1].x        # folded
    plan.bind_sources([y, noise_scale])
    assert not plan.rebindable and not plan.rebind([torch.randn(3000), torch.tensor(0.7)])
    plan.bind_sources([y])                                                 # without the scalar leaf: pointers only
    assert not plan.rebindable                                             # x is captured, not conditioned
    sites, specs = trace(lambda: mininf.sample("y", Normal(mininf.sample("a", Normal(0, 1)), 1.0), [3000]),
                         {"a": (abi.NORMAL, torch.tensor(0.1))}, {"y": y})
    plan = Plan(sites, specs, 4, CPU, dry_run=True)
    plan.bind_sources([y])
    assert plan.rebindable


def test_latent_vector_times_scalar_latent_is_a_one_feature_row_dot():
    from mininf_b200.engine.plan import assign_offsets, row_latent_names
    from mininf_b200.engine.trace import Linear, RowDot
    n = 40
    x, y = torch.randn(n), torch.poisson(torch.ones(n))

    def model(slope_of):
        def run():
            z = mininf.sample("z", Normal(0, 1), n)
            mininf.sample("x", Normal(z, 0.5))
            theta = mininf.sample("theta", Normal(0, 1), 3)
            slope = mininf.sample("slope", Normal(0, 1))
            mininf.sample("y", Poisson((0.5 + slope_of(theta, slope) * z).exp()))
        return run

    latents = {"z": (abi.NORMAL, torch.randn(n)), "theta": (abi.NORMAL, torch.randn(3)),
               "slope": (abi.NORMAL, torch.tensor(0.4))}
    sites, _ = trace(model(lambda theta, slope: slope), latents, {"x": x, "y": y})
    expr = sites[-1].distribution.rate._expr
    assert isinstance(expr, RowDot) and (expr.Z, expr.beta, expr.icpt_const, expr.transform) == ("z", "slope", 0.5, "exp")
    assert row_latent_names(sites) == {"z"}
    specs = assign_offsets([(k, f, v.shape) for k, (f, v) in latents.items()], row_latent_names(sites))
    assert [s.row_latent for s in specs] == [True, False, False]
    plan = Plan(sites, specs, 4, CPU, dry_run=True)
    assert plan.row_groups["z"].p == 1 and plan.row_groups["z"].beta_lat == specs[2].offset
    # an element of a longer vector is not a coefficient vector of p = 1 elements
    sites, _ = trace(model(lambda theta, slope: theta[0]), latents, {"x": x, "y": y})
    specs = assign_offsets([(k, f, v.shape) for k, (f, v) in latents.items()], row_latent_names(sites))
    with pytest.raises(NotImplementedError, match="do not match"):
        Plan(sites, specs, 4, CPU, dry_run=True)
    # without the row-latent treatment (z packed) the site is refused, not mis-scored
    with pytest.raises(NotImplementedError):
        Plan(*trace(model(lambda theta, slope: slope), latents, {"x": x, "y": y}), 4, CPU, dry_run=True)


def test_random_link_expressions_are_either_opaque_or_exact():
    """Property test of the link algebra: random expression trees over scalar latents, an
    element-wise latent vector, data tensors and constants. Whatever the tracer does not mark as
    opaque must evaluate - from the recorded link form alone - to the values torch computed. A wrong
    rule would lower a model to tables that silently score something else."""
    import random
    from mininf_b200.engine.trace import Linear, RowDot
    rng = random.Random(7)
    torch.manual_seed(7)
    n = 6
    latents = {"a": torch.tensor(0.7), "b": torch.tensor(-1.3), "theta": torch.randn(n)}
    x, w = torch.randn(n), torch.rand(n) + 0.5

    def leaf():
        kind = rng.choice(["a", "b", "theta", "x", "w", "const", "const"])
        if kind in ("a", "b"):
            return LinkTensor.wrap(latents[kind], Affine(a_lat=LatentRef(kind, 0)))
        if kind == "theta":
            return LinkTensor.wrap(latents["theta"], Affine(a_lat=LatentRef("theta")))
        if kind == "const":
            return rng.choice([2.0, -0.5, 3, 0.25])
        return {"x": x, "w": w}[kind]

    def grow(depth):
        if depth == 0 or rng.random() < 0.25:
            return leaf()
        op = rng.choice(["add", "sub", "mul", "div", "neg", "radd", "rsub", "shape"])
        left, right = grow(depth - 1), grow(depth - 1)
        if op == "shape":                              # layout-only operations keep (or drop) the link, never bend it
            if not isinstance(left, torch.Tensor):
                return left
            how = rng.choice(["reshape", "unsqueeze", "expand", "clone", "contiguous", "float", "index", "detach"])
            if how == "reshape":
                return left.reshape(-1) if left.ndim else left.reshape(())
            if how == "unsqueeze":
                return left.unsqueeze(0).squeeze(0)
            if how == "expand":
                return left.expand(n) if left.ndim == 0 or left.shape == (n,) else left
            if how == "index":
                return left[rng.randrange(n)] if left.ndim == 1 else left
            return getattr(left, how)()
        if op == "neg":
            return -left if isinstance(left, torch.Tensor) else -float(left)
        if op in ("add", "radd"):
            return left + right
        if op in ("sub", "rsub"):
            return left - right
        if op == "mul":
            return left * right
        if isinstance(right, torch.Tensor) and not isinstance(right, LinkTensor):
            return left / right                       # data denominators are bounded away from zero only for w
        return left / (right if not isinstance(right, torch.Tensor) and abs(right) > 1e-3 else 2.0)

    exact = opaque = linear = 0
    for _ in range(600):
        tree = grow(4)
        if rng.random() < 0.3 and isinstance(tree, LinkTensor):
            tree = torch.exp(tree.clamp(-3, 3)) if rng.random() < 0.2 else torch.exp(0.1 * tree)
        if not isinstance(tree, LinkTensor):
            continue                                   # no latent involved
        values = tree.unwrap()
        if not torch.isfinite(values).all():
            continue
        if tree._expr is None:
            opaque += 1
            continue
        expr = tree._expr
        if isinstance(expr, RowDot):                  # latent vector times scalar latent: Z @ beta with p = 1
            got = latents[expr.Z] * latents[expr.beta] + expr.icpt_const
            if expr.icpt_lat is not None:
                got = got + latents[expr.icpt_lat.name].reshape(-1)[expr.icpt_lat.index or 0]
            got = got.exp() if expr.transform == "exp" else got
        elif isinstance(expr, Linear):                # several scalar latents, each with its own covariate
            scalar = lambda ref: latents[ref.name].reshape(-1)[ref.index or 0]     # noqa: E731
            got = expr.icpt_const + sum(scalar(ref) * cov for ref, cov in expr.terms) + torch.zeros(values.shape)
            if expr.icpt_lat is not None:
                got = got + scalar(expr.icpt_lat)
            got = got.exp() if expr.transform == "exp" else got
            linear += 1
        else:
            assert isinstance(expr, Affine)
            got = _evaluate(expr, latents, values.shape)
        torch.testing.assert_close(got, values, rtol=1e-4, atol=1e-4)
        exact += 1
    assert exact > 100 and opaque > 50                 # both outcomes are exercised (Linear: the test below)


def test_several_covariates_lower_to_a_dense_site_over_a_design_matrix_built_at_trace_time():
    from mininf_b200.engine.trace import Linear
    n = 3000
    torch.manual_seed(5)
    x1, x2, x3, y = torch.randn(n), torch.randn(n), torch.rand(n), torch.randn(n)

    def model():
        a = mininf.sample("a", Normal(0, 1))
        b1 = mininf.sample("b1", Normal(0, 1))
        b2 = mininf.sample("b2", Normal(0, 1))
        beta = mininf.sample("beta", Normal(0, 1), [2])
        sigma = mininf.sample("sigma", Gamma(2, 2))
        eta = 0.5 + a + b1 * x1 - b2 * x2 / 2 + beta[0] * x3 + b1 * x3        # b1 rides on two covariates
        mininf.sample("y", Normal(eta, sigma))

    latents = {"a": (abi.NORMAL, torch.tensor(0.1)), "b1": (abi.NORMAL, torch.tensor(0.2)),
               "b2": (abi.NORMAL, torch.tensor(0.3)), "beta": (abi.NORMAL, torch.tensor([0.4, 0.5])),
               "sigma": (abi.GAMMA, torch.tensor(1.1))}
    sites, specs = trace(model, latents, {"y": y})
    expr = sites[-1].distribution.loc._expr
    assert isinstance(expr, Linear) and len(expr.terms) == 4 and expr.icpt_const == 0.5 and expr.icpt_lat.name == "a"
    plan = Plan(sites, specs, 8, CPU, dry_run=True)
    (site, mode), = plan.dense_sites
    assert (site.family, site.p, site.n_rows, site.theta_lat, site.icpt_lat) == (abi.NORMAL, 3, n, 1, 0)   # b1, b2, beta[0]
    assert mode == abi.DENSE_FP32 and site.scale.a_lat == 5
    X = next(t for t in plan.keepalive if t.data_ptr() == site.X)
    torch.testing.assert_close(X, torch.stack([x1 + x3, -x2 / 2, x3], dim=1))
    plan.bind_sources([y])
    assert not plan.rebindable                              # the design matrix is derived, a new batch retraces

    # slopes that are not neighbours in the packed latents are refused with a hint, not mis-read
    def gapped():
        b1 = mininf.sample("b1", Normal(0, 1))
        mininf.sample("a", Normal(0, 1))
        beta = mininf.sample("beta", Normal(0, 1), [2])
        mininf.sample("y", Normal(b1 * x1 + beta[1] * x2, 1.0))

    sites, specs = trace(gapped, latents, {"y": y})
    with pytest.raises(NotImplementedError, match="next to each other"):
        Plan(sites, specs, 8, CPU, dry_run=True)
    # a data-only term or two latent intercepts have no place in the form: opaque
    a = LinkTensor.wrap(torch.tensor(0.5), Affine(a_lat=LatentRef("a", 0)))
    b = LinkTensor.wrap(torch.tensor(2.0), Affine(a_lat=LatentRef("b", 0)))
    assert (a * x1 + b * x2 + x3)._expr is None and (a + a * x1 + b + b * x2)._expr is None
    assert isinstance(torch.exp(a * x1 + b * x2)._expr, Linear) and torch.exp(a * x1 + b * x2)._expr.transform == "exp"


def test_slopes_of_a_several_covariate_link_are_packed_next_to_each_other():
    from mininf_b200.engine.plan import _packing_order, assign_offsets, slope_groups
    n = 3000
    x1, x2, y = torch.randn(n), torch.randn(n), torch.randn(n)

    def model():
        b1 = mininf.sample("b1", Normal(0, 1))
        a = mininf.sample("a", Normal(0, 1))
        sigma = mininf.sample("sigma", Gamma(2, 2))
        b2 = mininf.sample("b2", Normal(0, 1))
        mininf.sample("y", Normal(a + b1 * x1 + b2 * x2, sigma))

    latents = {"b1": (abi.NORMAL, torch.tensor(0.2)), "a": (abi.NORMAL, torch.tensor(0.1)),
               "sigma": (abi.GAMMA, torch.tensor(1.1)), "b2": (abi.NORMAL, torch.tensor(0.3))}
    sites, in_order = trace(model, latents, {"y": y})
    with pytest.raises(NotImplementedError, match="next to each other"):       # the caller's order: b1 | a sigma | b2
        Plan(sites, in_order, 4, CPU, dry_run=True)
    assert slope_groups(sites) == [["b1", "b2"]]
    specs = assign_offsets([(k, f, v.shape) for k, (f, v) in latents.items()], (), slope_groups(sites))
    assert [(s.name, s.offset) for s in specs] == [("b1", 0), ("b2", 1), ("a", 2), ("sigma", 3)]
    (site, _), = Plan(sites, specs, 4, CPU, dry_run=True).dense_sites
    assert (site.p, site.theta_lat, site.icpt_lat, site.scale.a_lat) == (2, 0, 2, 3)
    # without groups nothing moves; chained groups end up in one run
    assert [s.name for s in assign_offsets([(k, f, v.shape) for k, (f, v) in latents.items()])] == list(latents)
    assert _packing_order(list("abcdef"), [["b", "e"], ["e", "c"], ["f", "a"]]) == ["f", "a", "b", "e", "c", "d"]


def test_indexing_a_latent_vector_with_group_labels_is_a_dense_link_over_an_indicator_matrix():
    alpha = LinkTensor.wrap(torch.tensor([0.1, -0.2, 0.3]), Affine(a_lat=LatentRef("alpha")))
    mu = LinkTensor.wrap(torch.tensor(1.5), Affine(a_lat=LatentRef("mu", 0)))
    group = torch.tensor([2, 0, 0, 1, 2, -1])
    picked = alpha[group]
    assert isinstance(picked._expr, Dense) and picked._expr.theta == "alpha"
    torch.testing.assert_close(picked._expr.X @ alpha.unwrap(), picked.unwrap())       # one_hot(group) @ alpha
    shifted = (mu + picked)._expr
    assert isinstance(shifted, Dense) and shifted.icpt_lat == LatentRef("mu", 0)
    # one element stays a scalar reference; float or boolean indices and 2-D sources are not this form
    assert alpha[torch.tensor(1)]._expr.a_lat == LatentRef("alpha", 1)
    assert alpha[torch.tensor([1])]._expr is None or alpha[torch.tensor([1])]._expr.a_lat == LatentRef("alpha", 1)
    assert alpha[torch.tensor([True, False, True])]._expr is None
    assert (alpha[group] * 2.0)._expr is None
