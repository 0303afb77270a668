"""What the lowered site table MEANS, checked without a GPU: a float64 host interpreter of the
C-ABI structures (``mnf_site_t`` / ``mnf_dense_site_t`` with their ``mnf_link_t`` parameters,
include/mininf_b200.h) evaluates the log-joint the kernels are asked to compute and is compared
with the oracle's restatement of the reference (oracle/handlers.py following mininf/core.py:211-273)
on the same latent values. This pins the host half of the drop-in boundary (trace -> link algebra
-> tables); the device half (tables -> numbers) is the `-m gpu` parity suite."""
import ctypes

import pytest
import torch
from torch.distributions import Bernoulli, Beta, Gamma, Normal, Poisson

import mininf_b200 as mininf
from mininf_b200.engine import abi
from mininf_b200.engine.plan import Plan, assign_offsets, row_latent_names, slope_groups
from mininf_b200.engine.trace import Affine, LatentRef, LinkTensor, SiteTableTracer
from oracle import configs, handlers

CPU = torch.device("cpu")
FAMILY = {Normal: abi.NORMAL, Gamma: abi.GAMMA, Beta: abi.BETA}


def lower(model, data, latents, n_particles=4):
    """Trace ``model`` with the latent values wrapped as links and lower it (dry run)."""
    draws, specs, offset = {}, [], 0
    for name, (family, value) in latents.items():
        numel = max(value.numel(), 1)
        ref = LatentRef(name, 0) if numel == 1 else LatentRef(name)
        draws[name] = LinkTensor.wrap(value.float(), Affine(a_lat=ref))
    with SiteTableTracer() as tracer:
        mininf.condition(mininf.condition(lambda: model(mininf), **data), **draws)()
    # as EvidenceLowerBoundLoss._build_plan: large Normal latents and the Z of a `Z @ beta` link are row latents
    entries = [(name, family, value.shape) for name, (family, value) in latents.items()]
    specs = assign_offsets(entries, row_latent_names(tracer.sites), slope_groups(tracer.sites))
    return Plan(tracer.sites, specs, n_particles, CPU, dry_run=True), specs


class Interpreter:
    """Evaluates the tables for one packed latent vector z[D] in float64."""

    def __init__(self, plan, z, rows=None):
        self.plan, self.z, self.rows = plan, z.double(), rows or {}

    def buffer(self, pointer, count, stride=1):
        for tensor in self.plan.keepalive:
            size = tensor.numel() * tensor.element_size()
            if tensor.data_ptr() <= pointer < tensor.data_ptr() + max(size, 1):
                offset = (pointer - tensor.data_ptr()) // tensor.element_size()
                flat = tensor.reshape(-1)
                index = offset + stride * torch.arange(count)
                return flat[index]
        raise AssertionError("a table points outside every tensor the plan keeps alive")

    def link(self, L, count):
        i = torch.arange(count)
        a = torch.full((count,), float(L.a_const), dtype=torch.float64)
        if L.a_lat >= 0:
            a = a + self.z[L.a_lat + L.a_stride * i]
        b = torch.full((count,), float(L.b_const), dtype=torch.float64)
        if L.b_lat >= 0:
            b = b + self.z[L.b_lat + L.b_stride * i]
        x = self.buffer(L.x, count, L.x_stride).double() if L.x else torch.ones(count, dtype=torch.float64)
        eta = a + b * x                              # csrc/common.cuh: x = 1 without a covariate
        return eta.exp() if L.transform == abi.T_EXP else eta

    @staticmethod
    def log_density(family, value, p0, p1):
        dist = {abi.NORMAL: lambda: Normal(p0, p1), abi.GAMMA: lambda: Gamma(p0, p1), abi.BETA: lambda: Beta(p0, p1),
                abi.BERNOULLI_PROBS: lambda: Bernoulli(probs=p0), abi.BERNOULLI_LOGITS: lambda: Bernoulli(logits=p0),
                abi.POISSON: lambda: Poisson(p0)}[family]()
        return dist.log_prob(value)

    def site(self, site):
        count = site.numel
        if site.value_lat >= 0:
            value = self.z[site.value_lat + torch.arange(count)]
        else:
            value = self.buffer(site.value, count).double()
        lp = self.log_density(site.family, value, self.link(site.param[0], count), self.link(site.param[1], count))
        if site.mask:
            lp = torch.where(self.buffer(site.mask, count), lp, torch.zeros_like(lp))
        return site.scale * lp.sum()

    def dense(self, site):
        X = self.buffer(site.X, site.n_rows * site.ldx).double().reshape(site.n_rows, site.ldx)[:, :site.p]
        eta = X @ self.z[site.theta_lat:site.theta_lat + site.p] + float(site.icpt_const)
        if site.icpt_lat >= 0:
            eta = eta + self.z[site.icpt_lat]
        y = self.buffer(site.y, site.n_rows).double()
        scale = self.link(site.scale, site.n_rows)
        lp = self.log_density(site.family, y, eta.exp() if site.family == abi.POISSON else eta, scale)
        if site.mask:
            lp = torch.where(self.buffer(site.mask, site.n_rows), lp, torch.zeros_like(lp))
        return site.weight * lp.sum()

    def row_latent(self, desc, Z):
        """``mnf_rowlatent_t``: prior of the latent matrix, its observed noisy copy, the response."""
        n, p = desc.n_rows, desc.p
        Z = Z.double().reshape(n, p)
        scalar = lambda L: self.link(L, 1)[0]                                      # noqa: E731
        out = Normal(scalar(desc.prior_loc), scalar(desc.prior_scale)).log_prob(Z).sum()
        if desc.feat:
            feat = self.buffer(desc.feat, n * p).double().reshape(n, p)
            out = out + Normal(Z, scalar(desc.feat_scale)).log_prob(feat).sum()
        if desc.resp:
            eta = Z @ self.z[desc.beta_lat:desc.beta_lat + p] + float(desc.icpt_const)
            if desc.icpt_lat >= 0:
                eta = eta + self.z[desc.icpt_lat]
            if desc.resp_transform == abi.T_EXP:
                eta = eta.exp()
            resp = self.buffer(desc.resp, n).double()
            out = out + self.log_density(desc.resp_family, resp, eta, scalar(desc.resp_scale)).sum()
        return out

    def total(self):
        plan = self.plan
        out = sum(self.site(group[i]) for group in plan.sweep_groups for i in range(len(group)))
        out = out + sum(self.site(s) for s in plan._small_observed_host + plan._small_global_host)
        out = out + sum(self.row_latent(desc, self.rows[name]) for name, desc in plan.row_groups.items())
        return out + sum(self.dense(site) for site, _ in plan.dense_sites)


def oracle_log_joint(model, data, values):
    data64 = {}
    for name, value in data.items():
        if isinstance(value, torch.masked.MaskedTensor):
            data64[name] = torch.masked.as_masked_tensor(value.get_data().double(), value.get_mask())
        else:
            data64[name] = value.double()
    conditioned = handlers.condition(handlers.condition(lambda: model(handlers), **data64),
                                     **{k: v.double() for k, v in values.items()})
    return sum(handlers.evaluate(conditioned, {}, validate=False).values())


def packed(specs, values):
    return torch.cat([values[spec.name].reshape(-1).double() for spec in specs if not spec.row_latent])


def check(model64, model32, data, latents):
    plan, specs = lower(model32, data, latents)
    values = {name: value for name, (_, value) in latents.items()}
    rows = {spec.name: values[spec.name] for spec in specs if spec.row_latent}
    got = Interpreter(plan, packed(specs, values), rows).total()
    expected = oracle_log_joint(model64, data, values)
    torch.testing.assert_close(got, expected.double(), rtol=2e-6, atol=1e-6)   # data are stored as float32
    return plan


def test_widened_links_mean_what_the_reference_scores():
    # oracle/configs.py::affine_links: a - b*x/2, (c + a)*x/2, Bernoulli(probs=sigmoid(c - b*x))
    for n in (100, 5000):
        config = configs.affine_links(n)
        latents = {"a": (abi.NORMAL, torch.tensor(0.23)), "b": (abi.NORMAL, torch.tensor(-0.41)),
                   "sigma": (abi.GAMMA, torch.tensor(0.9))}
        plan = check(config.model, config.model, config.data, latents)
        assert (len(plan.sweep_groups) == 1) == (n >= 2048)


@pytest.mark.parametrize("n", [50, 4000])
def test_missing_observations_table(n):
    config = configs.missing(n)
    latents = {k: (abi.NORMAL, torch.tensor(v)) for k, v in zip("abcd", (0.3, 0.5, -0.2, 0.8))}
    latents["sigma"] = (abi.GAMMA, torch.tensor(0.7))
    check(config.model, config.model, config.data, latents)


def test_coin_and_regression_tables():
    config = configs.coin()
    check(config.model, config.model, config.data, {"theta": (abi.BETA, torch.tensor(0.6))})
    torch.manual_seed(0)
    config = configs.regression(300, 24, sigma_latent=True)
    check(config.model, config.model, config.data,
          {"theta": (abi.NORMAL, 0.1 * torch.randn(24)), "sigma": (abi.GAMMA, torch.tensor(1.3))})
    config = configs.logistic(100_000, 500, p=64, intercept=True)          # `batch` scaling: weight 200
    latents = {name: (FAMILY[cls], 0.1 * torch.randn(params["loc"].shape)) for name, (cls, params) in config.families.items()}
    check(config.model, config.model, config.data, latents)


def test_element_wise_latents_and_scaled_vectors():
    # a vector latent with a coefficient, against a data vector of the same length
    torch.manual_seed(1)
    w = torch.rand(7) + 0.5
    y = torch.randn(7)

    def model(m):
        t = m.sample("t", Normal(0, 1), [7])
        m.sample("y", Normal(2.0 * t / w - 1.0, 0.5))

    check(model, model, {"y": y}, {"t": (abi.NORMAL, torch.randn(7))})
    assert ctypes.sizeof(abi.Link) == 40          # the layout the interpreter reads (tests/test_abi.py holds it to gcc)


def test_row_latent_descriptor():
    # C4 (examples/regression-with-feature-uncertainty.md:28-38 widened to p features): the 300 x 32
    # latent matrix is a row latent - prior, noisy features and the Poisson response in one descriptor
    torch.manual_seed(2)
    config = configs.feature_uncertainty(300, 32)
    latents = {"population_scale": (abi.GAMMA, torch.tensor(1.1)),
               "z": (abi.NORMAL, config.data["x"] + 0.3 * torch.randn(300, 32)),
               "intercept": (abi.NORMAL, torch.tensor(0.4)), "slope": (abi.NORMAL, 0.1 * torch.randn(32))}
    plan = check(config.model, config.model, config.data, latents)
    (desc,) = plan.row_groups.values()
    assert (desc.n_rows, desc.p, desc.resp_family, desc.resp_transform) == (300, 32, abi.POISSON, abi.T_EXP)
    assert not plan.sweep_groups and not plan.dense_sites


def test_the_feature_uncertainty_example_as_written():
    # examples/regression-with-feature-uncertainty.md:28-38 literally (oracle/configs.py::feature_example):
    # ONE latent feature per row, n = 30, `intercept + z * slope`, the noise scale conditioned on as a
    # known value. The product of the latent vector and the scalar latent is `Z @ beta` with p = 1, and
    # z becomes a row latent although it is small.
    torch.manual_seed(13)
    n = 30
    config = configs.feature_example(n)
    latents = {"population_scale": (abi.GAMMA, torch.tensor(1.2)),
               "z": (abi.NORMAL, config.data["x"] + 0.1 * torch.randn(n)),
               "intercept": (abi.NORMAL, torch.tensor(0.25)), "slope": (abi.NORMAL, torch.tensor(0.6))}
    plan = check(config.model, config.model, config.data, latents)
    (desc,) = plan.row_groups.values()
    assert (desc.n_rows, desc.p, desc.resp_family, desc.resp_transform) == (n, 1, abi.POISSON, abi.T_EXP)
    assert desc.feat_scale.a_const == pytest.approx(0.3) and desc.beta_lat >= 0 and desc.icpt_lat >= 0
    # the Gamma(2, 2) density of the conditioned noise scale is a small observed site of its own
    assert len(plan._small_observed_host) == 1 and plan._small_observed_host[0].family == abi.GAMMA


def test_a_gallery_of_model_forms():
    """Shapes of models users of the reference write (eight schools, Gamma-Poisson with an exposure,
    heteroscedastic noise, vector Beta-Bernoulli, unpacked coefficient vectors, log offsets, centred
    covariates, negated latents): each lowers to tables that mean what the reference scores."""
    torch.manual_seed(0)
    J, n = 8, 3000
    sigma_known, effects = torch.rand(J) + 0.5, 3 * torch.randn(J)
    exposure = torch.rand(n) + 0.5
    counts = torch.poisson(2 * exposure)
    x, w, y = torch.randn(n), torch.rand(n) + 0.5, torch.randn(n)
    flips = torch.bernoulli(torch.full((12,), 0.3))

    def schools(m):
        mu = m.sample("mu", Normal(0, 5))
        tau = m.sample("tau", Gamma(2, 0.5))
        theta = m.sample("theta", Normal(mu, tau), [J])
        m.sample("y", Normal(theta, sigma_known))

    check(schools, schools, {"y": effects}, {"mu": (abi.NORMAL, torch.tensor(0.5)), "tau": (abi.GAMMA, torch.tensor(1.5)),
                                             "theta": (abi.NORMAL, torch.randn(J))})

    def gamma_poisson(m):
        rate = m.sample("rate", Gamma(2, 2))
        m.sample("y", Poisson(rate * exposure))

    check(gamma_poisson, gamma_poisson, {"y": counts}, {"rate": (abi.GAMMA, torch.tensor(1.7))})

    def heteroscedastic(m):
        a = m.sample("a", Normal(0, 1))
        b = m.sample("b", Normal(0, 1))
        log_scale = m.sample("log_scale", Normal(0, 1))
        s = m.sample("s", Gamma(2, 2))
        m.sample("y", Normal(a + b * x, s * w))
        m.sample("y2", Normal(a - x, log_scale.exp()))

    check(heteroscedastic, heteroscedastic, {"y": y, "y2": y + 1},
          {"a": (abi.NORMAL, torch.tensor(0.2)), "b": (abi.NORMAL, torch.tensor(-0.3)),
           "log_scale": (abi.NORMAL, torch.tensor(0.1)), "s": (abi.GAMMA, torch.tensor(0.8))})

    def beta_bernoulli(m):
        p = m.sample("p", Beta(2, 2), [12])
        m.sample("k", Bernoulli(p))

    check(beta_bernoulli, beta_bernoulli, {"k": flips}, {"p": (abi.BETA, 0.8 * torch.rand(12) + 0.1)})

    def unpacked(m):
        a, b = m.sample("ab", Normal(0, 1), [2])
        m.sample("y", Normal(a + b * x, 1.0))

    check(unpacked, unpacked, {"y": y}, {"ab": (abi.NORMAL, torch.tensor([0.3, -0.2]))})

    def offset(m):
        a = m.sample("a", Normal(0, 1))
        m.sample("y", Poisson(torch.exp(a + torch.log(exposure))))

    check(offset, offset, {"y": counts}, {"a": (abi.NORMAL, torch.tensor(0.4))})

    x2 = torch.randn(n)

    def several_covariates(m):                       # a + b1 x1 + b2 x2: a dense site over [x1 x2], built at trace time
        a = m.sample("a", Normal(0, 1))
        b1 = m.sample("b1", Normal(0, 1))
        b2 = m.sample("b2", Normal(0, 1))
        s = m.sample("s", Gamma(2, 2))
        m.sample("y", Normal(a + b1 * x - b2 * x2 / 3, s))
        m.sample("k", Poisson(torch.exp(0.1 * (b1 * w + b2 * x2))))

    plan = check(several_covariates, several_covariates, {"y": y, "k": counts},
                 {"a": (abi.NORMAL, torch.tensor(0.2)), "b1": (abi.NORMAL, torch.tensor(-0.3)),
                  "b2": (abi.NORMAL, torch.tensor(0.6)), "s": (abi.GAMMA, torch.tensor(0.8))})
    assert [(site.p, site.theta_lat) for site, _ in plan.dense_sites] == [(2, 1), (2, 1)]

    G = 7
    group = torch.randint(0, G, (n,))

    def random_intercepts(m):                        # alpha[group]: a dense site over an indicator matrix
        mu = m.sample("mu", Normal(0, 1))
        tau = m.sample("tau", Gamma(2, 2))
        alpha = m.sample("alpha", Normal(0, tau), [G])
        s = m.sample("s", Gamma(2, 2))
        m.sample("y", Normal(mu + alpha[group], s))
        m.sample("k", Poisson(torch.exp(alpha[group - G])))          # negative indices wrap like torch's

    plan = check(random_intercepts, random_intercepts, {"y": y, "k": counts},
                 {"mu": (abi.NORMAL, torch.tensor(0.2)), "tau": (abi.GAMMA, torch.tensor(0.9)),
                  "alpha": (abi.NORMAL, 0.3 * torch.randn(G)), "s": (abi.GAMMA, torch.tensor(0.8))})
    assert [(site.p, site.theta_lat, site.icpt_lat) for site, _ in plan.dense_sites] == [(G, 2, 0), (G, 2, -1)]

    X4 = torch.randn(n, 4)

    def mixed(m):                                    # random intercepts + a named slope; a coefficient block + a named slope
        mu = m.sample("mu", Normal(0, 1))
        alpha = m.sample("alpha", Normal(0, 1), [G])
        s = m.sample("s", Gamma(2, 2))
        b = m.sample("b", Normal(0, 1))
        theta = m.sample("theta", Normal(0, 1), [4])
        m.sample("y", Normal(mu + alpha[group] + b * x, s))
        m.sample("k", Poisson(torch.exp(0.2 * (X4.to(theta.dtype) @ theta - b * w))))

    plan = check(mixed, mixed, {"y": y, "k": counts},
                 {"mu": (abi.NORMAL, torch.tensor(0.2)), "alpha": (abi.NORMAL, 0.3 * torch.randn(G)),
                  "s": (abi.GAMMA, torch.tensor(0.8)), "b": (abi.NORMAL, torch.tensor(-0.4)),
                  "theta": (abi.NORMAL, 0.2 * torch.randn(4))})
    # packed as mu | alpha b theta | s: both sites read one run of columns (alpha b, b theta)
    assert [(site.p, site.theta_lat) for site, _ in plan.dense_sites] == [(G + 1, 1), (5, 1 + G)]

    def centred(m):
        a = m.sample("a", Normal(0, 1))
        b = m.sample("b", Gamma(2, 2))
        m.sample("y", Normal(a - b * (x - x.mean()) / x.std(), 2.0))
        m.sample("y2", Normal(-a, 1.0), [n])

    check(centred, centred, {"y": y, "y2": y}, {"a": (abi.NORMAL, torch.tensor(0.4)), "b": (abi.GAMMA, torch.tensor(0.7))})


def _random_recipe(rng, depth):
    if depth == 0 or rng.random() < 0.25:
        return rng.choice(["a", "b", "theta", "x", "w", 2.0, -0.5, 3.0, 0.25])
    op = rng.choice(["add", "sub", "mul", "div", "neg"])
    if op == "neg":
        return ("neg", _random_recipe(rng, depth - 1))
    return (op, _random_recipe(rng, depth - 1), _random_recipe(rng, depth - 1))


def _run_recipe(recipe, env):
    if not isinstance(recipe, tuple):
        return env.get(recipe, recipe)
    args = [_run_recipe(r, env) for r in recipe[1:]]
    if recipe[0] == "neg":
        return -args[0]
    left, right = args
    if recipe[0] == "add":
        return left + right
    if recipe[0] == "sub":
        return left - right
    if recipe[0] == "mul":
        return left * right
    if isinstance(right, float):
        return left / (right if abs(right) > 1e-3 else 2.0)
    return left / right


def test_random_links_lower_to_tables_that_score_what_the_reference_scores():
    """The property test of tests/test_trace_lowering.py one level down: random link expressions as
    the location of a Normal site (or, exponentiated, as a Poisson rate) go through ``Plan`` and the
    table interpreter; what is not refused must reproduce the oracle's log-joint."""
    import random
    rng = random.Random(11)
    torch.manual_seed(11)
    n = 40
    x, w = torch.randn(n), torch.rand(n) + 0.5
    y, counts = torch.randn(n), torch.poisson(torch.ones(n))
    lowered = refused = 0
    for trial in range(250):
        recipe = _random_recipe(rng, 3)
        poisson = rng.random() < 0.3

        def model(m, recipe=recipe, poisson=poisson):
            env = {"a": m.sample("a", Normal(0, 1)), "b": m.sample("b", Normal(0, 1)),
                   "theta": m.sample("theta", Normal(0, 1), [n]), "x": x, "w": w}
            eta = _run_recipe(recipe, env)
            if not isinstance(eta, torch.Tensor):
                eta = torch.as_tensor(float(eta))
            if poisson:
                m.sample("counts", Poisson(torch.exp(0.05 * eta)), [] if eta.ndim else [n])
            else:
                m.sample("y", Normal(eta, 1.5), [] if eta.ndim else [n])

        latents = {"a": (abi.NORMAL, torch.tensor(0.7)), "b": (abi.NORMAL, torch.tensor(-1.3)),
                   "theta": (abi.NORMAL, 0.5 * torch.randn(n))}
        data = {"counts": counts} if poisson else {"y": y}
        values = {"a": latents["a"][1], "b": latents["b"][1], "theta": latents["theta"][1], "x": x, "w": w}
        reference_eta = _run_recipe(recipe, values)
        if isinstance(reference_eta, torch.Tensor) and not (reference_eta.abs().max() < 200.0):   # also NaN / inf
            continue
        try:
            check(model, model, data, latents)
            lowered += 1
        except NotImplementedError:
            refused += 1
    assert lowered > 60 and refused > 30
