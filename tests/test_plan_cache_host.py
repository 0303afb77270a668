"""``EvidenceLowerBoundLoss._plan_for`` - when a traced plan is reused, rebound or rebuilt - without
a GPU: ``_build_plan`` is replaced by its own tracing steps on CPU tensors with a dry-run ``Plan``
(no device is touched), everything else is the product code. The reference re-runs the model on
every call (mininf/nn.py:223-225); these cases are the ways a cached plan could go stale."""
import torch
from torch.distributions import Normal

import mininf_b200 as mininf
from mininf_b200.engine import abi
from mininf_b200.engine.plan import Plan, assign_offsets, latent_parameters, row_latent_names, slope_groups
from mininf_b200.engine.trace import Affine, LatentRef, LinkTensor, SiteTableTracer

CPU = torch.device("cpu")


def host_module(n_particles=4):
    module = mininf.nn.EvidenceLowerBoundLoss(n_particles)
    module.traces = 0

    def build(model, approximation):              # EvidenceLowerBoundLoss._build_plan minus the CUDA requirement
        module.traces += 1
        entries, draws = [], {}
        for name, factor in approximation.items():
            family, p0, _ = latent_parameters(factor)
            shape = factor.batch_shape
            entries.append((name, family, shape))
            ref = LatentRef(name, 0) if max(shape.numel(), 1) == 1 else LatentRef(name)
            draws[name] = LinkTensor.wrap(factor.sample(), Affine(a_lat=ref))
        with SiteTableTracer() as tracer:
            mininf.condition(model, **draws)()
        return Plan(tracer.sites, assign_offsets(entries, row_latent_names(tracer.sites), slope_groups(tracer.sites)), n_particles, CPU,
                    dry_run=True)

    module._build_plan = build
    return module


def regression(p=8):
    def model():
        theta = mininf.sample("theta", Normal(0, 1), p)
        with mininf.no_log_prob():
            X = mininf.sample("X", Normal(0, 1), (3000, p))
        mininf.sample("y", Normal(X @ theta, 1.0))
    return model


APPROX = {"theta": Normal(torch.zeros(8), torch.ones(8))}


def batch(seed):
    generator = torch.Generator().manual_seed(seed)
    return {"X": torch.randn(3000, 8, generator=generator), "y": torch.randn(3000, generator=generator)}


def test_same_tensors_reuse_new_batches_rebind_and_layout_changes_retrace():
    module, model = host_module(), regression()
    first, second = batch(1), batch(2)
    plan = module._plan_for(mininf.condition(model, **first), APPROX)
    (site, _), = plan.dense_sites
    assert (site.X, site.y) == (first["X"].data_ptr(), first["y"].data_ptr()) and plan.rebindable
    assert module._plan_for(mininf.condition(model, **first), APPROX) is plan and module.traces == 1
    # the next minibatch: same plan object, pointers patched, no trace
    assert module._plan_for(mininf.condition(model, **second), APPROX) is plan and module.traces == 1
    (site, _), = plan.dense_sites
    assert (site.X, site.y) == (second["X"].data_ptr(), second["y"].data_ptr())
    # written in place: address unchanged, version counter bumped - still the same plan (data are read in place)
    second["y"].add_(1.0)
    assert module._plan_for(mininf.condition(model, **second), APPROX) is plan and module.traces == 1
    # another layout (a ragged last batch) is a plan of its own; the full batches keep theirs
    short = {k: v[:1000].clone() for k, v in first.items()}

    def short_model():
        theta = mininf.sample("theta", Normal(0, 1), 8)
        with mininf.no_log_prob():
            X = mininf.sample("X", Normal(0, 1), (1000, 8))
        mininf.sample("y", Normal(X @ theta, 1.0))

    other = module._plan_for(mininf.condition(short_model, **short), APPROX)
    assert other is not plan and module.traces == 2
    assert module._plan_for(mininf.condition(model, **first), APPROX) is plan and module.traces == 2
    # a non-contiguous view cannot be rebound: traced again
    strided = {"X": torch.randn(3000, 16)[:, ::2], "y": first["y"]}
    assert module._plan_for(mininf.condition(model, **strided), APPROX) is not plan and module.traces == 3


def test_a_scalar_tensor_handed_to_condition_is_never_served_from_a_stale_plan():
    module = host_module()
    x, y = torch.randn(3000), torch.randn(3000)

    def model():
        a = mininf.sample("a", Normal(0, 1))
        scale = mininf.value("noise_scale")
        mininf.sample("y", Normal(a + x, scale))

    approx = {"a": Normal(torch.tensor(0.0), torch.tensor(1.0))}
    noise_scale = torch.tensor(0.5)
    plan = module._plan_for(mininf.condition(model, y=y, noise_scale=noise_scale), approx)
    assert plan.sweep_groups[0][0].param[1].a_const == 0.5
    assert module._plan_for(mininf.condition(model, y=y, noise_scale=noise_scale), approx) is plan      # untouched: reused
    noise_scale.fill_(0.7)                                                                              # in place
    fresh = module._plan_for(mininf.condition(model, y=y, noise_scale=noise_scale), approx)
    assert fresh is not plan and abs(fresh.sweep_groups[0][0].param[1].a_const - 0.7) < 1e-6
    newer = module._plan_for(mininf.condition(model, y=y, noise_scale=torch.tensor(0.9)), approx)       # new tensor
    assert newer is not fresh and abs(newer.sweep_groups[0][0].param[1].a_const - 0.9) < 1e-6
    assert module.traces == 3


def test_a_changed_python_constant_and_a_changed_approximation_structure_retrace():
    module = host_module()
    y = torch.randn(3000)
    settings = {"prior_scale": 1.0}

    def make(prior_scale):
        def model():
            a = mininf.sample("a", Normal(0, prior_scale))
            mininf.sample("y", Normal(a, 1.0), [3000])
        return model

    approx = {"a": Normal(torch.tensor(0.0), torch.tensor(1.0))}
    plan = module._plan_for(mininf.condition(make(settings["prior_scale"]), y=y), approx)
    assert module._plan_for(mininf.condition(make(1.0), y=y), approx) is plan           # re-created closure, same constant
    wider = module._plan_for(mininf.condition(make(2.5), y=y), approx)
    assert wider is not plan and wider._small_global_host[0].param[1].a_const == 2.5
    # the approximation's family is part of the key
    gamma = {"a": torch.distributions.Gamma(torch.tensor(2.0), torch.tensor(2.0))}
    assert module._plan_for(mininf.condition(make(1.0), y=y), gamma) is not plan
    assert module.traces == 3
    # at most eight plans are kept
    for i in range(10):
        module._plan_for(mininf.condition(make(3.0 + i), y=y), approx)
    assert len(module._plans) == 8
    assert abi.ABI_VERSION >= 7


def test_contexts_opened_around_the_loss_call_are_part_of_the_key():
    # `with mininf.batch(N): loss(...)` declares the full size from OUTSIDE the model (every site is then
    # batched, as in the reference); another N is another plan
    module = host_module()
    y = torch.randn(3000)

    def model():
        a = mininf.sample("a", Normal(0, 1), [3000])
        mininf.sample("y", Normal(a, 1.0))

    approx = {"a": Normal(torch.zeros(3000), torch.ones(3000))}
    scales = lambda plan: {site.scale for site in plan._small_observed_host + plan._small_global_host}  # noqa: E731
    plain = module._plan_for(mininf.condition(model, y=y), approx)
    assert scales(plain) == {1.0}
    with mininf.batch(9000):
        tripled = module._plan_for(mininf.condition(model, y=y), approx)
        assert tripled is not plain and scales(tripled) == {3.0}
        assert module._plan_for(mininf.condition(model, y=y), approx) is tripled
    with mininf.batch(30000):
        assert scales(module._plan_for(mininf.condition(model, y=y), approx)) == {10.0}
    assert module._plan_for(mininf.condition(model, y=y), approx) is plain and module.traces == 3
