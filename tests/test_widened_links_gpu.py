"""Beyond the BASELINE configurations, through the CUDA engine: links the tracer reduces to the
affine form (`a - b*x/2`, `(c + a)*x/2`, `Bernoulli(probs=sigmoid(c - b*x))`) and
the feature-uncertainty example exactly as the reference writes it (`intercept + z * slope`, one
latent feature per row, n = 30), and regression written with one named scalar slope per covariate
(`a + b1*x1 - b2*x2/3`: a dense site over a design matrix built at trace time). Each case is held to values the UNMODIFIED reference produced on
the same data and noise (tests/golden/{affine_links_small,affine_links,feature_example,several_covariates}.npz, made by
tests/golden/make_golden.py) and to a float64 evaluation of the reference algorithm with other
noise (oracle/elbo.py). The host half - trace, link algebra, tables - is covered without a GPU in
tests/test_trace_lowering.py and tests/test_site_table_semantics.py. Needs a B200."""
import numpy as np
import pytest
import torch

import mininf_b200 as mininf
from oracle import elbo

from conftest import EXTRA_GOLDEN_CASES, golden_noise, load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def engine_eval(config, noise, n_particles):
    approx, leaves = config.approximation(device=DEV)
    module = mininf.nn.EvidenceLowerBoundLoss(n_particles, dense_precision="fp32", check="sync")
    loss = module(mininf.condition(lambda: config.model(mininf), **config.data), approx, _noise=noise)
    loss.backward()
    return loss, leaves, module


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def check_plan(case, plan):
    if case == "feature_example":
        assert list(plan.row_groups) == ["z"] and plan.row_groups["z"].p == 1      # a row latent although small
    elif case == "several_covariates":
        assert [(site.p, site.theta_lat) for site, _ in plan.dense_sites] == [(2, 1), (2, 1)]  # [x1 x2], [x3 x2]
    else:
        assert (len(plan.sweep_groups) == 1) == (case == "affine_links") and not plan.row_groups


@pytest.mark.parametrize("case", list(EXTRA_GOLDEN_CASES))
def test_matches_reference_golden(case):
    """The tolerances of test_engine_gpu.py::test_matches_reference_golden: 1e-5 on the loss,
    1e-4 relative L2 on every gradient."""
    config, golden = load_golden(case, device=DEV)
    S = int(golden["n_particles"])
    loss, leaves, module = engine_eval(config, golden_noise(config, golden, DEV), S)
    check_plan(case, module.last_plan)
    assert abs(float(loss) - float(golden["loss"])) <= 1e-5 * abs(float(golden["loss"]))
    for key, leaf in leaves.items():
        assert leaf.grad is not None, key
        assert rel(leaf.grad.cpu().numpy(), golden[f"grad/{key}"]) < 1e-4, key


@pytest.mark.parametrize("case,S", [("affine_links_small", 5), ("affine_links", 16), ("feature_example", 7),
                                     ("several_covariates", 6)])
def test_against_the_float64_oracle(case, S):
    """Other particle counts and fresh noise, against float64 on the CPU."""
    from oracle import configs
    torch.manual_seed(S)
    cpu = EXTRA_GOLDEN_CASES[case](configs)
    gpu = EXTRA_GOLDEN_CASES[case](configs, device=DEV, gen_device="cpu")
    noise = {name: elbo.draw_noise(dist, S) for name, dist in cpu.approximation()[0].items()}
    approx64, leaves64 = cpu.approximation(dtype=torch.float64)
    # (a covariate the model closure captured stays float32: exact in the float64 arithmetic it enters)
    expected = elbo.neg_elbo(cpu.model, {k: v.double() for k, v in cpu.data.items()}, approx64,
                             {k: v.double() for k, v in noise.items()}, S)
    expected.backward()
    loss, leaves, module = engine_eval(gpu, {k: v.to(DEV) for k, v in noise.items()}, S)
    check_plan(case, module.last_plan)
    assert abs(float(loss) - float(expected)) <= 1e-5 * abs(float(expected))
    for key, leaf in leaves.items():
        assert rel(leaf.grad.cpu().numpy(), leaves64[key].grad.numpy()) < 2e-4, key
