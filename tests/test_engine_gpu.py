"""Parity of the CUDA engine (through the public API and the C-ABI underneath) with the oracle
and with the golden fixtures generated from the reference. Needs a B200."""
import numpy as np
import pytest
import torch

import mininf_b200 as mininf
from mininf_b200.engine import abi
from oracle import configs, elbo, handlers

from conftest import GOLDEN_CASES, golden_noise, load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def engine_eval(config, noise, n_particles, precision="fp32", approx=None, leaves=None, closed_form=False):
    if approx is None:
        approx, leaves = config.approximation(device=DEV)
    loss_module = mininf.nn.EvidenceLowerBoundLoss(n_particles, dense_precision=precision, check="sync",
                                                   closed_form=closed_form)
    conditioned = mininf.condition(lambda: config.model(mininf), **config.data)
    loss = loss_module(conditioned, approx, _noise=noise)
    loss.backward()
    return loss, leaves, loss_module


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


@pytest.mark.parametrize("closed_form", [False, True])
@pytest.mark.parametrize("case", list(GOLDEN_CASES))
def test_matches_reference_golden(case, closed_form):
    """fp32 mode against values produced by the unmodified reference (tests/golden): the default
    per-(particle, observation) sweeps and the closed-form statistics paths."""
    config, golden = load_golden(case, device=DEV)
    S = int(golden["n_particles"])
    loss, leaves, _ = engine_eval(config, golden_noise(config, golden, DEV), S, closed_form=closed_form)
    assert loss.ndim == 0 and loss.dtype == torch.float32
    assert abs(float(loss) - float(golden["loss"])) <= 1e-5 * abs(float(golden["loss"]))  # north-star tolerance
    for key, leaf in leaves.items():
        assert leaf.grad is not None, key
        assert rel(leaf.grad.cpu().numpy(), golden[f"grad/{key}"]) < 1e-4, key


@pytest.mark.parametrize("case", ["regression", "regression_sigma"])
def test_tf32_tensor_core_path_matches_golden(case):
    """tcgen05 kernel: TF32 operands (round to nearest), fp32 accumulate. Stated tolerances at
    these few-hundred-row sizes: 1e-4 relative on the loss, 2e-3 relative L2 on gradients. The
    rounding noise is unbiased and averages down as 1/sqrt(N): test_tf32_error_shrinks_with_rows
    asserts the north-star 1e-5 on the loss at N = 2e5."""
    config, golden = load_golden(case, device=DEV)
    S = int(golden["n_particles"])
    loss, leaves, module = engine_eval(config, golden_noise(config, golden, DEV), S, precision="tf32")
    assert module.last_plan.dense_sites[0][1] == abi.DENSE_TF32
    assert abs(float(loss) - float(golden["loss"])) <= 1e-4 * abs(float(golden["loss"]))
    for key, leaf in leaves.items():
        assert rel(leaf.grad.cpu().numpy(), golden[f"grad/{key}"]) < 2e-3, key


@pytest.mark.parametrize("n,p,S,sigma", [(1, 64, 1, False), (127, 64, 3, True), (128, 64, 64, False),
                                         (129, 64, 5, True), (5000, 64, 64, True), (1000, 7, 2, False),
                                         (777, 130, 9, True)])
@pytest.mark.parametrize("precision", ["fp32", "auto"])
def test_regression_against_oracle(n, p, S, sigma, precision):
    """Ragged row counts, one-row and one-particle edge cases, feature counts outside the tensor
    core kernel's shape (those fall to the fp32 kernel under 'auto')."""
    torch.manual_seed(n + p)
    cpu = configs.regression(n, p, sigma_latent=sigma)
    gpu = configs.regression(n, p, sigma_latent=sigma, device=DEV, gen_device="cpu")
    approx_c, leaves_c = cpu.approximation()
    noise = {name: elbo.draw_noise(dist, S) for name, dist in approx_c.items()}
    expected = elbo.neg_elbo(cpu.model, cpu.data, approx_c, noise, S)
    expected.backward()
    loss, leaves, module = engine_eval(gpu, {k: v.to(DEV) for k, v in noise.items()}, S, precision)
    mode = module.last_plan.dense_sites[0][1]
    # "auto": fp16 operands for p = 64 with at most 64 particles, the fp32 kernel for the other shapes here
    assert mode == (abi.DENSE_F16 if precision == "auto" and p == 64 else abi.DENSE_FP32)
    tensor_core = mode != abi.DENSE_FP32
    # tensor-core operands at a few hundred rows: what is left is the unbiased 11-bit rounding of X,
    # which shrinks as 1/sqrt(N) (1e-5 / 1e-4 from ~2e5 rows on, see the fp64-oracle tests below)
    assert abs(float(loss) - float(expected)) <= (1e-4 if tensor_core else 1e-5) * abs(float(expected))
    for key, leaf in leaves.items():
        assert rel(leaf.grad.cpu().numpy(), leaves_c[key].grad.numpy()) < (3e-3 if tensor_core else 1e-4), key


def test_tf32_error_shrinks_with_rows():
    """Rounding errors of the TF32 path are unbiased: relative gradient error at N=2e5 is far
    below the small-N tolerance (fp64 torch on the GPU is the yardstick here)."""
    n, p, S = 200_000, 64, 64
    gpu = configs.regression(n, p, device=DEV)
    approx, leaves = gpu.approximation(device=DEV)
    noise = {"theta": torch.randn(S, p, device=DEV)}
    loss, leaves, _ = engine_eval(gpu, noise, S, "tf32", approx, leaves)
    loc = leaves["theta.loc"].detach().double().requires_grad_()
    scale = leaves["theta.scale"].detach().double().requires_grad_()
    theta = loc + noise["theta"].double() * scale
    eta = gpu.data["X"].double() @ theta.T
    ll = torch.distributions.Normal(eta, 1.0).log_prob(gpu.data["y"].double()[:, None]).sum(0)
    prior = torch.distributions.Normal(0.0, 1.0).log_prob(theta).sum(1)
    ref = -((ll + prior).mean() + torch.distributions.Normal(loc, scale).entropy().sum())
    ref.backward()
    assert abs(float(loss) - float(ref)) <= 1e-5 * abs(float(ref))
    assert rel(leaves["theta.loc"].grad.cpu().numpy(), loc.grad.cpu().numpy()) < 2e-4
    assert rel(leaves["theta.scale"].grad.cpu().numpy(), scale.grad.cpu().numpy()) < 2e-4


@pytest.mark.parametrize("closed_form", [False, True])
@pytest.mark.parametrize("n,S", [(1, 1), (31, 2), (2048, 64), (50_000, 17)])
def test_missing_observations_against_oracle(n, S, closed_form):
    """Masked Poisson + Normal sites sharing a covariate (small-site kernel below 2048 rows,
    fused site sweep from there on)."""
    torch.manual_seed(n)
    cpu = configs.missing(n)
    gpu = configs.missing(n, device=DEV, gen_device="cpu")
    approx_c, leaves_c = cpu.approximation()
    noise = {name: elbo.draw_noise(dist, S) for name, dist in approx_c.items()}
    expected = elbo.neg_elbo(cpu.model, cpu.data, approx_c, noise, S)
    expected.backward()
    loss, leaves, module = engine_eval(gpu, {k: v.to(DEV) for k, v in noise.items()}, S, closed_form=closed_form)
    assert (len(module.last_plan.sweep_groups) == 1) == (n >= 2048)
    assert abs(float(loss) - float(expected)) <= 1e-5 * abs(float(expected))
    for key, leaf in leaves.items():
        np.testing.assert_allclose(leaf.grad.cpu().numpy(), leaves_c[key].grad.numpy(), rtol=2e-4, atol=1e-3)


@pytest.mark.parametrize("n,p,S", [(300, 32, 1), (2000, 8, 5), (520, 17, 64), (9000, 1, 3)])
def test_feature_uncertainty_against_oracle(n, p, S):
    """Config C4: per-observation latent features (row-latent kernel), external noise for parity:
    1-, 8- and 64-particle template instances (two passes at S = 64), p below the warp width."""
    torch.manual_seed(n + p)
    cpu = configs.feature_uncertainty(n, p)
    gpu = configs.feature_uncertainty(n, p, device=DEV, gen_device="cpu")
    approx_c, leaves_c = cpu.approximation()
    noise = {name: elbo.draw_noise(dist, S) for name, dist in approx_c.items()}
    expected = elbo.neg_elbo(cpu.model, cpu.data, approx_c, noise, S)
    expected.backward()
    loss, leaves, module = engine_eval(gpu, {k: v.to(DEV) for k, v in noise.items()}, S)
    assert list(module.last_plan.row_groups) == ["z"] and module.last_plan.D == 2 + p
    assert abs(float(loss) - float(expected)) <= 1e-5 * abs(float(expected))
    for key, leaf in leaves.items():
        assert rel(leaf.grad.cpu().numpy(), leaves_c[key].grad.numpy()) < 2e-4, key


def test_feature_uncertainty_philox_draws():
    """Without external noise the row latents draw in-kernel (Philox): the Monte Carlo estimate
    must agree with a large-sample oracle estimate and be reproducible under the same seed."""
    n, p, S = 4000, 16, 64
    gpu = configs.feature_uncertainty(n, p, device=DEV, gen_device="cpu")
    cpu = configs.feature_uncertainty(n, p)
    for config in (gpu, cpu):       # a tight approximation keeps the Monte Carlo noise of exp(.) small
        cls, params = config.families["z"]
        config.families["z"] = (cls, {"loc": params["loc"], "scale": 0.05 * params["scale"]})
        config.families["population_scale"] = (torch.distributions.Gamma, {
            "concentration": torch.tensor(400.0), "rate": torch.tensor(400.0)})
    approx, leaves = gpu.approximation(device=DEV)
    conditioned = mininf.condition(lambda: gpu.model(mininf), **gpu.data)
    loss_module = mininf.nn.EvidenceLowerBoundLoss(S, check="sync")
    torch.manual_seed(3)
    values = [float(loss_module(conditioned, approx)) for _ in range(4)]
    torch.manual_seed(3)
    again = float(mininf.nn.EvidenceLowerBoundLoss(S, check="sync")(conditioned, approx))
    assert again == values[0] and len(set(values)) == 4
    approx_c, _ = cpu.approximation(requires_grad=False)
    torch.manual_seed(0)
    reference = float(elbo.neg_elbo(cpu.model, cpu.data, approx_c, None, 48))
    assert abs(np.median(values) - reference) < 3e-3 * abs(reference)


def test_integer_exact_mask_and_count_sums():
    """Bit-exact integer work: number of observed entries and sum of observed counts."""
    gpu = configs.missing(100_003, device=DEV, gen_device="cpu")
    raw = gpu.extra["raw"]
    lib = abi.load()
    out = torch.zeros(2, dtype=torch.int64, device=DEV)
    stream = torch.cuda.current_stream().cuda_stream
    lib.call("mnf_masked_count", raw["counts"].data_ptr(), raw["m_counts"].data_ptr(), raw["counts"].numel(),
             out.data_ptr(), stream)
    assert out.tolist() == [int(raw["m_counts"].sum()), int(raw["counts"][raw["m_counts"]].double().sum())]
    lib.call("mnf_masked_count", raw["counts"].data_ptr(), None, raw["counts"].numel(), out.data_ptr(), stream)
    assert out.tolist() == [raw["counts"].numel(), int(raw["counts"].double().sum())]


def test_wide_tensor_core_kernel_matches_golden():
    """dense_tcr.cuh (p = 128, three particles, latent intercept) against the reference's values.
    Stated TF32 tolerances at 333 rows: 1e-4 relative on the loss, 3e-3 relative L2 on gradients."""
    config, golden = load_golden("logistic_wide", device=DEV)
    S = int(golden["n_particles"])
    loss, leaves, module = engine_eval(config, golden_noise(config, golden, DEV), S, precision="tf32")
    site, mode = module.last_plan.dense_sites[0]
    assert mode == abi.DENSE_TF32 and site.p == 128 and site.icpt_lat >= 0
    assert abs(float(loss) - float(golden["loss"])) <= 1e-4 * abs(float(golden["loss"]))
    for key, leaf in leaves.items():
        assert rel(leaf.grad.cpu().numpy(), golden[f"grad/{key}"]) < 3e-3, key


@pytest.mark.parametrize("n,p,S,intercept", [(1, 64, 1, True), (129, 128, 16, False), (1000, 256, 16, True),
                                             (5000, 192, 17, True), (4097, 448, 32, False),
                                             (70_000, 256, 32, True), (3000, 64, 64, True), (40_000, 64, 33, True),
                                             (3000, 128, 64, True), (2000, 192, 40, False), (900, 128, 100, True),
                                             (301, 24, 3, False), (5000, 100, 20, True), (700, 8, 8, True)])
def test_wide_tensor_core_kernel_against_oracle(n, p, S, intercept):
    """Shapes around the tile (128 rows), chunk (64 features), particle-slot (16 / 32) and
    drain-group (8 tiles) boundaries of dense_tcr.cuh, more than 32 particles (two to four passes
    of the wide kernel), feature counts that are not multiples of 64 (zero-filled last chunk),
    and p = 64 with an intercept (dense_tc.cuh),
    minibatch weight included; the fp32 kernel and the oracle agree to 1e-5, the TF32 kernels
    within the stated TF32 tolerance."""
    torch.manual_seed(n + p)
    cpu = configs.logistic(10 * n, n, p=p, intercept=intercept)
    gpu = configs.logistic(10 * n, n, p=p, intercept=intercept, device=DEV, gen_device="cpu")
    approx_c, leaves_c = cpu.approximation()
    noise = {name: elbo.draw_noise(dist, S) for name, dist in approx_c.items()}
    expected = elbo.neg_elbo(cpu.model, cpu.data, approx_c, noise, S)
    expected.backward()
    for precision in ("fp32", "tf32"):
        loss, leaves, module = engine_eval(gpu, {k: v.to(DEV) for k, v in noise.items()}, S, precision)
        assert module.last_plan.dense_sites[0][1] == (abi.DENSE_TF32 if precision == "tf32" else abi.DENSE_FP32)
        tol = 1e-5 if precision == "fp32" else 2e-4
        assert abs(float(loss) - float(expected)) <= tol * abs(float(expected)), precision
        for key, leaf in leaves.items():
            assert rel(leaf.grad.cpu().numpy(), leaves_c[key].grad.numpy()) < (1e-4 if precision == "fp32" else 5e-3), \
                (precision, key)


@pytest.mark.parametrize("p,S", [(128, 24), (64, 64)])
def test_tensor_core_kernels_normal_and_poisson_families_with_intercept(p, S):
    """The Normal (latent sigma, exp link) and Poisson (exp link) epilogues with a latent plus
    constant intercept and a row mask, dense_tcr.cuh (p = 128), dense_tc.cuh and the fp16-operand
    dense_th.cuh (p = 64), against the exact fp32 kernel on the same inputs (raw C-ABI, 20000 rows)."""
    import ctypes
    lib = abi.load()
    torch.manual_seed(9)
    n = 20_000
    D = p + 2
    X = torch.randn(n, p, device=DEV)
    z = (0.05 * torch.randn(S, D, device=DEV)).contiguous()
    stream = torch.cuda.current_stream().cuda_stream
    ws_bytes = lib.workspace_bytes(S, D)
    ws = torch.empty(ws_bytes, device=DEV, dtype=torch.uint8)
    status = torch.zeros(1, device=DEV, dtype=torch.int32)
    mask = (torch.rand(n, device=DEV) < 0.7).to(torch.uint8)
    for family in (abi.NORMAL, abi.POISSON):
        y = torch.randn(n, device=DEV) if family == abi.NORMAL else torch.poisson(torch.full((n,), 1.3, device=DEV))
        scale = abi.Link(x=None, a_const=0.0, a_lat=p + 1, a_stride=0, b_const=0.0, b_lat=-1, b_stride=0,
                         transform=abi.T_EXP) if family == abi.NORMAL else abi.const_link(1.0)
        site = abi.DenseSite(family=family, p=p, n_rows=n, ldx=p, X=X.data_ptr(), y=y.data_ptr(),
                             mask=mask.data_ptr(), theta_lat=0, icpt_lat=p, icpt_const=0.25, reserved=0,
                             scale=scale, weight=3.0)
        results = []
        for mode in (abi.DENSE_FP32, abi.DENSE_TF32) + ((abi.DENSE_F16,) if p == 64 else ()):
            acc = torch.zeros(S, D + 1, device=DEV, dtype=torch.float64)
            lib.call("mnf_dense_sweep", ctypes.byref(site), mode, z.data_ptr(), S, D, acc.data_ptr(), ws.data_ptr(),
                     ws_bytes, status.data_ptr(), stream)
            torch.cuda.synchronize()
            results.append(acc.cpu().numpy())
        exact = results[0]
        assert int(status.item()) == 0
        for fast in results[1:]:
            assert np.max(np.abs(fast[:, 0] - exact[:, 0]) / np.abs(exact[:, 0])) < 2e-4, family
            assert rel(fast[:, 1:1 + p], exact[:, 1:1 + p]) < 5e-3, family          # theta
            assert rel(fast[:, 1 + p], exact[:, 1 + p]) < 5e-3, family              # intercept
            if family == abi.NORMAL:
                assert rel(fast[:, 2 + p], exact[:, 2 + p]) < 1e-3                    # log-sigma


@pytest.mark.parametrize("batch_rows", [400, 1, 4097])
def test_minibatch_scaling_against_oracle(batch_rows):
    """`batch` rescaling (declared / actual rows) on a Bernoulli(logits = X @ theta) site."""
    torch.manual_seed(batch_rows)
    S = 4
    cpu = configs.logistic(100_000, batch_rows, p=64)
    gpu = configs.logistic(100_000, batch_rows, p=64, device=DEV, gen_device="cpu")
    approx_c, leaves_c = cpu.approximation()
    noise = {name: elbo.draw_noise(dist, S) for name, dist in approx_c.items()}
    expected = elbo.neg_elbo(cpu.model, cpu.data, approx_c, noise, S)
    expected.backward()
    for precision in ("fp32", "tf32"):
        loss, leaves, module = engine_eval(gpu, {k: v.to(DEV) for k, v in noise.items()}, S, precision)
        assert module.last_plan.dense_sites[0][0].weight == 100_000 / batch_rows
        tol = 1e-5 if precision == "fp32" else 1e-4
        assert abs(float(loss) - float(expected)) <= tol * abs(float(expected))
        for key, leaf in leaves.items():
            assert rel(leaf.grad.cpu().numpy(), leaves_c[key].grad.numpy()) < (1e-4 if precision == "fp32" else 5e-3)


def test_minibatch_stream_rebinds_the_plan_instead_of_retracing():
    """examples/minibatch.md:44-60 conditions the model on a fresh batch every step. Batches of the
    same layout reuse one plan with patched pointers; values equal those of a fresh evaluation;
    data written in place is picked up; out-of-support data in a later batch is still reported."""
    S, p, rows = 8, 128, 3000
    batches = [configs.logistic(90_000, rows, p=p, batch_id=i, device=DEV, gen_device="cpu", intercept=True)
               for i in range(3)]
    torch.manual_seed(3)
    approx, leaves = batches[0].approximation(device=DEV)
    noise = {name: elbo.draw_noise(dist, S).to(DEV) for name, dist in approx.items()}
    model = lambda: batches[0].model(mininf)   # noqa: E731  one model, many conditionings
    stream = mininf.nn.EvidenceLowerBoundLoss(S, check="sync")
    values, plans = [], []
    for config in batches + batches[:1]:
        values.append(float(stream(mininf.condition(model, **config.data), approx, _noise=noise)))
        plans.append(stream.last_plan)
    assert all(plan is plans[0] for plan in plans) and plans[0].rebindable
    assert values[3] == values[0]
    for config, value in zip(batches, values):
        fresh = mininf.nn.EvidenceLowerBoundLoss(S, check="sync")
        assert float(fresh(mininf.condition(model, **config.data), approx, _noise=noise)) == value
    assert len({round(v, 3) for v in values[:3]}) == 3
    # in-place refill of the same buffers (a pinned-memory loader writing into a staging batch)
    staging = {k: v.clone() for k, v in batches[0].data.items()}
    first = float(stream(mininf.condition(model, **staging), approx, _noise=noise))
    assert first == values[0]
    for key in staging:
        staging[key].copy_(batches[1].data[key])
    assert float(stream(mininf.condition(model, **staging), approx, _noise=noise)) == values[1]
    assert stream.last_plan is plans[0]
    # the support check of the reference (mininf/core.py:183) survives rebinding via the status word
    staging["y"][17] = 2.0
    with pytest.raises(ValueError, match="not in the support"):
        stream(mininf.condition(model, **staging), approx, _noise=noise)


def test_host_batch_stream_feeds_the_rebound_plan():
    """Out-of-core feed (examples/minibatch.md:68-88 with the data in host memory): batches staged
    through the double-buffered ring give exactly the losses of conditioning on device slices, the
    ragged last batch included, in natural and in shuffled order, without retracing."""
    from mininf_b200.stream import HostBatchStream
    S, p, rows, total = 4, 64, 4096, 3 * 4096 + 1000
    full = configs.logistic(total, total, p=p, device="cpu", intercept=True)
    host = {k: v for k, v in full.data.items()}
    torch.manual_seed(5)
    approx, _ = full.approximation(device=DEV)
    noise = {name: elbo.draw_noise(dist, S).to(DEV) for name, dist in approx.items()}

    def model():
        return full.model(mininf)

    def direct(lo, hi):
        module = mininf.nn.EvidenceLowerBoundLoss(S, check="sync")
        data = {k: v[lo:hi].to(DEV) for k, v in host.items()}
        return float(module(mininf.condition(model, **data), approx, _noise=noise))

    expected = [direct(lo, min(lo + rows, total)) for lo in range(0, total, rows)]
    for order in (None, torch.tensor([2, 0, 3, 1])):
        feed = HostBatchStream(host, rows, device=DEV, depth=2, order=order)
        assert len(feed) == 4 and feed.bytes_per_batch == rows * (p * 4 + 4)
        module = mininf.nn.EvidenceLowerBoundLoss(S, check="sync")
        got, plans = [], []
        for batch in feed:
            got.append(float(module(mininf.condition(model, **batch), approx, _noise=noise)))
            plans.append(module.last_plan)
        want = expected if order is None else [expected[i] for i in order.tolist()]
        assert got == want
        full_plans = [plan for plan, index in zip(plans, range(4) if order is None else order.tolist()) if index != 3]
        assert all(plan is full_plans[0] for plan in full_plans)      # full batches share one rebound plan


def test_host_batch_stream_epochs_do_not_overtake_their_consumers():
    """Several asynchronous epochs over an odd number of batches (depth 2): the first copies of an
    epoch reuse slots whose consumers - the previous epoch's last sweeps - may still be running;
    nothing synchronises inside the loop, losses are compared at the end."""
    from mininf_b200.stream import HostBatchStream
    S, p, rows, total = 8, 256, 60_000, 3 * 60_000
    full = configs.logistic(total, total, p=p, device="cpu")
    host = {k: v for k, v in full.data.items()}
    torch.manual_seed(6)
    approx, _ = full.approximation(device=DEV)
    noise = {name: elbo.draw_noise(dist, S).to(DEV) for name, dist in approx.items()}

    def model():
        return full.model(mininf)

    reference = mininf.nn.EvidenceLowerBoundLoss(S, check="sync")
    expected = []
    for lo in range(0, total, rows):
        data = {k: v[lo:lo + rows].to(DEV) for k, v in host.items()}
        expected.append(float(reference(mininf.condition(model, **data), approx, _noise=noise)))
    feed = HostBatchStream(host, rows, device=DEV, depth=2)
    module = mininf.nn.EvidenceLowerBoundLoss(S, check="lazy")
    losses = []
    for _ in range(4):
        for batch in feed:
            losses.append(module(mininf.condition(model, **batch), approx, _noise=noise).detach())
    module.synchronize()
    assert [float(value) for value in losses] == expected * 4


def test_graphed_step_trains_like_the_eager_loop():
    """`GraphedStep` replays the whole SVI step from one CUDA graph: every replay draws new Philox
    noise (device-side call index), parameters move, and the fit matches the eager README loop."""
    torch.manual_seed(0)
    config = configs.coin(device=DEV)
    approximation = mininf.nn.ParameterizedDistribution(
        torch.distributions.Beta, concentration0=torch.tensor(2.0, device=DEV),
        concentration1=torch.tensor(2.0, device=DEV))
    conditioned = mininf.condition(lambda: config.model(mininf), **config.data)
    optimizer = torch.optim.Adam(approximation.parameters(), lr=0.02, capturable=True)
    loss = mininf.nn.EvidenceLowerBoundLoss(n_particles=8)
    step = mininf.nn.GraphedStep(loss, conditioned, lambda: {"theta": approximation()}, optimizer)
    plan = loss.last_plan
    before = int(plan.step_counter.item())
    values = [float(step()) for _ in range(5)]
    assert int(plan.step_counter.item()) == before + 5
    assert len(set(values)) == 5                      # fresh noise on every replay
    for _ in range(600):
        step()
    loss.synchronize()
    assert abs(float(approximation().mean) - 11 / 14) < 0.05

    # dense regression: graphed and eager training reach the same posterior mean
    torch.manual_seed(1)
    config = configs.regression(20_000, 64, device=DEV, gen_device="cpu")
    fits = []
    for graphed in (False, True):
        q = mininf.nn.ParameterizedDistribution(torch.distributions.Normal, loc=torch.zeros(64, device=DEV),
                                                scale=0.1 * torch.ones(64, device=DEV))
        opt = torch.optim.Adam(q.parameters(), lr=0.05, capturable=graphed)
        module = mininf.nn.EvidenceLowerBoundLoss(n_particles=16)
        model = mininf.condition(lambda: config.model(mininf), **config.data)
        if graphed:
            run = mininf.nn.GraphedStep(module, model, lambda: {"theta": q()}, opt)
        else:
            def run():
                opt.zero_grad()
                module(model, {"theta": q()}).backward()
                opt.step()
        for _ in range(400):
            run()
        module.synchronize()
        fits.append(q().mean.detach().cpu().numpy())
    assert rel(fits[1], fits[0]) < 0.06        # two independent stochastic runs
    assert rel(fits[1], config.extra["theta_true"].cpu().numpy()) < 0.1


def test_readme_training_loop_recovers_posterior():
    """README.md:57-70: Beta approximation of the coin bias trained with Adam through the
    unchanged API; exact posterior is Beta(11, 3)."""
    torch.manual_seed(0)
    config = configs.coin(device=DEV)
    approximation = mininf.nn.ParameterizedDistribution(
        torch.distributions.Beta, concentration0=torch.tensor(2.0, device=DEV),
        concentration1=torch.tensor(2.0, device=DEV))
    conditioned = mininf.condition(lambda: config.model(mininf), **config.data)
    optimizer = torch.optim.Adam(approximation.parameters(), lr=0.02)
    loss = mininf.nn.EvidenceLowerBoundLoss(n_particles=8)
    for _ in range(600):
        optimizer.zero_grad()
        loss(conditioned, {"theta": approximation()}).backward()
        optimizer.step()
    fitted = approximation()
    mean = float(fitted.mean)
    assert abs(mean - 11 / 14) < 0.05
    assert 6 < float(fitted.concentration1) < 18 and 1.5 < float(fitted.concentration0) < 6


def test_loss_contract_and_parameter_gradients():
    """tests/test_nn.py:50-66 of the reference, on the engine: 0-dim finite loss with grad_fn,
    gradients appear on the unconstrained parameters only after backward."""
    def model():
        mininf.sample("x", torch.distributions.Normal(0, 1), 3)

    approximation = mininf.nn.ParameterizedDistribution(
        torch.distributions.Normal, loc=torch.tensor(0.0, device=DEV), scale=torch.ones(3, device=DEV))
    loss = mininf.nn.EvidenceLowerBoundLoss()
    value = loss(model, {"x": approximation()})
    assert value.grad_fn is not None and value.ndim == 0 and np.isfinite(value.item())
    assert all(p.grad is None for p in approximation.distribution_parameters.values())
    value.backward()
    assert all(p.grad is not None for p in approximation.distribution_parameters.values())
    with pytest.raises(TypeError, match="dictionaries of tensors"):
        loss(None, torch.distributions.Normal(0, 1))


def test_log_likelihood_loss_on_the_engine():
    """mininf/nn.py:231-257 / tests/test_nn.py:114-129: joint log-density at point values, with
    gradients, through the same kernels (compared with torch.distributions on the CPU)."""
    cpu = configs.regression(700, 64, sigma_latent=True)
    gpu = configs.regression(700, 64, sigma_latent=True, device=DEV, gen_device="cpu")
    torch.manual_seed(4)
    theta = 0.1 * torch.randn(64)
    sigma = torch.tensor(0.8)
    expected_params = {"theta": theta.clone().requires_grad_(), "sigma": sigma.clone().requires_grad_()}
    log_probs = handlers.evaluate(lambda: cpu.model(handlers), {**cpu.data, **expected_params})
    expected = -sum(value.sum() for value in log_probs.values())
    expected.backward()
    params = {"theta": theta.to(DEV).requires_grad_(), "sigma": sigma.to(DEV).requires_grad_()}
    loss = mininf.nn.LogLikelihoodLoss()(mininf.condition(lambda: gpu.model(mininf), **gpu.data), params)
    assert loss.grad_fn is not None and loss.ndim == 0
    loss.backward()
    assert abs(float(loss) - float(expected)) <= 2e-4 * abs(float(expected))     # TF32 dense path at N = 700
    for name in params:
        assert rel(params[name].grad.cpu().numpy(), expected_params[name].grad.numpy()) < 5e-3, name
    exact = mininf.nn.LogLikelihoodLoss(dense_precision="fp32")(
        mininf.condition(lambda: gpu.model(mininf), **gpu.data), {k: v.detach() for k, v in params.items()})
    assert abs(float(exact) - float(expected)) <= 1e-5 * abs(float(expected))


def test_philox_draws_are_standard_normal_and_seeded():
    """Without external noise Normal factors draw eps in-kernel (Philox4x32-10)."""
    config = configs.regression(256, 64, device=DEV)
    conditioned = mininf.condition(lambda: config.model(mininf), **config.data)
    approx, _ = config.approximation(device=DEV)
    loss = mininf.nn.EvidenceLowerBoundLoss(64, check="sync")
    torch.manual_seed(5)
    first = float(loss(conditioned, approx))
    eps = loss.last_plan.noise.clone()
    assert abs(float(eps.mean())) < 0.06 and abs(float(eps.std()) - 1) < 0.05
    second = float(loss(conditioned, approx))
    assert first != second                      # the stream advances between calls
    torch.manual_seed(5)
    again = mininf.nn.EvidenceLowerBoundLoss(64, check="sync")
    assert float(again(conditioned, approx)) == first


def test_philox_gamma_and_beta_draws_have_the_right_moments():
    """Gamma / Beta factors draw in-kernel too (Marsaglia-Tsang on Philox): check the first two
    moments of the draws the kernel leaves in the plan's z buffer."""
    def model():
        mininf.sample("g", torch.distributions.Gamma(2.0, 2.0), 3)
        mininf.sample("b", torch.distributions.Beta(2.0, 2.0), 3)

    conc = torch.tensor([0.4, 2.5, 30.0], device=DEV)
    rate = torch.tensor([1.5, 0.5, 10.0], device=DEV)
    c1 = torch.tensor([0.5, 3.0, 20.0], device=DEV)
    c0 = torch.tensor([0.7, 1.5, 40.0], device=DEV)
    approx = {"g": torch.distributions.Gamma(conc, rate), "b": torch.distributions.Beta(c1, c0)}
    module = mininf.nn.EvidenceLowerBoundLoss(64, check="sync")
    draws = []
    for _ in range(64):
        assert np.isfinite(float(module(model, approx)))
        draws.append(module.last_plan.z.clone())
    z = torch.cat(draws).double()                     # [4096, 6]
    g, b = z[:, :3], z[:, 3:]
    mean_g, var_g = (conc / rate).double(), (conc / rate ** 2).double()
    mean_b = (c1 / (c1 + c0)).double()
    var_b = (c1 * c0 / ((c1 + c0) ** 2 * (c1 + c0 + 1))).double()
    assert torch.allclose(g.mean(0), mean_g, rtol=0.08) and torch.allclose(g.var(0), var_g, rtol=0.25)
    assert torch.allclose(b.mean(0), mean_b, rtol=0.05) and torch.allclose(b.var(0), var_b, rtol=0.25)
    assert (g > 0).all() and ((b > 0) & (b < 1)).all()


def test_invalid_values_are_reported():
    config = configs.regression(300, 64, device=DEV)
    config.data["y"][17] = float("nan")
    conditioned = mininf.condition(lambda: config.model(mininf), **config.data)
    approx, _ = config.approximation(device=DEV)
    with pytest.raises(ValueError, match="not in the support"):
        mininf.nn.EvidenceLowerBoundLoss(2, check="sync")(conditioned, approx)   # caught at trace time


def test_no_cpu_fallback():
    config = configs.regression(64, 8)
    conditioned = mininf.condition(lambda: config.model(mininf), **config.data)
    approx, _ = config.approximation()
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        mininf.nn.EvidenceLowerBoundLoss()(conditioned, approx)


def test_unsupported_link_raises():
    x = torch.randn(50, device=DEV)

    def model():
        a = mininf.sample("a", torch.distributions.Normal(0, 1))
        mininf.sample("y", torch.distributions.Normal(torch.sin(a * x), 1.0))

    approx = {"a": torch.distributions.Normal(torch.tensor(0.0, device=DEV), torch.tensor(1.0, device=DEV))}
    with pytest.raises(NotImplementedError, match="not a supported link"):
        mininf.nn.EvidenceLowerBoundLoss()(mininf.condition(model, y=torch.randn(50, device=DEV)), approx)


# ---- BASELINE.json sizes: size-independent properties (the oracle cannot run at N = 1e8) ---------
def _bench_regression(n, seed):
    g = torch.Generator(device=DEV)
    g.manual_seed(seed)
    p = 64
    theta_true = torch.randn(p, generator=g, device=DEV) / p ** 0.5
    X = torch.randn(n, p, generator=g, device=DEV)
    y = X @ theta_true + torch.randn(n, generator=g, device=DEV)

    def model():
        theta = mininf.sample("theta", torch.distributions.Normal(0, 1), p)
        with mininf.no_log_prob():
            Xv = mininf.sample("X", torch.distributions.Normal(0, 1), (n, p))
        mininf.sample("y", torch.distributions.Normal(Xv @ theta, 1.0))
    return model, X, y


def _evaluate(model, X, y, noise, precision, S=64):
    loc = (0.05 * torch.ones(64, device=DEV)).requires_grad_()
    scale = (0.1 * torch.ones(64, device=DEV)).requires_grad_()
    module = mininf.nn.EvidenceLowerBoundLoss(S, dense_precision=precision, check="sync")
    loss = module(mininf.condition(model, X=X, y=y), {"theta": torch.distributions.Normal(loc, scale)},
                  _noise={"theta": noise})
    loss.backward()
    return float(loss), loc.grad.clone(), scale.grad.clone()


def test_full_size_tensor_core_kernel_agrees_with_exact_kernel():
    """N = 1e8, p = 64, S = 64 (BASELINE.json config[1]): the tcgen05 paths (fp16 operands - the
    default - and TF32 operands) against the exact fp32 SIMT kernel (itself pinned to the reference
    at small N): the north-star 1e-5 on the loss, 1e-4 relative L2 on gradients."""
    n = 100_000_000
    model, X, y = _bench_regression(n, 123)
    noise = torch.randn(64, 64, generator=torch.Generator(device=DEV).manual_seed(5), device=DEV)
    loss_ex, gl_ex, gs_ex = _evaluate(model, X, y, noise, "fp32")
    for precision in ("auto", "tf32"):
        loss_tc, gl_tc, gs_tc = _evaluate(model, X, y, noise, precision)
        assert abs(loss_tc - loss_ex) <= 1e-5 * abs(loss_ex), precision
        assert float((gl_tc - gl_ex).norm() / gl_ex.norm()) < 1e-4, precision
        assert float((gs_tc - gs_ex).norm() / gs_ex.norm()) < 1e-4, precision
    # integer-exact row accounting at full size: every row is live exactly once
    out = torch.zeros(2, dtype=torch.int64, device=DEV)
    ones = torch.ones(n, device=DEV)
    abi.load().call("mnf_masked_count", ones.data_ptr(), None, n, out.data_ptr(),
                    torch.cuda.current_stream().cuda_stream)
    assert out.tolist() == [n, n]


def _fp64_oracle(config, noise, S):
    """float64 evaluation of the reference algorithm (oracle/elbo.py) on the CPU."""
    data64 = {k: v.double().cpu() for k, v in config.data.items()}
    approx64, leaves64 = config.approximation(dtype=torch.float64)
    expected = elbo.neg_elbo(config.model, data64, approx64, {k: v.double().cpu() for k, v in noise.items()}, S,
                             validate=False)
    expected.backward()
    return float(expected), {k: v.grad.numpy() for k, v in leaves64.items()}


@pytest.mark.parametrize("precision", ["f16", "tf32"])
@pytest.mark.parametrize("rows", [200_000, 2_000_000])
def test_c2_black_box_kernel_meets_the_north_star_tolerance_against_the_fp64_oracle(rows, precision):
    """BASELINE.json config[1] (p = 64, S = 64) through the per-(particle, observation) tcgen05
    kernels (csrc/dense_th.cuh with fp16 operands - the default - and csrc/dense_tc.cuh with TF32
    operands; the default `closed_form=False` path) on seeded row subsamples of
    the benchmark's data recipe, against a float64 evaluation of the reference algorithm: 1e-5
    relative on the loss (the north-star tolerance), 1e-4 relative L2 on the gradients. The error
    that is left is the unbiased TF32 rounding of X (it shrinks as 1/sqrt(N)); theta enters the
    tensor core as hi + lo TF32 pairs, so nothing systematic per particle remains."""
    S = 64
    torch.manual_seed(rows)
    config = configs.regression(rows, 64, device=DEV)          # chunk-seeded recipe, device generator
    approx, leaves = config.approximation(device=DEV)
    noise = {name: elbo.draw_noise(dist, S) for name, dist in config.approximation()[0].items()}
    module = mininf.nn.EvidenceLowerBoundLoss(S, dense_precision=precision, check="sync")
    loss = module(mininf.condition(lambda: config.model(mininf), **config.data), approx,
                  _noise={k: v.to(DEV) for k, v in noise.items()})
    loss.backward()
    (site, mode), = module.last_plan.dense_sites
    assert mode == (abi.DENSE_F16 if precision == "f16" else abi.DENSE_TF32)
    assert abi.load().raw("mnf_dense_tf32_kernel")(site.family, site.p, S) == 1
    expected, grads64 = _fp64_oracle(config, noise, S)
    assert abs(float(loss) - expected) <= 1e-5 * abs(expected)
    for key, leaf in leaves.items():
        assert rel(leaf.grad.cpu().numpy(), grads64[key]) < 1e-4, key


def test_c3_black_box_kernel_meets_the_north_star_tolerance_against_the_fp64_oracle():
    """BASELINE.json config[2] shape (p = 256, S = 16, Bernoulli logits under `batch` scaling)
    through the wide tcgen05 kernel (csrc/dense_tcr.cuh) on a seeded 500 000-row batch against the
    float64 oracle: 1e-5 on the loss, 1e-4 on the gradients."""
    S, rows = 16, 500_000
    torch.manual_seed(3)
    config = configs.logistic(1_000_000_000, rows, p=256, device=DEV)
    approx, leaves = config.approximation(device=DEV)
    noise = {name: elbo.draw_noise(dist, S) for name, dist in config.approximation()[0].items()}
    module = mininf.nn.EvidenceLowerBoundLoss(S, dense_precision="tf32", check="sync")
    loss = module(mininf.condition(lambda: config.model(mininf), **config.data), approx,
                  _noise={k: v.to(DEV) for k, v in noise.items()})
    loss.backward()
    (site, mode), = module.last_plan.dense_sites
    assert mode == abi.DENSE_TF32 and abi.load().raw("mnf_dense_tf32_kernel")(site.family, site.p, S) == 2
    expected, grads64 = _fp64_oracle(config, noise, S)
    assert abs(float(loss) - expected) <= 1e-5 * abs(expected)
    for key, leaf in leaves.items():
        assert rel(leaf.grad.cpu().numpy(), grads64[key]) < 1e-4, key


def test_full_size_wide_kernel_agrees_with_exact_kernel():
    """The C3 batch at full size (1e7 x 256, S = 16): TF32 tcgen05 kernel against the exact fp32
    SIMT kernel, 1e-5 on the loss and 1e-4 on the gradients."""
    S, rows = 16, 10_000_000
    config = configs.logistic(1_000_000_000, rows, p=256, device=DEV)
    noise = {"theta": torch.randn(S, 256, generator=torch.Generator().manual_seed(8)).to(DEV)}
    results = {}
    for precision in ("tf32", "fp32"):
        approx, leaves = config.approximation(device=DEV)
        module = mininf.nn.EvidenceLowerBoundLoss(S, dense_precision=precision, check="sync")
        loss = module(mininf.condition(lambda: config.model(mininf), **config.data), approx, _noise=noise)
        loss.backward()
        results[precision] = (float(loss), {k: v.grad.clone() for k, v in leaves.items()})
    assert abs(results["tf32"][0] - results["fp32"][0]) <= 1e-5 * abs(results["fp32"][0])
    for key, grad in results["tf32"][1].items():
        exact = results["fp32"][1][key]
        assert float((grad - exact).norm() / exact.norm()) < 1e-4, key


def test_row_permutation_invariance_and_determinism():
    """The joint is a sum over rows: permuting (X, y) jointly changes nothing beyond fp32
    reassociation, and repeating a call is bit-identical (fixed-order reductions)."""
    n = 20_000_003                      # ragged last tile
    model, X, y = _bench_regression(n, 321)
    noise = torch.randn(64, 64, generator=torch.Generator(device=DEV).manual_seed(6), device=DEV)
    first = _evaluate(model, X, y, noise, "tf32")
    again = _evaluate(model, X, y, noise, "tf32")
    assert first[0] == again[0] and torch.equal(first[1], again[1]) and torch.equal(first[2], again[2])
    perm = torch.randperm(n, device=DEV)
    shuffled = _evaluate(model, X[perm].contiguous(), y[perm].contiguous(), noise, "tf32")
    assert abs(shuffled[0] - first[0]) <= 2e-6 * abs(first[0])
    assert float((shuffled[1] - first[1]).norm() / first[1].norm()) < 2e-5


# ---------------------------------------------------------------------------------------------
# Normal likelihood, p <= 64, no mask: data-only Gram statistics (csrc/dense_gram.cuh)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("masked", [False, True])
@pytest.mark.parametrize("n,S,p", [(1, 37, 64), (127, 37, 64), (20_001, 37, 64), (1_000_003, 37, 64), (5_000, 100, 64),
                                   (20_001, 37, 32), (3_000, 5, 4), (200_003, 64, 48), (5_000, 200, 64)])
def test_gram_statistics_kernel_against_exact_and_per_particle_kernels(n, S, p, masked, monkeypatch):
    """Raw C-ABI, Normal(a + X theta, exp(s)) with a latent + constant intercept, a latent scale,
    features with non-zero means (the sums of squares do not cancel in the Gram form) and no row
    mask: the Gram path (TF32 X, exact products, fp32/fp64 sums) and the per-particle tcgen05
    kernel (mode MNF_DENSE_TF32; four passes of the wide kernel at S = 100, where the Gram path still
    reads X once) against the exact fp32 SIMT kernel. Ragged and one-row tiles; feature counts below
    64 (columns past p are zero-filled by the TMA unit); with a row mask (30 % of the rows missing, their
    responses NaN, at an address that is not 4-byte aligned) the masked rows must leave every sum."""
    import ctypes
    lib = abi.load()
    torch.manual_seed(n)
    D = p + 2
    X = (torch.randn(n, p, device=DEV) + 0.5 * torch.rand(p, device=DEV)).contiguous()
    y = torch.randn(n, device=DEV) + 0.3
    z = (0.05 * torch.randn(S, D, device=DEV)).contiguous()
    stream = torch.cuda.current_stream().cuda_stream
    ws_bytes = lib.workspace_bytes(S, D)
    ws = torch.empty(ws_bytes, device=DEV, dtype=torch.uint8)
    status = torch.zeros(1, device=DEV, dtype=torch.int32)
    scale = abi.Link(x=None, a_const=0.0, a_lat=p + 1, a_stride=0, b_const=0.0, b_lat=-1, b_stride=0,
                     transform=abi.T_EXP)
    mask_ptr = None
    if masked:
        mask_store = torch.zeros(n + 1, device=DEV, dtype=torch.uint8)
        mask = mask_store[1:] if (n > 127 and p == 64) else mask_store[:n]   # odd address: per-element loads
        mask.copy_(torch.rand(n, device=DEV) < 0.7)
        if n == 1:
            mask.fill_(1)
        y = torch.where(mask.bool(), y, torch.full_like(y, float("nan")))
        mask_ptr = mask.data_ptr()
    site = abi.DenseSite(family=abi.NORMAL, p=p, n_rows=n, ldx=p, X=X.data_ptr(), y=y.data_ptr(), mask=mask_ptr,
                         theta_lat=0, icpt_lat=p, icpt_const=0.25, reserved=0, scale=scale, weight=3.0)

    def sweep(mode):
        acc = torch.zeros(S, D + 1, device=DEV, dtype=torch.float64)
        lib.call("mnf_dense_sweep", ctypes.byref(site), mode, z.data_ptr(), S, D, acc.data_ptr(), ws.data_ptr(),
                 ws_bytes, status.data_ptr(), stream)
        torch.cuda.synchronize()
        assert int(status.item()) == 0
        return acc.cpu().numpy()

    exact = sweep(abi.DENSE_FP32)
    monkeypatch.delenv("MNF_DENSE_NO_GRAM", raising=False)
    gram = sweep(abi.DENSE_TF32_CLOSED_FORM)
    assert np.array_equal(gram, sweep(abi.DENSE_TF32_CLOSED_FORM))      # fixed-order reductions
    if S > 128:                                        # beyond the per-particle kernels: only the Gram path
        assert lib.raw("mnf_dense_tf32_kernel")(abi.NORMAL, p, S) == 3
        monkeypatch.setenv("MNF_DENSE_NO_GRAM", "1")
        assert lib.raw("mnf_dense_tf32_kernel")(abi.NORMAL, p, S) == 0
        monkeypatch.delenv("MNF_DENSE_NO_GRAM")
        per_particle = exact
    else:
        per_particle = sweep(abi.DENSE_TF32)           # the black-box mode never takes the closed form
    assert not np.array_equal(gram, per_particle)                        # the switch selects another kernel (or: not the exact one)
    for name, fast in (("gram", gram), ("per-particle", per_particle)):
        # TF32 rounding of X is unbiased: few rows state the loose tolerance, 1e6 rows the tight one
        tol = 5e-4 if n < 100_000 else 2e-5
        assert np.max(np.abs(fast[:, 0] - exact[:, 0]) / np.abs(exact[:, 0])) < tol, name
        gtol = 5e-3 if n < 100_000 else 5e-4
        assert rel(fast[:, 1:1 + p], exact[:, 1:1 + p]) < gtol, name      # theta
        assert rel(fast[:, 1 + p], exact[:, 1 + p]) < gtol, name          # intercept
        assert rel(fast[:, 2 + p], exact[:, 2 + p]) < gtol, name          # log-sigma


def test_gram_statistics_kernel_flags_nan_responses_and_bad_scales():
    """Per-step device checks of the Gram path (raw C-ABI): a NaN response is outside the Normal
    support (mininf/core.py:186-188) -> MNF_ST_BAD_VALUE; a non-positive scale is an invalid
    parameter (TORCH normal.py:57 arg_constraints) -> MNF_ST_BAD_PARAM."""
    import ctypes
    lib = abi.load()
    n, p, S = 1000, 64, 4
    D = p
    X = torch.randn(n, p, device=DEV)
    z = (0.05 * torch.randn(S, D, device=DEV)).contiguous()
    stream = torch.cuda.current_stream().cuda_stream
    ws_bytes = lib.workspace_bytes(S, D)
    ws = torch.empty(ws_bytes, device=DEV, dtype=torch.uint8)

    def flags(y, sigma, mask=None):
        status = torch.zeros(1, device=DEV, dtype=torch.int32)
        site = abi.DenseSite(family=abi.NORMAL, p=p, n_rows=n, ldx=p, X=X.data_ptr(), y=y.data_ptr(),
                             mask=None if mask is None else mask.data_ptr(),
                             theta_lat=0, icpt_lat=-1, icpt_const=0.0, reserved=0, scale=abi.const_link(sigma),
                             weight=1.0)
        acc = torch.zeros(S, D + 1, device=DEV, dtype=torch.float64)
        lib.call("mnf_dense_sweep", ctypes.byref(site), abi.DENSE_TF32, z.data_ptr(), S, D, acc.data_ptr(),
                 ws.data_ptr(), ws_bytes, status.data_ptr(), stream)
        torch.cuda.synchronize()
        return int(status.item())

    y = torch.randn(n, device=DEV)
    assert flags(y, 1.0) == 0
    assert flags(y, -1.0) == abi.ST_BAD_PARAM
    y[777] = float("nan")
    assert flags(y, 1.0) == abi.ST_BAD_VALUE
    mask = torch.ones(n, device=DEV, dtype=torch.uint8)
    mask[777] = 0
    assert flags(y, 1.0, mask) == 0                    # a masked NaN is not an observation
    mask[777], mask[5] = 1, 0
    assert flags(y, 1.0, mask) == abi.ST_BAD_VALUE


# ---------------------------------------------------------------------------------------------
# Normal sites through the sufficient-statistics sweep (csrc/site_sweep.cuh::normal_stats_kernel)
# ---------------------------------------------------------------------------------------------
def _normal_site_case(n, offset=0.0, masked=True, covariate=True, misalign=0, seed=0):
    """w ~ Normal(c + d x, sigma) (or Normal(c, sigma) without a covariate) on n elements; the
    float64 copy feeds the oracle, the float32 copy - optionally a misaligned view - the engine."""
    from torch.distributions import Gamma, Normal
    g = torch.Generator().manual_seed(seed)
    x64 = torch.randn(n + misalign, generator=g, dtype=torch.float64)
    w64 = offset - 0.2 + 0.8 * x64 + 0.7 * torch.randn(n + misalign, generator=g, dtype=torch.float64)
    mask = torch.rand(n + misalign, generator=g) > 0.3
    # the engine sees float32 data: the oracle must score exactly those values
    x32, w32 = x64.float(), w64.float()
    x64, w64 = x32.double(), w32.double()

    def make(xv, wv, mv, device):
        xv, wv, mv = xv.to(device)[misalign:], wv.to(device)[misalign:], mv.to(device)[misalign:]

        def model(m):
            c = m.sample("c", Normal(0, 1))
            sigma = m.sample("sigma", Gamma(2, 2))
            if covariate:
                d = m.sample("d", Normal(0, 1))
                m.sample("w", Normal(c + d * xv, sigma))
            else:
                m.sample("w", Normal(c, sigma), wv.shape)

        data = {"w": torch.masked.as_masked_tensor(wv, mv) if masked else wv}
        return model, data

    families = {"c": (Normal, {"loc": torch.tensor(offset + 0.1), "scale": torch.tensor(0.2)}),
                "d": (Normal, {"loc": torch.tensor(0.5), "scale": torch.tensor(0.2)}),
                "sigma": (Gamma, {"concentration": torch.tensor(3.0), "rate": torch.tensor(3.0)})}
    if not covariate:
        del families["d"]
    return make, (x64, w64, mask), (x32, w32, mask), families


@pytest.mark.parametrize("n,offset,masked,covariate,misalign", [
    (4096, 0.0, True, True, 0),          # vector path, whole float4 groups
    (4099, 0.0, True, True, 0),          # ragged tail of three elements
    (5000, 0.0, False, True, 0),         # no mask
    (5001, 0.0, True, True, 1),          # misaligned views: scalar path
    (70_001, 50.0, True, True, 0),       # large common offset: cancellation in sum r^2 (fp64 statistics)
    (3000, 0.0, True, False, 0),         # no covariate: Normal(c, sigma)
])
def test_normal_site_sufficient_statistics(n, offset, masked, covariate, misalign):
    """The Normal site sweep keeps six data-only sums and evaluates every particle from them; it
    must agree with a float64 evaluation of the reference algorithm on the same float32 data."""
    S = 16
    make, data64, data32, families = _normal_site_case(n, offset, masked, covariate, misalign, seed=n)
    torch.manual_seed(n)
    cpu = configs.Config("normal_site", None, {}, families)
    approx64, leaves64 = cpu.approximation(dtype=torch.float64)
    noise = {name: elbo.draw_noise(dist, S) for name, dist in approx64.items()}
    model64, cond64 = make(*data64, "cpu")
    expected = elbo.neg_elbo(lambda m: model64(m), cond64, approx64, noise, S)
    expected.backward()

    model32, cond32 = make(*data32, DEV)
    approx, leaves = cpu.approximation(device=DEV)
    loss_module = mininf.nn.EvidenceLowerBoundLoss(S, check="sync", closed_form=True)
    loss = loss_module(mininf.condition(lambda: model32(mininf), **cond32), approx,
                       _noise={k: v.float().to(DEV) for k, v in noise.items()})
    loss.backward()
    assert len(loss_module.last_plan.sweep_groups) == 1          # the site sweep ran, not the small-site kernel
    assert abs(float(loss) - float(expected)) <= 1e-5 * abs(float(expected))
    for key, leaf in leaves.items():
        np.testing.assert_allclose(leaf.grad.cpu().numpy(), leaves64[key].grad.numpy(), rtol=3e-4, atol=2e-3)


# ---------------------------------------------------------------------------------------------
# Poisson(exp(a + b x)) sites through the Chebyshev-moment sweep (csrc/site_sweep.cuh::
# poisson_moment_kernel) with its device-side fallback to the per-particle kernel
# ---------------------------------------------------------------------------------------------
def _poisson_site_case(n, x_scale=1.0, x_shift=0.0, b_loc=0.5, masked=True, misalign=0, seed=0, constant_x=False):
    from torch.distributions import Normal, Poisson
    g = torch.Generator().manual_seed(seed)
    x64 = x_shift + x_scale * torch.randn(n + misalign, generator=g, dtype=torch.float64)
    if constant_x:
        x64 = torch.full_like(x64, 0.75)
    x32 = x64.float()
    x64 = x32.double()
    counts = torch.poisson(torch.exp((0.3 + b_loc * x64).clamp(max=8.0)), generator=g)
    mask = torch.rand(n + misalign, generator=g) > 0.3

    def make(xv, device, dtype):
        xv, cv, mv = xv.to(device)[misalign:], counts.to(device=device, dtype=dtype)[misalign:], mask.to(device)[misalign:]

        def model(m):
            a = m.sample("a", Normal(0, 1))
            b = m.sample("b", Normal(0, 1))
            m.sample("counts", Poisson((a + b * xv).exp()))

        return model, {"counts": torch.masked.as_masked_tensor(cv, mv) if masked else cv}

    families = {"a": (Normal, {"loc": torch.tensor(0.25), "scale": torch.tensor(0.1)}),
                "b": (Normal, {"loc": torch.tensor(b_loc), "scale": torch.tensor(0.1)})}
    return make, x64, x32, families


def _poisson_site_eval(make, x64, x32, families, S, seed):
    torch.manual_seed(seed)
    cpu = configs.Config("poisson_site", None, {}, families)
    approx64, leaves64 = cpu.approximation(dtype=torch.float64)
    noise = {name: elbo.draw_noise(dist, S) for name, dist in approx64.items()}
    model64, cond64 = make(x64, "cpu", torch.float64)
    expected = elbo.neg_elbo(lambda m: model64(m), cond64, approx64, noise, S)
    expected.backward()
    model32, cond32 = make(x32, DEV, torch.float32)
    approx, leaves = cpu.approximation(device=DEV)
    loss_module = mininf.nn.EvidenceLowerBoundLoss(S, check="sync", closed_form=True)
    loss = loss_module(mininf.condition(lambda: model32(mininf), **cond32), approx,
                       _noise={k: v.float().to(DEV) for k, v in noise.items()})
    loss.backward()
    assert len(loss_module.last_plan.sweep_groups) == 1
    return float(loss), {k: v.grad.cpu().numpy() for k, v in leaves.items()}, float(expected), \
        {k: v.grad.numpy() for k, v in leaves64.items()}


@pytest.mark.parametrize("n,kwargs", [
    (4096, {}),                                   # moment path, whole float4 groups
    (4099, {}),                                   # ragged tail of three elements
    (5000, {"masked": False}),                    # no mask
    (5001, {"misalign": 1}),                      # misaligned views: per-particle kernel
    (60_001, {"b_loc": -0.8}),                    # negative slope: alternating Bessel weights
    (30_000, {"x_scale": 0.01, "x_shift": 40.0, "b_loc": 0.05}),   # narrow range far from the origin
    (3000, {"constant_x": True}),                 # zero-width range
    (20_000, {"x_scale": 12.0, "b_loc": 0.4}),    # |b| * half-range > 12: the device check falls back
])
def test_poisson_site_moment_sweep_against_float64_oracle(n, kwargs):
    """R_s = sum exp(a_s + b_s x_i) and its x-weighted twin come from 33 data-only Chebyshev
    moments (or, when the in-kernel check rejects the expansion, from the per-particle kernel);
    either way they must agree with a float64 evaluation of the reference algorithm."""
    make, x64, x32, families = _poisson_site_case(n, seed=n, **kwargs)
    loss, grads, expected, grads64 = _poisson_site_eval(make, x64, x32, families, S=16, seed=n)
    assert abs(loss - expected) <= 1e-5 * abs(expected)
    for key in grads:
        np.testing.assert_allclose(grads[key], grads64[key], rtol=3e-4, atol=2e-3 + 1e-6 * abs(expected))


def test_poisson_site_moment_sweep_agrees_with_per_particle_kernel():
    """A/B on identical inputs: the moment path (closed_form=True) against the per-particle MUFU
    kernel (the default), N = 3e6, S = 64, both within 2e-6 of each other on the loss and 2e-5 on
    the gradients."""
    make, x64, x32, families = _poisson_site_case(3_000_000, seed=5)

    def run(closed_form):
        torch.manual_seed(11)
        cfg = configs.Config("poisson_site", None, {}, families)
        approx, leaves = cfg.approximation(device=DEV)
        gen = torch.Generator().manual_seed(3)
        noise = {k: torch.randn(64, generator=gen).to(DEV) for k in families}
        model32, cond32 = make(x32, DEV, torch.float32)
        loss = mininf.nn.EvidenceLowerBoundLoss(64, check="sync", closed_form=closed_form)(
            mininf.condition(lambda: model32(mininf), **cond32), approx, _noise=noise)
        loss.backward()
        return float(loss), torch.stack([v.grad for v in leaves.values()]).double().cpu()

    fast = run(True)
    again = run(True)
    assert fast[0] == again[0] and torch.equal(fast[1], again[1])          # fixed-order reductions
    exact = run(False)
    assert abs(fast[0] - exact[0]) <= 2e-6 * abs(exact[0])
    assert float((fast[1] - exact[1]).norm() / exact[1].norm()) < 2e-5


def test_poisson_site_moment_sweep_flags_invalid_counts_and_propagates_nan_covariates():
    """The per-step device checks of the moment path: a count outside the Poisson support raises
    like the reference's validation; a NaN covariate makes the expansion check fail on the device,
    the per-particle kernel takes over and the non-finite loss is reported."""
    from torch.distributions import Normal, Poisson
    n = 8192
    x = torch.randn(n, device=DEV)
    counts = torch.poisson(torch.exp(0.2 + 0.3 * x))
    approx = {"a": Normal(torch.tensor(0.1, device=DEV), torch.tensor(0.1, device=DEV)),
              "b": Normal(torch.tensor(0.3, device=DEV), torch.tensor(0.1, device=DEV))}

    def model():
        a = mininf.sample("a", Normal(0, 1))
        b = mininf.sample("b", Normal(0, 1))
        mininf.sample("counts", Poisson((a + b * x).exp()))

    module = mininf.nn.EvidenceLowerBoundLoss(8, check="sync", closed_form=True)
    assert torch.isfinite(module(mininf.condition(model, counts=counts), approx))
    counts[17] = 2.5                  # same tensors (cached plan), now with a non-integer count
    with pytest.raises(ValueError, match="support"):
        module(mininf.condition(model, counts=counts), approx)
    counts[17] = 2.0
    assert torch.isfinite(module(mininf.condition(model, counts=counts), approx))
    x[5] = float("nan")
    with pytest.raises(ValueError, match="not finite"):
        module(mininf.condition(model, counts=counts), approx)


def test_python_hyperparameter_changed_between_steps_takes_effect():
    """The reference re-runs the model every step (mininf/nn.py:223-225); the cached plan must not
    freeze a Python constant the model captured: the prior scale below is a closure variable."""
    from torch.distributions import Normal
    x = torch.randn(500, device=DEV)
    y = 0.5 * x + torch.randn(500, device=DEV)
    approx = {"b": Normal(torch.tensor(0.4, device=DEV), torch.tensor(0.2, device=DEV))}
    noise = {"b": torch.linspace(-1, 1, 8).to(DEV)}
    prior_scale = 1.0

    def model():
        b = mininf.sample("b", Normal(0.0, prior_scale))
        mininf.sample("y", Normal(b * x, 1.0))

    module = mininf.nn.EvidenceLowerBoundLoss(8, check="sync")

    def step():
        return float(module(mininf.condition(model, y=y), approx, _noise=noise))

    def expected(scale):
        b = 0.4 + 0.2 * noise["b"].double().cpu()
        log_prior = Normal(0.0, scale).log_prob(b)
        resid = y.double().cpu()[None, :] - b[:, None] * x.double().cpu()[None, :]
        log_lik = Normal(0.0, 1.0).log_prob(resid).sum(1)
        return -float((log_prior + log_lik).mean() + approx["b"].entropy().double().cpu())

    first = step()
    assert abs(first - expected(1.0)) <= 1e-5 * abs(expected(1.0))
    assert step() == first
    prior_scale = 0.1
    changed = step()
    assert abs(changed - expected(0.1)) <= 1e-5 * abs(expected(0.1))
    assert abs(changed - first) > 1.0


def test_poisson_site_moment_sweep_keeps_the_covariate_range_across_steps():
    """The covariate range of the moment path is kept in a device slot across steps and verified
    against every live element: repeated steps on the same data are bit-identical to the first
    (which measured the range), data that leave the cached range are served exactly in the step
    that meets them and by a freshly measured range afterwards, and data that shrink into a
    fraction of the range make the next step measure it again. Every value is held to a float64
    evaluation of the reference algorithm."""
    from torch.distributions import Normal, Poisson
    # the slots are keyed by (covariate address, mask address, element count): an element count no other
    # test uses keeps a slot left behind at a recycled address from entering the bit-equality checks
    n, S = 201_336, 16
    g = torch.Generator(device=DEV).manual_seed(123)
    x = torch.rand(n, generator=g, device=DEV) * 2 - 1
    counts = torch.poisson(torch.exp(0.2 + 0.4 * x), generator=g)
    mask = torch.rand(n, generator=g, device=DEV) > 0.3
    loc = {k: torch.tensor(v, device=DEV) for k, v in (("a", 0.2), ("b", 0.4))}
    approx = {k: Normal(v, torch.tensor(0.05, device=DEV)) for k, v in loc.items()}
    gen = torch.Generator().manual_seed(4)
    noise = {k: torch.randn(S, generator=gen).to(DEV) for k in approx}

    def model():
        a = mininf.sample("a", Normal(0, 1))
        b = mininf.sample("b", Normal(0, 1))
        mininf.sample("counts", Poisson((a + b * x).exp()))

    module = mininf.nn.EvidenceLowerBoundLoss(S, check="sync", closed_form=True)

    def step():
        return float(module(mininf.condition(model, counts=torch.masked.as_masked_tensor(counts, mask)), approx,
                            _noise=noise))

    def expected():
        x64, c64, m = x.double().cpu(), counts.double().cpu(), mask.cpu()
        total = 0.0
        for s in range(S):
            z = {k: float(loc[k]) + 0.05 * float(noise[k][s]) for k in loc}
            eta = z["a"] + z["b"] * x64[m]
            total += float((c64[m] * eta - eta.exp() - torch.lgamma(c64[m] + 1)).sum())
            total += sum(float(Normal(0.0, 1.0).log_prob(torch.tensor(v, dtype=torch.float64))) for v in z.values())
        entropy = sum(float(d.entropy()) for d in approx.values())
        return -(total / S + entropy)

    first = step()
    assert abs(first - expected()) <= 1e-5 * abs(expected())
    assert step() == first and step() == first                  # cached range: same (mid, 1 / half), same sums
    x[7] = 3.0                                                   # same tensors, one live element outside the range
    mask[7] = True
    counts[7] = 4.0
    outside = step()                                             # served by the per-particle kernel
    assert abs(outside - expected()) <= 1e-5 * abs(expected())
    again = step()                                               # range measured again
    assert abs(again - expected()) <= 1e-5 * abs(expected())
    assert step() == again
    x.mul_(0.1)                                                  # the data now use a tenth of the cached range
    narrow = step()
    assert abs(narrow - expected()) <= 1e-5 * abs(expected())
    assert abs(step() - expected()) <= 1e-5 * abs(expected())
    # the developer switch gives the uncached path on identical inputs
    import os
    os.environ["MNF_POISSON_NO_RANGE_CACHE"] = "1"
    try:
        assert abs(step() - expected()) <= 1e-5 * abs(expected())
    finally:
        del os.environ["MNF_POISSON_NO_RANGE_CACHE"]


# ---------------------------------------------------------------------------------------------
# full-size properties of the site sweeps (config C5) and the row-latent sweep (config C4)
# ---------------------------------------------------------------------------------------------
def test_full_size_site_sweeps_permutation_invariance_determinism_and_counts():
    """N = 1e8 elements, S = 64, 30 % missing (BASELINE.json config[4]): the joint is a sum over
    elements, so permuting (x, counts, w, masks) jointly changes nothing beyond reassociation;
    repeating a call is bit-identical; the observed-entry and count sums are integer-exact."""
    from torch.distributions import Gamma, Normal, Poisson
    n, S = 100_000_000, 64
    g = torch.Generator(device=DEV).manual_seed(77)
    x = torch.randn(n, generator=g, device=DEV)
    counts = torch.poisson(torch.exp(0.3 + 0.5 * x), generator=g)
    w = -0.2 + 0.8 * x + 0.7 * torch.randn(n, generator=g, device=DEV)
    m_counts = torch.rand(n, generator=g, device=DEV) > 0.3
    m_w = torch.rand(n, generator=g, device=DEV) > 0.3

    def evaluate(x, counts, w, m_counts, m_w):
        def model():
            a = mininf.sample("a", Normal(0, 1))
            b = mininf.sample("b", Normal(0, 1))
            c = mininf.sample("c", Normal(0, 1))
            d = mininf.sample("d", Normal(0, 1))
            sigma = mininf.sample("sigma", Gamma(2, 2))
            mininf.sample("counts", Poisson((a + b * x).exp()))
            mininf.sample("w", Normal(c + d * x, sigma))

        leaves = {k: (torch.tensor(0.1, device=DEV).requires_grad_(), torch.tensor(0.2, device=DEV).requires_grad_())
                  for k in "abcd"}
        approx = {k: Normal(*v) for k, v in leaves.items()}
        conc, rate = torch.tensor(2.0, device=DEV).requires_grad_(), torch.tensor(2.0, device=DEV).requires_grad_()
        approx["sigma"] = Gamma(conc, rate)
        gen = torch.Generator().manual_seed(9)
        noise = {k: torch.randn(S, generator=gen).to(DEV) for k in "abcd"}
        noise["sigma"] = torch._standard_gamma(torch.full((S,), 2.0), generator=gen).to(DEV)
        module = mininf.nn.EvidenceLowerBoundLoss(S, check="sync", closed_form=True)
        loss = module(mininf.condition(model, counts=torch.masked.as_masked_tensor(counts, m_counts),
                                       w=torch.masked.as_masked_tensor(w, m_w)), approx, _noise=noise)
        loss.backward()
        grads = torch.stack([t.grad for pair in leaves.values() for t in pair] + [conc.grad, rate.grad])
        return float(loss), grads

    first = evaluate(x, counts, w, m_counts, m_w)
    again = evaluate(x, counts, w, m_counts, m_w)
    assert first[0] == again[0] and torch.equal(first[1], again[1])
    perm = torch.randperm(n, device=DEV)
    shuffled = evaluate(x[perm], counts[perm], w[perm], m_counts[perm], m_w[perm])
    assert abs(shuffled[0] - first[0]) <= 2e-6 * abs(first[0])
    assert float((shuffled[1] - first[1]).norm() / first[1].norm()) < 2e-5
    out = torch.zeros(2, dtype=torch.int64, device=DEV)
    abi.load().call("mnf_masked_count", counts.data_ptr(), m_counts.data_ptr(), n, out.data_ptr(),
                    torch.cuda.current_stream().cuda_stream)
    assert out.tolist() == [int(m_counts.sum()), int(counts[m_counts].double().sum())]


def test_full_size_row_latent_sweep_is_seeded_and_its_gradients_are_consistent():
    """N = 1e7 rows, p = 32, S = 32 (BASELINE.json config[3]) with in-kernel Philox draws: the
    same seed reproduces the loss and every gradient bit for bit, and because the draws are a fixed
    function of (seed, call index) the loss is a smooth function of the parameters whose central
    difference along the intercept location must match the analytic gradient."""
    from torch.distributions import Gamma, Normal, Poisson
    n, p, S = 10_000_000, 32, 32
    g = torch.Generator(device=DEV).manual_seed(41)
    slope_true = torch.randn(p, generator=g, device=DEV) / p ** 0.5
    zt = torch.randn(n, p, generator=g, device=DEV)
    y = torch.poisson(torch.exp(0.5 + zt @ slope_true), generator=g)
    x = zt + 0.5 * torch.randn(n, p, generator=g, device=DEV)
    del zt

    def model():
        population_scale = mininf.sample("population_scale", Gamma(2, 2))
        z = mininf.sample("z", Normal(0, population_scale), (n, p))
        mininf.sample("x", Normal(z, 0.5))
        intercept = mininf.sample("intercept", Normal(0, 1))
        slope = mininf.sample("slope", Normal(0, 1), p)
        mininf.sample("y", Poisson((intercept + z @ slope).exp()))

    conditioned = mininf.condition(model, x=x, y=y)
    z_loc, z_scale = x.clone().requires_grad_(), torch.full((n, p), 0.3, device=DEV).requires_grad_()

    def evaluate(intercept_loc, backward=True):
        icpt = torch.tensor(intercept_loc, device=DEV).requires_grad_()
        approx = {"population_scale": Gamma(torch.tensor(40.0, device=DEV), torch.tensor(40.0, device=DEV)),
                  "z": Normal(z_loc, z_scale), "intercept": Normal(icpt, torch.tensor(0.05, device=DEV)),
                  "slope": Normal(0.1 * slope_true, torch.full((p,), 0.02, device=DEV))}
        torch.manual_seed(123)                       # the engine derives its Philox key from torch's generator
        loss = mininf.nn.EvidenceLowerBoundLoss(S, check="sync")(conditioned, approx)
        if backward:
            z_loc.grad = z_scale.grad = None
            loss.backward()
            return float(loss), icpt.grad.item(), z_loc.grad.clone(), z_scale.grad.clone()
        return float(loss)

    first = evaluate(0.4)
    again = evaluate(0.4)
    assert first[0] == again[0] and first[1] == again[1]
    assert torch.equal(first[2], again[2]) and torch.equal(first[3], again[3])
    assert bool(torch.isfinite(first[2]).all()) and bool(torch.isfinite(first[3]).all())
    h = 0.03     # fp32 loss resolution (~2e2 of 2e9) and the cubic term both stay near 1e-3 of the slope
    numeric = (evaluate(0.4 + h, backward=False) - evaluate(0.4 - h, backward=False)) / (2 * h)
    assert abs(numeric - first[1]) <= 5e-3 * abs(first[1]), (numeric, first[1])


# ---------------------------------------------------------------------------------------------
# one native call per step: mnf_plan_create / mnf_elbo_fwd_bwd / mnf_svi_step (include/mininf_b200.h)
# ---------------------------------------------------------------------------------------------
def test_a_step_is_one_native_call_with_few_kernels():
    """The drop-in loss enqueues its whole evaluation through mnf_elbo_fwd_bwd: for the regression
    model that is rsample, the dense sweep, its partial reduction and ONE tail kernel (priors +
    finalize), as counted by the library itself."""
    torch.manual_seed(0)
    config = configs.regression(5000, 64, sigma_latent=True, device=DEV, gen_device="cpu")
    approx, _ = config.approximation(device=DEV)
    module = mininf.nn.EvidenceLowerBoundLoss(16, dense_precision="tf32", check="sync")
    loss = module(mininf.condition(lambda: config.model(mininf), **config.data), approx)
    assert torch.isfinite(loss)
    assert module.last_plan.gpu_launches_per_step == 4


def _svi_case(name):
    from torch.distributions import Gamma, Normal
    PD = mininf.nn.ParameterizedDistribution

    def scalar(value):
        return torch.tensor(value, device=DEV)

    if name == "regression":
        config = configs.regression(3000, 64, sigma_latent=True, device=DEV, gen_device="cpu")

        def modules():
            return {"theta": PD(Normal, loc=torch.zeros(64, device=DEV), scale=0.1 * torch.ones(64, device=DEV)),
                    "sigma": PD(Gamma, concentration=scalar(2.0), rate=scalar(2.0))}
    else:       # per-observation latents (config C4): 2000 x 16 location / scale parameters
        config = configs.feature_uncertainty(2000, 16, device=DEV, gen_device="cpu")

        def modules():
            return {"population_scale": PD(Gamma, concentration=scalar(2.0), rate=scalar(2.0)),
                    "z": PD(Normal, loc=config.data["x"].clone(), scale=torch.ones(2000, 16, device=DEV)),
                    "intercept": PD(Normal, loc=scalar(0.1), scale=scalar(0.2)),
                    "slope": PD(Normal, loc=torch.zeros(16, device=DEV), scale=0.2 * torch.ones(16, device=DEV))}
    return config, modules


@pytest.mark.parametrize("case", ["regression", "features"])
@pytest.mark.parametrize("graph", [False, True])
def test_fused_svi_step_matches_the_torch_adam_loop(graph, case):
    """FusedSVIStep (mnf_svi_step: transforms, ELBO + gradient kernels, chain rule and Adam inside
    the engine; for per-observation latents inside the row-latent sweep itself) against the
    reference-style loop `zero_grad; loss.backward(); Adam.step()` on the same model with the same
    Philox draws: losses and parameters after 25 steps agree to fp32 round-off."""
    config, modules = _svi_case(case)
    conditioned = mininf.condition(lambda: config.model(mininf), **config.data)
    steps, lr, S = 25, 0.02, 8
    torch.manual_seed(7)
    fused_modules = modules()
    fused_loss = mininf.nn.EvidenceLowerBoundLoss(S, dense_precision="fp32")
    fused = mininf.nn.FusedSVIStep(fused_loss, conditioned, fused_modules, lr=lr, graph=graph)
    first_offset, seed = fused._offset, fused._seed
    losses = [float(fused()) for _ in range(steps)]
    fused_loss.synchronize()
    warm = 2 if graph else 0              # graph capture ran two warm-up steps that also updated the parameters
    assert int(fused.steps) == steps + warm
    assert fused.kernels_per_step == 4      # draws, sweep, its partial reduction, tail

    # the same loop with torch: replay the fused step's Philox stream through the plan directly
    from mininf_b200.engine.plan import latent_parameters
    from mininf_b200.nn import _EngineFunction
    torch.manual_seed(7)
    ref_modules = modules()
    optimizer = torch.optim.Adam([p for m in ref_modules.values() for p in m.parameters()], lr=lr)
    ref_loss = mininf.nn.EvidenceLowerBoundLoss(S, dense_precision="fp32", check="off")
    reference_losses = []
    for step in range(steps + warm):
        optimizer.zero_grad()
        approx = {name: module() for name, module in ref_modules.items()}
        plan = ref_loss._plan_for(conditioned, approx)
        params = []
        for spec in plan.all_latents:
            _, p0, p1 = latent_parameters(approx[spec.name])
            shape = spec.shape if len(spec.shape) else torch.Size([])
            params += [p0.expand(shape), p1.expand(shape)]
        loss = _EngineFunction.apply(ref_loss, plan, None, {}, seed, first_offset + step, None, True, False, *params)
        loss.backward()
        optimizer.step()
        reference_losses.append(float(loss))
    np.testing.assert_allclose(losses, reference_losses[warm:], rtol=2e-5)
    for name in ref_modules:
        for key, parameter in ref_modules[name].distribution_parameters.items():
            np.testing.assert_allclose(fused_modules[name].distribution_parameters[key].detach().cpu().numpy(),
                                       parameter.detach().cpu().numpy(), rtol=5e-4, atol=5e-6, err_msg=f"{name}.{key}")


def test_c_program_runs_a_step_from_host_tables_without_python(tmp_path):
    """tests/c/plan_step.c: a plain C consumer of include/mininf_b200.h builds the flat tables of a
    regression model, calls mnf_plan_create / mnf_elbo_fwd_bwd / mnf_svi_step and checks the loss
    against its own double-precision evaluation (no Python, no torch in the process)."""
    import subprocess
    from mininf_b200.engine import build
    root = build.PACKAGE_DIR.parent
    binary = tmp_path / "plan_step"
    subprocess.run(["gcc", "-O1", "-std=c11", "-I", str(root / "include"), "-I", "/usr/local/cuda/include",
                    str(root / "tests" / "c" / "plan_step.c"), "-o", str(binary), str(build.LIB_PATH),
                    "-L/usr/local/cuda/lib64", "-lcudart", "-lm", f"-Wl,-rpath,{build.LIB_DIR}",
                    "-Wl,-rpath,/usr/local/cuda/lib64"], check=True)
    result = subprocess.run([str(binary)], capture_output=True, text=True)
    assert result.returncode == 0, result.stdout + result.stderr
    assert "OK" in result.stdout


def test_fp16_operand_kernel_reports_entries_outside_its_range():
    """MNF_DENSE_F16 is selected per plan only when the design matrix fits fp16; a plan REBOUND to a
    batch with an entry at or above 2^15 is caught by the kernel's per-step check (MNF_ST_RANGE)."""
    torch.manual_seed(1)
    n, p, S = 4096, 64, 8
    config = configs.regression(n, p, device=DEV, gen_device="cpu")
    approx, _ = config.approximation(device=DEV)
    module = mininf.nn.EvidenceLowerBoundLoss(S, check="sync")
    model = lambda: config.model(mininf)  # noqa: E731
    assert torch.isfinite(module(mininf.condition(model, **config.data), approx))
    assert module.last_plan.dense_sites[0][1] == abi.DENSE_F16
    other = {k: v.clone() for k, v in config.data.items()}
    other["X"][17, 3] = 1.0e5
    with pytest.raises(ValueError, match="fp16"):
        module(mininf.condition(model, **other), approx)          # same layout: the plan is rebound
    # traced from scratch on such data the TF32 kernel is chosen instead
    fresh = mininf.nn.EvidenceLowerBoundLoss(S, check="sync")
    assert torch.isfinite(fresh(mininf.condition(model, **other), approx))
    assert fresh.last_plan.dense_sites[0][1] == abi.DENSE_TF32


# ---------------------------------------------------------------------------------------------
# batched posterior predictive: broadcast_samples on the site table (engine/predictive.py, csrc/predict.cuh)
# ---------------------------------------------------------------------------------------------
def test_broadcast_samples_on_the_device_matches_the_reference_contract():
    """tests/test_core.py:324-333 of the reference on CUDA samples: deterministic `value` sites are
    evaluated per sample, given samples pass through, conditioned values are replicated."""
    def model():
        a = mininf.value("a")
        x = mininf.sample("x", torch.distributions.Normal(0, 1))
        assert x.shape == ()
        mininf.value("y", x + a)

    x = torch.randn(7, device=DEV)
    states = mininf.broadcast_samples(mininf.condition(model, a=1.3), x=x)
    assert isinstance(states, mininf.State) and set(states) == {"a", "x", "y"}
    torch.testing.assert_close(states["y"], x + 1.3)
    torch.testing.assert_close(states["x"], x)
    assert states["a"].shape == (7,) and bool((states["a"] == 1.3).all())


def test_broadcast_samples_predictive_example_in_one_launch():
    """examples/predictive.md:22-38 and :82-86: quadratic regression, posterior samples of theta and
    sigma broadcast over the model conditioned on new covariates. The device path (one trace, one
    launch) must give the per-sample loop's results: `prediction` exactly (it is deterministic), the
    drawn `y` in distribution, every site with a leading batch dimension."""
    from torch.distributions.constraints import nonnegative_integer

    def model():
        n = mininf.value("n", 30, support=nonnegative_integer)
        p = mininf.value("p", 3, support=nonnegative_integer)
        x = mininf.sample("x", torch.distributions.Normal(0, 1), n)
        X = mininf.value("X", x[:, None] ** torch.arange(p, device=x.device))
        theta = mininf.sample("theta", torch.distributions.Normal(0, 1), p)
        prediction = mininf.value("prediction", X @ theta)
        sigma = mininf.sample("sigma", torch.distributions.Gamma(2, 2))
        mininf.sample("y", torch.distributions.Normal(prediction, sigma))

    torch.manual_seed(3)
    B, nlin = 5000, 101
    theta = torch.randn(B, 3)
    sigma = 0.2 + torch.rand(B)
    lin = torch.linspace(-2.0, 2.0, nlin)
    host = mininf.broadcast_samples(mininf.condition(model, n=nlin, x=lin), theta=theta[:50], sigma=sigma[:50])
    device = mininf.broadcast_samples(mininf.condition(model, n=nlin, x=lin.to(DEV)), theta=theta.to(DEV), sigma=sigma.to(DEV))
    assert set(device) == set(host) == {"n", "p", "x", "X", "theta", "prediction", "sigma", "y"}
    for key in host:
        assert tuple(device[key].shape) == (B,) + tuple(host[key].shape[1:]), key
    torch.testing.assert_close(device["prediction"][:50].cpu(), host["prediction"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(device["x"][7].cpu(), lin)
    assert int(device["n"][0]) == nlin and device["X"].shape == (B, nlin, 3)
    standardized = ((device["y"] - device["prediction"]) / device["sigma"][:, None]).double()
    assert abs(float(standardized.mean())) < 5.0 / (B * nlin) ** 0.5
    assert abs(float(standardized.var()) - 1.0) < 0.01
    per_sample_spread = (device["y"] - device["prediction"]).std(dim=1)
    assert float((per_sample_spread / device["sigma"] - 1).abs().mean()) < 0.08      # each sample's own sigma


@pytest.mark.parametrize("family", ["poisson_small", "poisson_large", "bernoulli_logits", "bernoulli_probs", "gamma", "beta"])
def test_predictive_sampler_families_have_the_right_moments(family):
    """Every device sampler of csrc/predict.cuh against the analytic mean and variance, with
    parameters that depend on the broadcast samples (so the link is exercised too)."""
    from torch.distributions import Bernoulli, Beta, Gamma, Normal, Poisson
    B, n = 4000, 64
    torch.manual_seed(4)
    a = (0.2 * torch.randn(B)).to(DEV)

    def model():
        a_ = mininf.sample("a", Normal(0, 1))
        if family == "poisson_small":
            mininf.sample("y", Poisson((1.0 + a_).exp()), n)
        elif family == "poisson_large":
            mininf.sample("y", Poisson((4.0 + a_).exp()), n)
        elif family == "bernoulli_logits":
            mininf.sample("y", Bernoulli(logits=0.5 + a_), n)
        elif family == "bernoulli_probs":
            mininf.sample("y", Bernoulli(probs=(a_ - 1.0).exp()), n)
        elif family == "gamma":
            mininf.sample("y", Gamma((1.0 + a_).exp(), 2.0), n)
        else:       # torch's Beta stacks its concentrations (no link through that): constant parameters
            mininf.sample("y", Beta(2.0, 3.0), n)

    y = mininf.broadcast_samples(model, a=a)["y"].double()
    assert y.shape == (B, n)
    a64 = a.double()
    if family.startswith("poisson"):
        rate = ((1.0 if family == "poisson_small" else 4.0) + a64).exp()
        mean, var = rate, rate
        assert bool((y == y.floor()).all()) and bool((y >= 0).all())
    elif family.startswith("bernoulli"):
        prob = torch.sigmoid(0.5 + a64) if family == "bernoulli_logits" else (a64 - 1.0).exp()
        mean, var = prob, prob * (1 - prob)
        assert bool(((y == 0) | (y == 1)).all())
    elif family == "gamma":
        conc = (1.0 + a64).exp()
        mean, var = conc / 2.0, conc / 4.0
    else:
        c1, c0 = torch.full_like(a64, 2.0), torch.full_like(a64, 3.0)
        mean = c1 / (c1 + c0)
        var = c1 * c0 / ((c1 + c0) ** 2 * (c1 + c0 + 1))
    z = (y.mean(1) - mean) / (var / n).sqrt()              # per-sample standardized mean: ~ N(0, 1)
    assert abs(float(z.mean())) < 5.0 / B ** 0.5
    assert abs(float(z.var()) - 1.0) < 0.15
    ratio = float((y.var(1, unbiased=True) / var).mean())   # variance of the draws themselves
    assert abs(ratio - 1.0) < 0.05
