"""Multi-rank logic on CPU (gloo, world_size 2): rows shard by chunk ownership, every rank
reduces its partial [S][1+D] accumulator over its own rows, ONE all-reduce combines them, and the
latent-valued (prior) sites plus the entropy are added once afterwards - the decomposition
``Plan.step(reduce_fn=...)`` relies on. The per-rank arithmetic is done by the
oracle here; the kernels are covered by the GPU tests."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import configs, elbo, handlers

N, P, S = 1000, 8, 3


def test_rank_rows_partition_the_data_set():
    for n in (1000, 256, 257, 100_000_003, 5):
        for world in (1, 2, 4, 8):
            bounds = [configs.rank_rows(n, r, world) for r in range(world)]
            assert bounds[0][0] == 0 and bounds[-1][1] == n
            assert all(bounds[r][1] == bounds[r + 1][0] for r in range(world - 1))
    full = configs.regression(N, P)
    pieces = [configs.regression(N, P, rows=configs.rank_rows(N, r, 2)).data for r in range(2)]
    assert torch.equal(torch.cat([piece["X"] for piece in pieces]), full.data["X"])
    assert torch.equal(torch.cat([piece["y"] for piece in pieces]), full.data["y"])


def _partial_accumulator(config, z):
    """[S][1+D] float64: column 0 the observed-site log-density of particle s on this rank's rows,
    column 1+d its gradient w.r.t. z[s][d]."""
    acc = torch.zeros(S, 1 + P, dtype=torch.float64)
    for s in range(S):
        theta = z[s].clone().requires_grad_()
        scored = handlers.evaluate(handlers.condition(lambda: config.model(handlers), **config.data), {"theta": theta})
        observed = scored["y"]
        observed.backward()
        acc[s, 0] = observed.detach().double()
        acc[s, 1:] = theta.grad.double()
    return acc


def _worker(rank, world, port, queue):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)                                  # identical draws on every rank
    loc, scale = 0.1 * torch.randn(P), 0.1 + 0.05 * torch.rand(P)
    eps = torch.randn(S, P)
    z = loc + eps * scale
    shard = configs.regression(N, P, rows=configs.rank_rows(N, rank, world))
    acc = _partial_accumulator(shard, z)
    dist.all_reduce(acc)                                  # the single collective of a step
    # global sites (prior of theta) and entropy, identically on every rank
    prior = torch.distributions.Normal(0.0, 1.0)
    acc[:, 0] += prior.log_prob(z).sum(1).double()
    acc[:, 1:] += (-z).double()
    entropy = torch.distributions.Normal(loc, scale).entropy().sum().double()
    loss = -(acc[:, 0].mean() + entropy)
    grad_loc = -acc[:, 1:].mean(0)
    grad_scale = -((acc[:, 1:] * eps.double()).mean(0) + 1.0 / scale.double())
    queue.put((rank, float(loss), grad_loc.numpy(), grad_scale.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        return sock.getsockname()[1]


@pytest.mark.timeout(180)
def test_two_ranks_reproduce_the_single_process_elbo():
    context = mp.get_context("spawn")
    queue = context.Queue()
    port = _free_port()
    procs = [context.Process(target=_worker, args=(rank, 2, port, queue)) for rank in range(2)]
    for proc in procs:
        proc.start()
    results = sorted(queue.get(timeout=150) for _ in procs)
    for proc in procs:
        proc.join(timeout=30)
        assert proc.exitcode == 0

    torch.manual_seed(0)
    loc = (0.1 * torch.randn(P)).requires_grad_()
    scale = (0.1 + 0.05 * torch.rand(P)).requires_grad_()
    eps = torch.randn(S, P)
    full = configs.regression(N, P)
    expected = elbo.neg_elbo(full.model, full.data, {"theta": torch.distributions.Normal(loc, scale)},
                             {"theta": eps}, S)
    expected.backward()
    for rank, loss, grad_loc, grad_scale in results:
        np.testing.assert_allclose(loss, float(expected), rtol=1e-5)
        np.testing.assert_allclose(grad_loc, loc.grad.numpy(), rtol=2e-4, atol=1e-3)
        np.testing.assert_allclose(grad_scale, scale.grad.numpy(), rtol=2e-4, atol=1e-3)
    assert results[0][1] == results[1][1]                 # ranks agree bit for bit after the reduce


class _RecordingLibrary:
    """Stands in for the native library: records the order of the C-ABI calls of a step."""

    def __init__(self, log):
        self.log = log

    def call(self, name, *args):
        if name == "mnf_elbo_fwd_bwd":
            self.log.append((name, int(args[4])))          # the step flags
        elif name == "mnf_plan_launches":
            args[1]._obj.value = 2
        else:
            self.log.append((name,))

    def raw(self, name):
        raise AssertionError(f"unexpected raw call {name}")


def test_plan_step_orders_the_reduction_between_the_two_halves_of_the_step():
    """Product code on the CPU (no kernels): ``Plan.step(reduce_fn=...)`` must enqueue the sweeps over
    the rank's observed rows, THEN reduce the [S][1+D] accumulator across ranks, THEN add the
    latent-valued (prior) sites and finalize - exactly once each; without a reduction (one rank, or
    the engine's own peer exchange) the step is a single native call."""
    import mininf_b200 as mininf
    from mininf_b200.engine import abi
    from mininf_b200.engine.plan import LatentSpec, Plan
    from mininf_b200.engine.trace import Affine, LatentRef, LinkTensor, SiteTableTracer

    config = configs.regression(N, P)
    draws = {"theta": LinkTensor.wrap(torch.randn(P), Affine(a_lat=LatentRef("theta")))}
    with SiteTableTracer() as tracer:
        mininf.condition(mininf.condition(lambda: config.model(mininf), **config.data), **draws)()
    plan = Plan(tracer.sites, [LatentSpec("theta", abi.NORMAL, torch.Size([P]), P, 0)], S, torch.device("cpu"),
                dense_mode="fp32", dry_run=True)
    log = []
    plan.lib = _RecordingLibrary(log)
    plan.step(None, seed=1, offset=2, reduce_fn=lambda acc: log.append(("reduce", tuple(acc.shape))))
    assert log == [("mnf_elbo_fwd_bwd", abi.STEP_ENTROPY | abi.STEP_PRE), ("reduce", (S, 1 + P)),
                   ("mnf_elbo_fwd_bwd", abi.STEP_ENTROPY | abi.STEP_POST)]
    assert plan.gpu_launches_per_step == 4                 # both halves are counted
    log.clear()
    plan.step(None, seed=1, offset=3, with_entropy=False)
    assert log == [("mnf_elbo_fwd_bwd", abi.STEP_ALL)]
