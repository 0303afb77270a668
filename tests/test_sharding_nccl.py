"""Row sharding on real GPUs (world_size 2): every rank runs the engine on its chunk-owned rows
with ``EvidenceLowerBoundLoss(process_group=True)``; combining the [S][1+D] accumulators - by the
engine's own peer-memory exchange over NVLink (``reduce="peer"``, csrc/small.cuh) or by one NCCL
all-reduce (``reduce="nccl"``) - must reproduce the one-GPU evaluation of the whole data set
(SURVEY.md §8e). A sharded FusedSVIStep replayed from a CUDA graph keeps the replicas identical.
Skipped on boxes with fewer than two GPUs (the gloo tests cover the host logic on CPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

CASES = {
    # name: (config factory name, kwargs, particles, dense precision)
    "regression": ("regression", {"n": 40_000, "p": 64, "sigma_latent": True}, 8, "fp32"),
    "regression_tf32": ("regression", {"n": 40_000, "p": 64}, 64, "tf32"),
    "missing": ("missing", {"n": 50_000}, 16, "fp32"),
}


def _build(case, rows, device):
    from oracle import configs
    factory, kwargs, S, precision = CASES[case]
    kwargs = dict(kwargs)
    n = kwargs.pop("n")
    config = getattr(configs, factory)(n, **kwargs, rows=rows, device=device, gen_device="cpu")
    return config, S, precision


def _evaluate(config, S, precision, device, noise, group, reduce="peer"):
    import mininf_b200 as mininf
    approx, leaves = config.approximation(device=device)
    loss_module = mininf.nn.EvidenceLowerBoundLoss(S, dense_precision=precision, check="sync", process_group=group,
                                                   reduce=reduce)
    loss = loss_module(mininf.condition(lambda: config.model(mininf), **config.data), approx,
                       _noise={k: v.to(device) for k, v in noise.items()})
    loss.backward()
    return float(loss), {k: v.grad.cpu().numpy() for k, v in leaves.items()}


def _worker(rank, world, port, case, reduce, queue):
    import torch.distributed as dist
    from oracle import configs, elbo
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    device = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
    n = CASES[case][1]["n"]
    shard, S, precision = _build(case, configs.rank_rows(n, rank, world), device)
    torch.manual_seed(11)                                 # identical draws on every rank
    approx_cpu, _ = shard.approximation()
    noise = {name: elbo.draw_noise(dist_, S) for name, dist_ in approx_cpu.items()}
    loss, grads = _evaluate(shard, S, precision, device, noise, True, reduce)
    expected = None
    if rank == 0:                                         # the whole data set on one GPU, no group
        full, _, _ = _build(case, None, device)
        expected = _evaluate(full, S, precision, device, noise, None)
    queue.put((rank, loss, grads, expected))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        return sock.getsockname()[1]


@pytest.mark.timeout(300)
@pytest.mark.parametrize("reduce", ["peer", "nccl"])
@pytest.mark.parametrize("case", list(CASES))
def test_two_gpus_reproduce_the_single_gpu_elbo(case, reduce):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    context = mp.get_context("spawn")
    queue = context.Queue()
    port = _free_port()
    procs = [context.Process(target=_worker, args=(rank, 2, port, case, reduce, queue)) for rank in range(2)]
    for proc in procs:
        proc.start()
    results = sorted((queue.get(timeout=240) for _ in procs), key=lambda item: item[0])
    for proc in procs:
        proc.join(timeout=60)
        assert proc.exitcode == 0
    expected_loss, expected_grads = results[0][3]
    tol_loss, tol_grad = (1e-5, 2e-4) if CASES[case][3] == "fp32" else (1e-4, 5e-3)
    for rank, loss, grads, _ in results:
        assert abs(loss - expected_loss) <= tol_loss * abs(expected_loss), (rank, loss, expected_loss)
        for key, grad in grads.items():
            err = np.linalg.norm(grad - expected_grads[key]) / max(np.linalg.norm(expected_grads[key]), 1e-30)
            assert err < tol_grad, (rank, key, err)
    assert results[0][1] == results[1][1]                 # ranks agree bit for bit after the reduce


def _svi_worker(rank, world, port, queue):
    """A sharded FusedSVIStep replayed from a CUDA graph: no NCCL call on the step path."""
    import torch.distributed as dist
    import mininf_b200 as mininf
    from oracle import configs
    from torch.distributions import Normal
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    device = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
    n, p, S = 200_000, 64, 16
    shard = configs.regression(n, p, rows=configs.rank_rows(n, rank, world), device=device, gen_device="cpu")
    torch.manual_seed(100 + rank)          # different generator seeds on purpose: rank 0's is broadcast
    modules = {"theta": mininf.nn.ParameterizedDistribution(Normal, loc=torch.zeros(p, device=device),
                                                            scale=0.1 * torch.ones(p, device=device))}
    loss_module = mininf.nn.EvidenceLowerBoundLoss(S, dense_precision="tf32", process_group=True)
    step = mininf.nn.FusedSVIStep(loss_module, mininf.condition(lambda: shard.model(mininf), **shard.data), modules,
                                  lr=0.05, graph=True)
    losses = [float(step()) for _ in range(30)]
    loss_module.synchronize()
    parameters = torch.cat([q.detach().reshape(-1) for q in modules["theta"].parameters()]).cpu().numpy()
    truth = shard.extra["theta_true"].cpu().numpy()
    queue.put((rank, losses, parameters, truth, step.kernels_per_step))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_fused_svi_step_replays_from_a_cuda_graph_and_keeps_replicas_identical():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    context = mp.get_context("spawn")
    queue = context.Queue()
    port = _free_port()
    procs = [context.Process(target=_svi_worker, args=(rank, 2, port, queue)) for rank in range(2)]
    for proc in procs:
        proc.start()
    results = sorted((queue.get(timeout=240) for _ in procs), key=lambda item: item[0])
    for proc in procs:
        proc.join(timeout=60)
        assert proc.exitcode == 0
    (_, losses0, params0, truth, kernels), (_, losses1, params1, _, _) = results
    assert losses0 == losses1 and np.array_equal(params0, params1)      # bit-identical replicas
    assert losses0[-1] < losses0[0] and np.all(np.isfinite(params0))
    assert kernels == 5                                                  # rsample, sweep, reduction, push, tail
    loc = params0[:64]
    assert np.linalg.norm(loc - truth) < np.linalg.norm(truth)           # moving towards the truth
