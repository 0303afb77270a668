"""Row sharding on real GPUs (NCCL, world_size 2): every rank runs the engine on its chunk-owned
rows with ``EvidenceLowerBoundLoss(process_group=True)``; the single all-reduce of the [S][1+D]
accumulator must reproduce the one-GPU evaluation of the whole data set.
Skipped on boxes with fewer than two GPUs (the gloo test covers the decomposition on CPU)."""
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


def _evaluate(config, S, precision, device, noise, group):
    import mininf_b200 as mininf
    approx, leaves = config.approximation(device=device)
    loss_module = mininf.nn.EvidenceLowerBoundLoss(S, dense_precision=precision, check="sync", process_group=group)
    loss = loss_module(mininf.condition(lambda: config.model(mininf), **config.data), approx,
                       _noise={k: v.to(device) for k, v in noise.items()})
    loss.backward()
    return float(loss), {k: v.grad.cpu().numpy() for k, v in leaves.items()}


def _worker(rank, world, port, case, queue):
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
    loss, grads = _evaluate(shard, S, precision, device, noise, True)
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
@pytest.mark.parametrize("case", list(CASES))
def test_two_gpus_reproduce_the_single_gpu_elbo(case):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    context = mp.get_context("spawn")
    queue = context.Queue()
    port = _free_port()
    procs = [context.Process(target=_worker, args=(rank, 2, port, case, queue)) for rank in range(2)]
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
