"""Developer timing of config C3 (minibatch logistic regression, p = 256, batch 1e7, S = 16) through
the public API: SVI step time with one resident batch and with two alternating batches (the
streaming case: same plan, pointers rebound). Usage: python tools/c3_check.py [batch_rows] [S]"""
import sys
import time

import torch

sys.path.insert(0, ".")
import mininf_b200 as mininf  # noqa: E402
from oracle import configs  # noqa: E402

rows = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dev = "cuda:0"
p = 256
batches = [configs.logistic(1_000_000_000, rows, p=p, batch_id=i, device=dev) for i in range(2)]
model = lambda: batches[0].model(mininf)   # noqa: E731
approximation = mininf.nn.ParameterizedDistribution(torch.distributions.Normal, loc=torch.zeros(p, device=dev),
                                                    scale=0.1 * torch.ones(p, device=dev))
optimizer = torch.optim.Adam(approximation.parameters(), lr=0.01)
loss_fn = mininf.nn.EvidenceLowerBoundLoss(S)


def step(config):
    optimizer.zero_grad()
    loss = loss_fn(mininf.condition(model, **config.data), {"theta": approximation()})
    loss.backward()
    optimizer.step()
    return loss


for name, order in (("one resident batch", [0]), ("two alternating batches", [0, 1])):
    for i in range(4):
        step(batches[order[i % len(order)]])
    torch.cuda.synchronize()
    reps = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(reps):
        loss = step(batches[order[i % len(order)]])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    host_ms = (time.perf_counter() - t0) * 1e3 / reps
    print(f"{name}: {ms:.3f} ms/step (wall {host_ms:.3f})  {rows * S / ms / 1e6:.2f} G evals/s  "
          f"{rows * (4 * p + 4) / ms / 1e6:.0f} GB/s algorithmic  loss {float(loss):.1f}  "
          f"mode {loss_fn.last_plan.dense_sites[0][1]}", flush=True)
