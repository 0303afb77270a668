"""Throughput of the fused masked site sweep (config C5 shape) through the public API."""
import sys, time
import torch
sys.path.insert(0, ".")
import mininf_b200 as mininf
from oracle import configs
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
S = 64
dev = "cuda:0"
cfg = configs.missing(N, device=dev)
approx, leaves = cfg.approximation(device=dev)
loss = mininf.nn.EvidenceLowerBoundLoss(S, check="sync")
cond = mininf.condition(lambda: cfg.model(mininf), **cfg.data)
for _ in range(3):
    l = loss(cond, approx); l.backward()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
reps = 5
for _ in range(reps):
    l = loss(cond, approx); l.backward()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
bytes_alg = N * (4 + 4 + 4 + 2)
print(f"C5 N={N} S={S}: {ms:.3f} ms/step, {N*S/ms/1e6:.1f} G evals/s (rows x particles), {bytes_alg/ms/1e6:.1f} GB/s algorithmic, loss {float(l):.4e}")
