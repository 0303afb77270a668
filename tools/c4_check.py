"""Throughput of the row-latent sweep (config C4 shape) through the public API."""
import sys
import torch
sys.path.insert(0, ".")
import mininf_b200 as mininf
from oracle import configs
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
p, S = 32, 32
dev = "cuda:0"
cfg = configs.feature_uncertainty(N, p, device=dev)
approx, leaves = cfg.approximation(device=dev)
loss = mininf.nn.EvidenceLowerBoundLoss(S, check="sync")
cond = mininf.condition(lambda: cfg.model(mininf), **cfg.data)
for _ in range(2):
    l = loss(cond, approx); l.backward()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
reps = 3
for _ in range(reps):
    l = loss(cond, approx); l.backward()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
bytes_alg = N * p * 20 + N * 4
print(f"C4 N={N} p={p} S={S}: {ms:.3f} ms/step, {N*S/ms/1e6:.2f} G evals/s (rows x particles), "
      f"{N*p*S/ms/1e6:.1f} G normal draws/s, {bytes_alg/ms/1e6:.1f} GB/s algorithmic, loss {float(l):.4e}")
