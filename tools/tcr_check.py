"""Developer check of the wide-p tcgen05 kernel (csrc/dense_tcr.cuh) through the raw C-ABI against
torch fp64 and the fp32 SIMT kernel on the same GPU (not a test; parity tests live in tests/).
Usage: python tools/tcr_check.py N p S family(normal|bernoulli|poisson) icpt(0|1) [reps]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, ".")
from mininf_b200.engine import abi  # noqa: E402

torch.manual_seed(0)
dev = torch.device("cuda:0")
lib = abi.Library(os.environ["MNF_LIB"]) if os.environ.get("MNF_LIB") else abi.load()   # debug builds
N, p, S = int(float(sys.argv[1])), int(sys.argv[2]), int(sys.argv[3])
family = sys.argv[4] if len(sys.argv) > 4 else "bernoulli"
icpt = bool(int(sys.argv[5])) if len(sys.argv) > 5 else True
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 5
D = p + 1
X = torch.randn(N, p, device=dev)
theta_true = torch.randn(p, device=dev) / p ** 0.5
eta_true = X @ theta_true + 0.3
if family == "normal":
    y = eta_true + torch.randn(N, device=dev)
    fam = abi.NORMAL
elif family == "bernoulli":
    y = torch.bernoulli(torch.sigmoid(eta_true))
    fam = abi.BERNOULLI_LOGITS
else:
    y = torch.poisson(torch.exp(0.3 * eta_true))
    fam = abi.POISSON
z = (0.1 * torch.randn(S, D, device=dev)).contiguous()
acc = torch.zeros(S, D + 1, device=dev, dtype=torch.float64)
status = torch.zeros(1, device=dev, dtype=torch.int32)
ws_bytes = lib.workspace_bytes(S, D)
ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
stream = torch.cuda.current_stream().cuda_stream
dense = abi.DenseSite(family=fam, p=p, n_rows=N, ldx=p, X=X.data_ptr(), y=y.data_ptr(), mask=None,
                      theta_lat=0, icpt_lat=p if icpt else -1, icpt_const=0.0, reserved=0,
                      scale=abi.const_link(0.9), weight=1.0)


def sweep(mode):
    acc.zero_()
    lib.call("mnf_dense_sweep", C.byref(dense), mode, z.data_ptr(), S, D, acc.data_ptr(), ws.data_ptr(),
             ws_bytes, status.data_ptr(), stream)


if N <= 4_000_000:
    z64 = z.double().requires_grad_()
    eta = X.double() @ z64[:, :p].T + (z64[:, p] if icpt else 0.0)
    y64 = y.double()[:, None]
    if family == "normal":
        ll = torch.distributions.Normal(eta, 0.9).log_prob(y64).sum(0)
    elif family == "bernoulli":
        ll = torch.distributions.Bernoulli(logits=eta).log_prob(y64).sum(0)
    else:
        ll = torch.distributions.Poisson(eta.exp()).log_prob(y64).sum(0)
    ll.sum().backward()
    ref = torch.cat([ll.detach()[:, None], z64.grad], 1)
    if not icpt:
        ref[:, 1 + p] = 0
else:
    sweep(abi.DENSE_FP32)
    torch.cuda.synchronize()
    ref = acc.clone()

for name, mode in (("fp32", abi.DENSE_FP32), ("tf32", abi.DENSE_TF32)):
    status.zero_()
    sweep(mode)
    torch.cuda.synchronize()
    ll_err = ((acc[:, 0] - ref[:, 0]).abs() / ref[:, 0].abs()).max()
    g_err = (acc[:, 1:1 + p] - ref[:, 1:1 + p]).norm() / ref[:, 1:1 + p].norm()
    i_err = (acc[:, 1 + p] - ref[:, 1 + p]).norm() / ref[:, 1 + p].norm().clamp_min(1e-30)
    print(f"{name}: per-particle ll max rel {ll_err:.3e} | grad theta rel-l2 {g_err:.3e} | grad icpt rel-l2 "
          f"{i_err:.3e} | status {status.item()}", flush=True)

for name, mode, r in (("tf32", abi.DENSE_TF32, reps), ("fp32", abi.DENSE_FP32, 1)):
    for _ in range(2):
        sweep(mode)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(r):
        sweep(mode)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / r
    print(f"{name}: {ms:.3f} ms/sweep  {N * (4 * p + 4) / ms / 1e6:.1f} GB/s algorithmic  "
          f"{N * S / ms / 1e6:.2f} G evals/s", flush=True)
