"""Accuracy of the Gram path against the exact fp32 kernel as the signal-to-residual ratio grows
(the quadratic form is a difference of large sums when the residual is small): y = X theta* + noise
with |theta*| = snr_root, particles scattered tightly around theta*. Prints relative errors of the
per-particle log-likelihood and gradients for the Gram path (per MNF_GRAM_DEV_SKIP flush override)
and the per-particle tcgen05 kernel. Usage: python tools/gram_accuracy.py [rows]"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mininf_b200.engine import abi  # noqa: E402

DEV = "cuda:0"
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20_000_000
p, S = 64, 64
D = p + 1
lib = abi.load()
torch.manual_seed(0)
X = torch.empty(n, p, device=DEV)
for c in range(0, n, 10_000_000):
    X[c:c + 10_000_000].normal_()
stream = torch.cuda.current_stream().cuda_stream
ws_bytes = lib.workspace_bytes(S, D)
ws = torch.empty(ws_bytes, device=DEV, dtype=torch.uint8)
status = torch.zeros(1, device=DEV, dtype=torch.int32)
scale = abi.Link(x=None, a_const=0.0, a_lat=p, a_stride=0, b_const=0.0, b_lat=-1, b_stride=0, transform=abi.T_EXP)


def sweep(site, z, mode):
    acc = torch.zeros(S, D + 1, device=DEV, dtype=torch.float64)
    lib.call("mnf_dense_sweep", ctypes.byref(site), mode, z.data_ptr(), S, D, acc.data_ptr(), ws.data_ptr(), ws_bytes,
             status.data_ptr(), stream)
    torch.cuda.synchronize()
    return acc.cpu().numpy()


def rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


for snr_root, spread in ((1.0, 0.003), (1.0, 0.1), (3.0, 0.003), (10.0, 0.003), (10.0, 0.1), (30.0, 0.003), (30.0, 0.1)):
    theta = torch.randn(p, device=DEV)
    theta *= snr_root / theta.norm()
    y = X @ theta + torch.randn(n, device=DEV)
    z = torch.cat([theta[None, :] + spread * torch.randn(S, p, device=DEV), torch.zeros(S, 1, device=DEV)], 1).contiguous()
    site = abi.DenseSite(family=abi.NORMAL, p=p, n_rows=n, ldx=p, X=X.data_ptr(), y=y.data_ptr(), mask=None,
                         theta_lat=0, icpt_lat=-1, icpt_const=0.0, reserved=0, scale=scale, weight=1.0)
    for k in ("MNF_GRAM_DEV_SKIP", "MNF_DENSE_NO_GRAM"):
        os.environ.pop(k, None)
    exact = sweep(site, z, abi.DENSE_FP32)
    rows = []
    for label, env in (("gram", {}), ("gram flush 1", {"MNF_GRAM_DEV_SKIP": str(1 << 8)}),
                       ("per-particle", {"MNF_DENSE_NO_GRAM": "1"})):
        for k in ("MNF_GRAM_DEV_SKIP", "MNF_DENSE_NO_GRAM"):
            os.environ.pop(k, None)
        os.environ.update(env)
        fast = sweep(site, z, abi.DENSE_TF32)
        rows.append(f"{label}: logp {np.max(np.abs(fast[:, 0] - exact[:, 0]) / np.abs(exact[:, 0])):.2e} "
                    f"theta {rel(fast[:, 1:1 + p], exact[:, 1:1 + p]):.2e} sigma {rel(fast[:, 1 + p], exact[:, 1 + p]):.2e}")
    print(f"signal/residual = {snr_root ** 2:g}, particle spread {spread}: " + " | ".join(rows), flush=True)
