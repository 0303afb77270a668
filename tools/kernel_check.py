"""Developer check of the raw C-ABI kernels against torch fp64 on the same GPU (not a test;
the parity tests proper live in tests/). Usage: python tools/kernel_check.py [N]"""
import ctypes as C
import sys
import time

import torch

sys.path.insert(0, ".")
from mininf_b200.engine import abi  # noqa: E402

torch.manual_seed(0)
dev = torch.device("cuda:0")
import os
lib = abi.Library(os.environ["MNF_LIB"]) if os.environ.get("MNF_LIB") else abi.load()
info = lib.device_info()
print("device", info.sm_count, info.cc_major, info.cc_minor, info.max_smem_optin)

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000
p, S = 64, 64
D = p
X = torch.randn(N, p, device=dev)
theta_true = torch.randn(p, device=dev) / p ** 0.5
y = X @ theta_true + torch.randn(N, device=dev)
loc = (0.1 * torch.randn(p, device=dev)).contiguous()
scale = (0.1 * torch.rand(p, device=dev) + 0.05).contiguous()
eps = torch.randn(S, D, device=dev)

latents = (abi.Latent * 1)(abi.Latent(family=abi.NORMAL, numel=p, offset=0, reserved=0,
                                      p0=loc.data_ptr(), p1=scale.data_ptr()))
lat_dev = torch.frombuffer(bytearray(bytes(latents)), dtype=torch.uint8).to(dev)
z = torch.empty(S, D, device=dev)
noise = torch.empty(S, D, device=dev)
acc = torch.empty(S, D + 1, device=dev, dtype=torch.float64)
status = torch.zeros(1, device=dev, dtype=torch.int32)
ws_bytes = lib.workspace_bytes(S, D)
ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
out = torch.empty(1 + 2 * D, device=dev)
stream = torch.cuda.current_stream().cuda_stream

# prior site on theta: Normal(0, 1), value = latent
prior = abi.Site(family=abi.NORMAL, value_lat=0, value=None, mask=None, numel=p, scale=1.0)
prior.param[0] = abi.const_link(0.0)
prior.param[1] = abi.const_link(1.0)
sites_dev = torch.frombuffer(bytearray(bytes((abi.Site * 1)(prior))), dtype=torch.uint8).to(dev)

dense = abi.DenseSite(family=abi.NORMAL, p=p, n_rows=N, ldx=p, X=X.data_ptr(), y=y.data_ptr(),
                      mask=None, theta_lat=0, icpt_lat=-1, icpt_const=0.0, reserved=0,
                      scale=abi.const_link(1.0), weight=1.0)


def step(mode):
    lib.call("mnf_rsample", lat_dev.data_ptr(), 1, S, D, eps.data_ptr(), 0, 0, None, z.data_ptr(),
             noise.data_ptr(), acc.data_ptr(), status.data_ptr(), stream)
    lib.call("mnf_dense_sweep", C.byref(dense), mode, z.data_ptr(), S, D, acc.data_ptr(),
             ws.data_ptr(), ws_bytes, status.data_ptr(), stream)
    lib.call("mnf_small_sites", sites_dev.data_ptr(), 1, p, z.data_ptr(), S, D, acc.data_ptr(),
             status.data_ptr(), stream)
    lib.call("mnf_finalize", lat_dev.data_ptr(), 1, S, D, z.data_ptr(), noise.data_ptr(),
             acc.data_ptr(), 1, out.data_ptr(), None, status.data_ptr(), stream)


# fp64 torch reference on the same device
loc64 = loc.double().requires_grad_()
scale64 = scale.double().requires_grad_()
theta = loc64 + eps.double() * scale64            # [S, p]
eta = X.double() @ theta.T                        # [N, S]
ll = torch.distributions.Normal(eta, 1.0).log_prob(y.double()[:, None]).sum(0)
prior_lp = torch.distributions.Normal(0.0, 1.0).log_prob(theta).sum(1)
ent = torch.distributions.Normal(loc64, scale64).entropy().sum()
loss_ref = -((ll + prior_lp).mean() + ent)
loss_ref.backward()
ref = torch.cat([loss_ref.detach()[None], loc64.grad, scale64.grad])

for name, mode in (("fp32", abi.DENSE_FP32), ("tf32", abi.DENSE_TF32), ("f16", 3)):
    status.zero_()
    try:
        step(mode)
        torch.cuda.synchronize()
    except Exception as ex:  # noqa: BLE001
        print(name, "FAILED", ex)
        continue
    o = out.double()
    rel_loss = abs(o[0] - ref[0]) / abs(ref[0])
    gl = (o[1:1 + D] - ref[1:1 + D]).norm() / ref[1:1 + D].norm()
    gs = (o[1 + D:] - ref[1 + D:]).norm() / ref[1 + D:].norm()
    # per-particle log-lik column
    lj = acc[:, 0] - prior_lp.detach()
    ll_err = ((lj - ll.detach()).abs() / ll.detach().abs()).max()
    print(f"{name}: loss {o[0]:.6f} ref {ref[0]:.6f} rel {rel_loss:.3e} | grad loc rel-l2 {gl:.3e} "
          f"scale rel-l2 {gs:.3e} | per-particle ll max rel {ll_err:.3e} | status {status.item()}")

# timing
for name, mode in (("fp32", abi.DENSE_FP32), ("tf32", abi.DENSE_TF32), ("f16", 3)):
    try:
        for _ in range(3):
            step(mode)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10 if mode != abi.DENSE_FP32 else 2
        e0.record()
        for _ in range(reps):
            step(mode)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        gbs = N * (4 * p + 4) / ms / 1e6
        print(f"{name}: {ms:.3f} ms/step  {gbs:.1f} GB/s algorithmic  {N * S / ms / 1e6:.2f} G evals/s")
    except Exception as ex:  # noqa: BLE001
        print(name, "timing FAILED", ex)
