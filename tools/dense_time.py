"""Raw C-ABI timing of the dense sweep (no reference computation, so N = 1e8 fits):
python tools/dense_time.py N mode(1 tf32 | 2 closed form | 3 f16) [reps] [family normal|bernoulli|poisson]"""
import ctypes as C
import os
import subprocess
import sys

import torch

sys.path.insert(0, ".")
from mininf_b200.engine import abi  # noqa: E402

lib = abi.Library(os.environ["MNF_LIB"]) if os.environ.get("MNF_LIB") else abi.load()
N, mode = int(float(sys.argv[1])), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
family = sys.argv[4] if len(sys.argv) > 4 else "normal"
p, S, D = 64, 64, 64
dev = torch.device("cuda:0")
torch.manual_seed(0)
X = torch.empty(N, p, device=dev)
for c in range(0, N, 10_000_000):
    X[c:c + 10_000_000].normal_()
theta = torch.randn(p, device=dev) / 8
if family == "normal":
    y, fam = X @ theta + torch.randn(N, device=dev), abi.NORMAL
elif family == "bernoulli":
    y, fam = torch.bernoulli(torch.sigmoid(X @ theta)), abi.BERNOULLI_LOGITS
else:
    y, fam = torch.poisson(torch.exp(0.3 * (X @ theta))), abi.POISSON
z = (0.1 * torch.randn(S, D, device=dev)).contiguous()
acc = torch.zeros(S, D + 1, device=dev, dtype=torch.float64)
status = torch.zeros(1, device=dev, dtype=torch.int32)
ws_bytes = lib.workspace_bytes(S, D)
ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
stream = torch.cuda.current_stream().cuda_stream
site = abi.DenseSite(family=fam, p=p, n_rows=N, ldx=p, X=X.data_ptr(), y=y.data_ptr(), mask=None, theta_lat=0,
                     icpt_lat=-1, icpt_const=0.0, reserved=0, scale=abi.const_link(1.0), weight=1.0)


def sweep():
    lib.call("mnf_dense_sweep", C.byref(site), mode, z.data_ptr(), S, D, acc.data_ptr(), ws.data_ptr(), ws_bytes,
             status.data_ptr(), stream)


for _ in range(3):
    sweep()
torch.cuda.synchronize()
for block in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        sweep()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    smi = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader"],
                         capture_output=True, text=True).stdout.strip()
    print(f"mode {mode} {family} N={N:.0e}: {ms:.3f} ms/sweep  {N * (4 * p + 4) / ms / 1e6:.0f} GB/s algorithmic "
          f"({N * (4 * p + 4) / ms / 1e6 / 6548.2:.3f} of 6548) | status {int(status.item())} | {smi}", flush=True)
