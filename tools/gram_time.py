"""Times mnf_dense_sweep (Normal, p = 64, S = 64, no mask) through the raw C-ABI under the developer
switches of csrc/dense.cu: MNF_DENSE_NO_GRAM (per-particle kernel) and MNF_GRAM_DEV_SKIP (1 = no
MMAs, 2 = no X'y / X'1 sums, 3 = TMA streaming only; results of those runs are wrong by design)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mininf_b200.engine import abi  # noqa: E402

DEV = "cuda:0"
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
p, S = 64, 64
D = p
lib = abi.load()
X = torch.empty(n, p, device=DEV)
for c in range(0, n, 10_000_000):
    X[c:c + 10_000_000].normal_()
y = torch.randn(n, device=DEV)
z = (0.05 * torch.randn(S, D, device=DEV)).contiguous()
ws_bytes = lib.workspace_bytes(S, D)
ws = torch.empty(ws_bytes, device=DEV, dtype=torch.uint8)
status = torch.zeros(1, device=DEV, dtype=torch.int32)
acc = torch.zeros(S, D + 1, device=DEV, dtype=torch.float64)
mask = None
if "masked" in sys.argv[2:]:            # 30 % of the rows missing (their responses NaN)
    mask = (torch.rand(n, device=DEV) < 0.7).to(torch.uint8)
    y = torch.where(mask.bool(), y, torch.full_like(y, float("nan")))
site = abi.DenseSite(family=abi.NORMAL, p=p, n_rows=n, ldx=p, X=X.data_ptr(), y=y.data_ptr(),
                     mask=None if mask is None else mask.data_ptr(),
                     theta_lat=0, icpt_lat=-1, icpt_const=0.0, reserved=0, scale=abi.const_link(1.0), weight=1.0)
stream = torch.cuda.current_stream().cuda_stream


def timed(label, reps=15):
    for _ in range(3):
        lib.call("mnf_dense_sweep", ctypes.byref(site), abi.DENSE_TF32, z.data_ptr(), S, D, acc.data_ptr(),
                 ws.data_ptr(), ws_bytes, status.data_ptr(), stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        lib.call("mnf_dense_sweep", ctypes.byref(site), abi.DENSE_TF32, z.data_ptr(), S, D, acc.data_ptr(),
                 ws.data_ptr(), ws_bytes, status.data_ptr(), stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{label:28s} {ms:7.3f} ms  {n * (4 * p + 4) / ms / 1e6:7.0f} GB/s", flush=True)


if "once" in sys.argv[2:]:      # a few launches for an ncu capture
    timed("gram", reps=2)
    sys.exit(0)

if "series" in sys.argv[2:]:
    # sustained behaviour: 20 back-to-back blocks of 10 launches with the SM clock and power beside them
    import subprocess
    for blk in range(20):
        q = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.active",
                            "--format=csv,noheader", "-i", "0"], capture_output=True, text=True).stdout.strip()
        timed(f"gram block {blk} [{q}]", reps=10)
    sys.exit(0)

for rnd in range(2):
    for label, env in (("gram", {}), ("gram no-mma", {"MNF_GRAM_DEV_SKIP": "1"}),
                       ("gram no-simt", {"MNF_GRAM_DEV_SKIP": "2"}), ("gram tma only", {"MNF_GRAM_DEV_SKIP": "3"}),
                       ("per-particle", {"MNF_DENSE_NO_GRAM": "1"})):
        for k in ("MNF_GRAM_DEV_SKIP", "MNF_DENSE_NO_GRAM"):
            os.environ.pop(k, None)
        os.environ.update(env)
        timed(label)
