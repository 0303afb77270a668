"""Per-role wait/compute cycle breakdown of the tcgen05 dense kernel (debug build, CTA 0)."""
import ctypes as C
import sys
import torch
sys.path.insert(0, ".")
from mininf_b200.engine import abi
dll = C.CDLL(sys.argv[1])
N = int(float(sys.argv[2])) if len(sys.argv) > 2 else 4_000_000
MODE = int(sys.argv[3]) if len(sys.argv) > 3 else 1
torch.manual_seed(0)
dev = torch.device("cuda:0")
p, S, D = 64, 64, 64
X = torch.randn(N, p, device=dev); y = torch.randn(N, device=dev)
loc = torch.zeros(p, device=dev); scale = torch.ones(p, device=dev)
eps = torch.randn(S, D, device=dev)
lat = (abi.Latent * 1)(abi.Latent(family=0, numel=p, offset=0, reserved=0, p0=loc.data_ptr(), p1=scale.data_ptr()))
lat_dev = torch.frombuffer(bytearray(bytes(lat)), dtype=torch.uint8).to(dev)
z = torch.empty(S, D, device=dev); noise = torch.empty(S, D, device=dev)
acc = torch.empty(S, D + 1, device=dev, dtype=torch.float64)
status = torch.zeros(1, device=dev, dtype=torch.int32)
ws = torch.zeros(64 << 20, device=dev, dtype=torch.uint8)
dbg = torch.zeros(16384 + 64, device=dev)
dense = abi.DenseSite(family=0, p=p, n_rows=N, ldx=p, X=X.data_ptr(), y=y.data_ptr(), mask=None, theta_lat=0,
                      icpt_lat=-1, icpt_const=0.0, reserved=0, scale=abi.const_link(1.0), weight=1.0)
dll.mnf_debug_buffer.argtypes = [C.c_void_p]
dll.mnf_rsample.argtypes = abi.EXPORTS["mnf_rsample"][1]
dll.mnf_dense_sweep.argtypes = abi.EXPORTS["mnf_dense_sweep"][1]
assert dll.mnf_debug_buffer(dbg.data_ptr()) == 0
st = torch.cuda.current_stream().cuda_stream
dll.mnf_rsample(lat_dev.data_ptr(), 1, S, D, eps.data_ptr(), 0, 0, None, z.data_ptr(), noise.data_ptr(), acc.data_ptr(), status.data_ptr(), st)
for _ in range(2):
    dll.mnf_dense_sweep(C.byref(dense), MODE, z.data_ptr(), S, D, acc.data_ptr(), ws.data_ptr(), ws.numel(), status.data_ptr(), st)
torch.cuda.synchronize()
dbg.zero_()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
dll.mnf_dense_sweep(C.byref(dense), MODE, z.data_ptr(), S, D, acc.data_ptr(), ws.data_ptr(), ws.numel(), status.data_ptr(), st)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
tiles = (N + 127) // 128
per_cta = tiles / 148
cnt = dbg[16384:].view(torch.int64).tolist()
names = ["epi(th): wait eta_full", "epi(th): tcgen05.ld + scores + st issue", "epi(th): wait st + arrive", "epi(th): drain", "mma: wait operands", "mma: issue eta", "mma: wait r_ready", "mma: issue G",
         "epi: wait eta_full", "epi: compute", "conv: wait", "conv: convert"]
print(f"N={N} {ms:.3f} ms, {N*260/ms/1e6:.0f} GB/s, tiles/CTA {per_cta:.0f}, ~{ms*1e-3/per_cta*1e9:.0f} ns/tile")
for i, n in enumerate(names):
    div = per_cta
    print(f"  {n:28s} {cnt[i]/div:10.0f} cycles per tile (of that role)")
print(f"  CTA0 total {cnt[15]} cycles -> SM clock {cnt[15]/ms/1e3:.0f} MHz, {cnt[15]/per_cta:.0f} cycles/tile")
