"""Debug the tcgen05 dense kernel: single tile, dump eta and the G accumulator from TMEM."""
import ctypes as C
import sys
import torch
sys.path.insert(0, ".")
from mininf_b200.engine import abi
dll = C.CDLL("/root/repo/gpurun_out/libdebug.so" if len(sys.argv) < 2 else sys.argv[1])
torch.manual_seed(0)
dev = torch.device("cuda:0")
N, p, S, D = 128, 64, 64, 64
X = torch.randn(N, p, device=dev)
y = torch.randn(N, device=dev)
loc = torch.zeros(p, device=dev); scale = torch.ones(p, device=dev)
eps = torch.randn(S, D, device=dev)
lat = (abi.Latent * 1)(abi.Latent(family=0, numel=p, offset=0, reserved=0, p0=loc.data_ptr(), p1=scale.data_ptr()))
lat_dev = torch.frombuffer(bytearray(bytes(lat)), dtype=torch.uint8).to(dev)
z = torch.empty(S, D, device=dev); noise = torch.empty(S, D, device=dev)
acc = torch.empty(S, D + 1, device=dev, dtype=torch.float64)
status = torch.zeros(1, device=dev, dtype=torch.int32)
ws = torch.zeros(64 << 20, device=dev, dtype=torch.uint8)
dbg = torch.zeros(16384, device=dev)
dense = abi.DenseSite(family=0, p=p, n_rows=N, ldx=p, X=X.data_ptr(), y=y.data_ptr(), mask=None, theta_lat=0,
                      icpt_lat=-1, icpt_const=0.0, reserved=0, scale=abi.const_link(1.0), weight=1.0)
dll.mnf_debug_buffer.argtypes = [C.c_void_p]
assert dll.mnf_debug_buffer(dbg.data_ptr()) == 0
dll.mnf_rsample.argtypes = abi.EXPORTS["mnf_rsample"][1]
dll.mnf_dense_sweep.argtypes = abi.EXPORTS["mnf_dense_sweep"][1]
st = torch.cuda.current_stream().cuda_stream
assert dll.mnf_rsample(lat_dev.data_ptr(), 1, S, D, eps.data_ptr(), 0, 0, z.data_ptr(), noise.data_ptr(), acc.data_ptr(), status.data_ptr(), st) == 0
rc = dll.mnf_dense_sweep(C.byref(dense), 1, z.data_ptr(), S, D, acc.data_ptr(), ws.data_ptr(), ws.numel(), status.data_ptr(), st)
print("rc", rc)
torch.cuda.synchronize()
eta = dbg[:8192].view(128, 64).double()
G = dbg[8192:].view(128, 64).double()
eta_ref = X.double() @ z.double().T
print("eta max abs err", (eta - eta_ref).abs().max().item(), "ref max", eta_ref.abs().max().item())
print("eta err per particle (max over rows):", (eta - eta_ref).abs().max(0).values[:8])
R = (y.double()[:, None] - eta_ref)
G_ref = X.double().T @ R       # [p, S]
print("G tmem: nonzero lanes", (G.abs().sum(1) > 0).nonzero().flatten().tolist())
print("G tmem abs max", G.abs().max().item(), "G_ref abs max", G_ref.abs().max().item())
# try to match the lanes against reference rows
for lane in [0, 1, 15, 16, 31, 32, 33, 48, 64, 96, 112]:
    row = G[lane]
    d = (G_ref - row[None, :]).abs().max(1).values
    j = int(d.argmin())
    print(f"lane {lane}: best ref feature {j} err {d[j].item():.3e} | row abs max {row.abs().max().item():.3e}")
# and transposed hypothesis: lane == particle
GT_ref = G_ref.T
for lane in [0, 1, 16, 32]:
    row = G[lane]
    d = (GT_ref - row[None, :]).abs().max(1).values
    j = int(d.argmin())
    print(f"[T] lane {lane}: best ref particle {j} err {d[j].item():.3e}")
