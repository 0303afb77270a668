"""CPU simulation (numpy, float64 reference) of the operand-precision schemes of dense_th at the bench's
operating point: how far is the per-particle Normal log-likelihood and its theta-gradient from float64 when

  single : theta rounded once to fp16 (round 1's scheme, with TF32's equal 11-bit significand)
  hilo   : theta = hi + lo fp16 rows (the committed kernel)
  tiles  : ONE fp16 row, re-rounded per 128-row tile: element e goes to the fp16 neighbour away from zero iff its
           residual fraction exceeds the van der Corput value of the tile index (tools/experiments/README.md)

X is rounded to fp16 in all three (as in the kernel); products exact, sums in float64 (the accumulation
order of the tensor core is not modelled). Usage: python tools/experiments/rerounded_theta_sim.py [N]"""
import sys

import numpy as np

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
p, S, TILE = 64, 16, 128
rng = np.random.default_rng(0)
X = rng.standard_normal((N, p)).astype(np.float32)
theta_true = rng.standard_normal(p) / np.sqrt(p)
y = X @ theta_true + rng.standard_normal(N)
theta = (0.1 * rng.standard_normal((S, p)) + theta_true).astype(np.float32)        # particles around the optimum
X16 = X.astype(np.float16).astype(np.float64)
up = 2.0 ** (11 - np.ceil(np.log2(np.abs(theta).max())))                           # largest |theta| 2^k in [2^10, 2^11)


def stats(eta):          # Normal(., 1): log-likelihood without constants and the theta-gradient, per particle
    r = y[None, :] - eta
    return -0.5 * (r * r).sum(1), r @ X16


exact_ll, exact_g = stats(theta.astype(np.float64) @ X16.T)


def report(name, eta):
    ll, g = stats(eta)
    print(f"{name:7s} per-particle log-lik max rel {np.abs(ll / exact_ll - 1).max():.2e} | loss rel "
          f"{abs(ll.mean() / exact_ll.mean() - 1):.2e} | grad rel-l2 {np.linalg.norm(g - exact_g) / np.linalg.norm(exact_g):.2e}")


t = theta.astype(np.float64) * up
single = t.astype(np.float16).astype(np.float64)
report("single", (single / up) @ X16.T)
lo = (t - single).astype(np.float16).astype(np.float64)
report("hilo", ((single + lo) / up) @ X16.T)

# neighbours of t towards / away from zero in fp16, residual fraction
h = t.astype(np.float16)
toward = np.where(np.abs(h.astype(np.float64)) > np.abs(t), np.nextafter(h, np.float16(0)), h).astype(np.float16)
away = np.nextafter(toward, np.where(t >= 0, np.float16(np.inf), np.float16(-np.inf)).astype(np.float16))
dn, upv = toward.astype(np.float64), away.astype(np.float64)
frac = np.where(upv != dn, (t - dn) / np.where(upv != dn, upv - dn, 1.0), 0.0)
eta = np.empty((S, N))
for tile in range((N + TILE - 1) // TILE):
    c = int(f"{tile:032b}"[::-1], 2) / 2.0 ** 32                                    # van der Corput
    th = np.where(frac > c, upv, dn) / up
    rows = slice(tile * TILE, min(N, (tile + 1) * TILE))
    eta[:, rows] = th @ X16[rows].T
report("tiles", eta)
