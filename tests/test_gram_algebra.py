"""The closed forms that csrc/dense_gram.cuh::gram_finish_kernel evaluates, restated in numpy and
checked against the direct per-particle sums that the reference computes (Normal.log_prob over
X @ theta, TORCH normal.py:87-103, and its gradients): the expansion of the Normal likelihood around
the particle mean in the data-only statistics A = X'X, c = X'r0, sx = X'1, R1, R2, n."""
import numpy as np
import pytest


def direct(X, y, theta, a, sigma):
    r = y[None, :] - a[:, None] - theta @ X.T                       # [S, n] residuals
    n = X.shape[0]
    logp = -0.5 * (r ** 2).sum(1) / sigma ** 2 - n * (np.log(sigma) + 0.5 * np.log(2 * np.pi))
    return logp, (r @ X) / sigma[:, None] ** 2, r.sum(1) / sigma ** 2, ((r ** 2).sum(1) / sigma ** 2 - n) / sigma


def through_statistics(X, y, theta, a, sigma, live=None):
    if live is not None:                                            # masked rows leave every sum
        X, y = X[live], y[live]
    theta0, a0 = theta.mean(0), a.mean()
    r0 = y - a0 - X @ theta0
    A, c, sx, R1, R2, n = X.T @ X, X.T @ r0, X.sum(0), r0.sum(), (r0 ** 2).sum(), float(X.shape[0])
    delta, alpha = theta - theta0, a - a0
    cj = c[None, :] - alpha[:, None] * sx[None, :]
    rj = cj - delta @ A
    Q = R2 - 2 * alpha * R1 + n * alpha ** 2 - (delta * cj).sum(1) - (delta * rj).sum(1)
    logp = -0.5 * Q / sigma ** 2 - n * (np.log(sigma) + 0.5 * np.log(2 * np.pi))
    return logp, rj / sigma[:, None] ** 2, (R1 - n * alpha - delta @ sx) / sigma ** 2, (Q / sigma ** 2 - n) / sigma


@pytest.mark.parametrize("n,p,S,masked", [(1, 4, 1, False), (257, 64, 7, False), (1000, 32, 64, True)])
def test_closed_forms_equal_the_direct_sums(n, p, S, masked):
    rng = np.random.default_rng(n + p)
    X = rng.normal(size=(n, p)) + rng.uniform(0, 0.5, size=p)
    theta_true = rng.normal(size=p)
    y = X @ theta_true + 0.7 + 0.1 * rng.normal(size=n)             # signal / residual ratio of a few thousand
    theta = theta_true + 0.05 * rng.normal(size=(S, p))
    a = 0.7 + 0.05 * rng.normal(size=S)
    sigma = np.exp(0.1 * rng.normal(size=S))
    live = rng.uniform(size=n) < 0.7 if masked else None
    if live is not None:
        expected = direct(X[live], y[live], theta, a, sigma)
    else:
        expected = direct(X, y, theta, a, sigma)
    got = through_statistics(X, y, theta, a, sigma, live)
    for e, g in zip(expected, got):
        np.testing.assert_allclose(g, e, rtol=1e-9, atol=1e-9 * np.abs(e).max())
