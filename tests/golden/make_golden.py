"""Generate the golden fixtures from the REAL reference (tillahoffmann/mininf at /root/reference).

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

For every configuration the unmodified ``mininf.nn.EvidenceLowerBoundLoss`` (mininf/nn.py:212-228)
is evaluated S times through the reference's own ``condition`` / ``LogProbTracer`` path on fixed
data and fixed reparameterisation noise; loss and gradients (w.r.t. the constrained parameters of
the approximation) are averaged over the S evaluations and stored. The approximation object
handed to the reference implements ``rsample`` with the stored noise using exactly the value and
backward functions of torch's own ``rsample`` (oracle/elbo.py); a separate check below confirms
that, for the same generator state, it reproduces ``FactorizedDistribution.rsample`` bit for bit.
"""
import importlib.util
import pathlib
import sys
import warnings

import numpy as np
import torch

ROOT = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import configs, elbo  # noqa: E402

warnings.filterwarnings("ignore", message=".*MaskedTensors is in prototype.*")


def load_reference():
    spec = importlib.util.spec_from_file_location(
        "mininf_ref", "/root/reference/mininf/__init__.py",
        submodule_search_locations=["/root/reference/mininf"])
    module = importlib.util.module_from_spec(spec)
    sys.modules["mininf_ref"] = module
    spec.loader.exec_module(module)
    return module


class FixedNoise:
    """Approximation handed to the reference. Deliberately NOT a dict: the reference re-wraps
    dictionaries in its own FactorizedDistribution (mininf/nn.py:215-216) and would draw fresh
    noise; any other object with ``rsample`` / ``entropy`` is used as is (mininf/nn.py:217,226)."""
    def __init__(self, dists, noise):
        self.dists = dists
        self.noise = noise
        self.particle = 0

    def rsample(self, sample_shape=None):
        return {name: elbo.rsample_given(dist, self.noise[name][self.particle])
                for name, dist in self.dists.items()}

    def entropy(self):
        return sum(dist.entropy().sum() for dist in self.dists.values())


def reference_loss(ref, cfg, approx, noise, n_particles):
    conditioned = ref.condition(lambda: cfg.model(ref), **cfg.data)
    loss_module = ref.nn.EvidenceLowerBoundLoss()
    wrapped = FixedNoise(approx, noise)
    total = 0.0
    for s in range(n_particles):
        wrapped.particle = s
        total = total + loss_module(conditioned, wrapped)
    return total / n_particles


CASES = {
    "coin": (lambda: configs.coin(), 8),
    "regression": (lambda: configs.regression(512, 64), 4),
    "regression_sigma": (lambda: configs.regression(384, 64, sigma_latent=True), 4),
    "regression_ragged": (lambda: configs.regression(301, 24, sigma_latent=True), 3),
    "logistic": (lambda: configs.logistic(20000, 400, p=32), 2),
    "logistic_wide": (lambda: configs.logistic(50000, 333, p=128, intercept=True), 3),
    "missing": (lambda: configs.missing(600), 4),
    "features": (lambda: configs.feature_uncertainty(300, 32), 2),
    # beyond BASELINE.json (tests/conftest.py::EXTRA_GOLDEN_CASES)
    "affine_links_small": (lambda: configs.affine_links(100), 3),
    "affine_links": (lambda: configs.affine_links(5000), 8),
    "feature_example": (lambda: configs.feature_example(30), 4),
    "several_covariates": (lambda: configs.several_covariates(4000), 5),
}


def main():
    ref = load_reference()
    out_dir = pathlib.Path(__file__).resolve().parent
    for case, (build, n_particles) in CASES.items():
        cfg = build()
        torch.manual_seed(1234)
        approx, leaves = cfg.approximation()
        # move the parameters off their initial values so no gradient is accidentally zero
        with torch.no_grad():
            for key, leaf in leaves.items():
                leaf.mul_(1.0 + 0.3 * torch.rand_like(leaf)).add_(0.05 * torch.randn_like(leaf) if key.endswith("loc") else 0.0)
        approx, leaves = cfg.approximation.__func__(
            type("C", (), {"families": {name: (cls, {k: leaves[f"{name}.{k}"].detach() for k in params})
                                        for name, (cls, params) in cfg.families.items()}})())
        noise = {name: elbo.draw_noise(dist, n_particles) for name, dist in approx.items()}
        loss = reference_loss(ref, cfg, approx, noise, n_particles)
        loss.backward()
        # the oracle on the same inputs, for the record (tests re-check this on every run)
        approx_o, leaves_o = cfg.approximation.__func__(
            type("C", (), {"families": {name: (cls, {k: leaves[f"{name}.{k}"].detach() for k in params})
                                        for name, (cls, params) in cfg.families.items()}})())
        loss_o = elbo.neg_elbo(cfg.model, cfg.data, approx_o, noise, n_particles)
        loss_o.backward()
        assert torch.allclose(loss, loss_o, rtol=1e-6), (case, loss, loss_o)
        arrays = {"loss": loss.detach().numpy(), "n_particles": np.array(n_particles)}
        for key, leaf in leaves.items():
            arrays[f"param/{key}"] = leaf.detach().numpy()
            arrays[f"grad/{key}"] = leaf.grad.numpy()
            assert torch.allclose(leaf.grad, leaves_o[key].grad, rtol=1e-5, atol=1e-6), (case, key)
        for name, value in noise.items():
            arrays[f"noise/{name}"] = value.numpy()
        # checksums of the regenerated data guard against generator drift
        for name, value in cfg.data.items():
            dense = value.get_data() if isinstance(value, torch.masked.MaskedTensor) else value
            arrays[f"datasum/{name}"] = dense.double().sum().numpy()
        np.savez_compressed(out_dir / f"{case}.npz", **arrays)
        print(f"{case:20s} S={n_particles} loss={float(loss):.6f} (oracle {float(loss_o):.6f})")

    # the replayed-noise approximation reproduces the reference's own rsample stream
    cfg = configs.regression(64, 8, sigma_latent=True)
    approx, _ = cfg.approximation()
    torch.manual_seed(7)
    mine = {name: elbo.rsample_given(dist, elbo.draw_noise(dist, 1)[0]) for name, dist in approx.items()}
    torch.manual_seed(7)
    theirs = ref.nn.FactorizedDistribution(approx).rsample()
    for name in approx:
        assert torch.equal(mine[name], theirs[name]), name
    # and a seeded evaluation of the reference with its own rsample, as a smoke value
    torch.manual_seed(11)
    loss_ref = ref.nn.EvidenceLowerBoundLoss()(ref.condition(lambda: cfg.model(ref), **cfg.data), approx)
    torch.manual_seed(11)
    loss_orc = elbo.neg_elbo(cfg.model, cfg.data, approx, None, 1)
    assert torch.equal(loss_ref, loss_orc), (loss_ref, loss_orc)
    np.savez_compressed(out_dir / "seeded_stream.npz", loss=loss_ref.detach().numpy())
    print("stream check: rsample replay bit-exact, seeded loss", float(loss_ref))


if __name__ == "__main__":
    main()
