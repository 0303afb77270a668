import pathlib
import sys
import warnings

import numpy as np
import pytest
import torch

ROOT = pathlib.Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = pathlib.Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    warnings.filterwarnings("ignore", message=".*MaskedTensors is in prototype.*")
    warnings.filterwarnings("ignore", message=".*Converting a tensor with requires_grad=True to a scalar.*")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


GOLDEN_CASES = {
    "coin": lambda cfgs, **kw: cfgs.coin(**{k: v for k, v in kw.items() if k == "device"}),
    "regression": lambda cfgs, **kw: cfgs.regression(512, 64, **kw),
    "regression_sigma": lambda cfgs, **kw: cfgs.regression(384, 64, sigma_latent=True, **kw),
    "regression_ragged": lambda cfgs, **kw: cfgs.regression(301, 24, sigma_latent=True, **kw),
    "logistic": lambda cfgs, **kw: cfgs.logistic(20000, 400, p=32, **kw),
    "logistic_wide": lambda cfgs, **kw: cfgs.logistic(50000, 333, p=128, intercept=True, **kw),
    "missing": lambda cfgs, **kw: cfgs.missing(600, **kw),
    "features": lambda cfgs, **kw: cfgs.feature_uncertainty(300, 32, **kw),
}


# not BASELINE configurations: links that reduce to the affine form, and the feature-uncertainty example
# as the reference writes it. Held by the CPU oracle test and by tests/test_widened_links_gpu.py only.
EXTRA_GOLDEN_CASES = {
    "affine_links_small": lambda cfgs, **kw: cfgs.affine_links(100, **kw),
    "affine_links": lambda cfgs, **kw: cfgs.affine_links(5000, **kw),
    "feature_example": lambda cfgs, **kw: cfgs.feature_example(30, **kw),
    "several_covariates": lambda cfgs, **kw: cfgs.several_covariates(4000, **kw),
}


def load_golden(case, device="cpu"):
    """(config on ``device`` with CPU-generated data, golden arrays)."""
    from oracle import configs
    golden = dict(np.load(GOLDEN / f"{case}.npz"))
    kwargs = {"device": device}
    if case != "coin":
        kwargs["gen_device"] = "cpu"
    config = {**GOLDEN_CASES, **EXTRA_GOLDEN_CASES}[case](configs, **kwargs)
    # install the golden parameter values as the approximation's initial values
    families = {}
    for name, (cls, params) in config.families.items():
        families[name] = (cls, {key: torch.from_numpy(golden[f"param/{name}.{key}"]) for key in params})
    config.families = families
    for name, value in config.data.items():
        dense = value.get_data() if isinstance(value, torch.masked.MaskedTensor) else value
        np.testing.assert_allclose(float(dense.double().sum()), float(golden[f"datasum/{name}"]), rtol=1e-9,
                                   err_msg="regenerated data differs from the data the golden was made with")
    return config, golden


def golden_noise(config, golden, device="cpu"):
    return {name: torch.from_numpy(golden[f"noise/{name}"]).to(device) for name in config.families}
