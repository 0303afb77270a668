"""Dump the lowered site tables (pointer fields reduced to present / absent) of the BASELINE model shapes as JSON.

Used to show that a change of the host code leaves the tables the GPU-validated kernels read unchanged:

    git archive <validated commit> mininf_b200 oracle include tests/conftest.py | tar -x -C /tmp/old
    cp -r mininf_b200/_lib /tmp/old/mininf_b200/
    python tools/dump_tables.py /tmp/old > old.json; python tools/dump_tables.py . > new.json; cmp old.json new.json

End of round 2: identical for 12 configurations x {black-box, closed form} between the last commit whose
kernels and host code ran on a B200 (0d5da7f) and HEAD (no GPU time was left to re-run the parity suite).
"""
import sys, json, ctypes as C, warnings
warnings.filterwarnings("ignore")
root = sys.argv[1]
sys.path.insert(0, root)
import torch
import mininf_b200 as mininf
from mininf_b200.engine import abi
from mininf_b200.engine.plan import Plan, assign_offsets, latent_parameters
from mininf_b200.engine.trace import Affine, LatentRef, LinkTensor, SiteTableTracer
from oracle import configs
CPU = torch.device("cpu")

def norm(struct):
    out = {}
    for name, ctype in struct._fields_:
        v = getattr(struct, name)
        if ctype is C.c_void_p:
            out[name] = int(bool(v))
        elif isinstance(v, C.Structure):
            out[name] = norm(v)
        elif isinstance(v, C.Array):
            out[name] = [norm(e) if isinstance(e, C.Structure) else e for e in v]
        else:
            out[name] = v
    return out

def build(config, S, closed_form=False, mode="auto"):
    torch.manual_seed(0)
    approx, _ = config.approximation()
    entries, draws = [], {}
    for name, factor in approx.items():
        family, p0, _ = latent_parameters(factor)
        shape = factor.batch_shape
        entries.append((name, family, shape))
        ref = LatentRef(name, 0) if max(shape.numel(), 1) == 1 else LatentRef(name)
        draws[name] = LinkTensor.wrap(factor.sample(), Affine(a_lat=ref))
    model = mininf.condition(lambda: config.model(mininf), **config.data)
    with SiteTableTracer() as tracer:
        mininf.condition(model, **draws)()
    try:
        from mininf_b200.engine.plan import row_latent_names, slope_groups
        specs = assign_offsets(entries, row_latent_names(tracer.sites), slope_groups(tracer.sites))
    except ImportError:
        specs = assign_offsets(entries)
    plan = Plan(tracer.sites, specs, S, CPU, dense_mode=mode, dry_run=True, closed_form=closed_form)
    leaves = []
    for v in config.data.values():
        if isinstance(v, torch.masked.MaskedTensor):
            leaves += [v.get_data(), v.get_mask()]
        else:
            leaves.append(v)
    plan.bind_sources(leaves)
    return {"D": plan.D, "rebindable": plan.rebindable,
            "specs": [(s.name, s.family, s.numel, s.offset, s.row_latent) for s in specs],
            "dense": [(norm(site), mode) for site, mode in plan.dense_sites],
            "groups": [[norm(g[i]) for i in range(len(g))] for g in plan.sweep_groups],
            "small_observed": [norm(s) for s in plan._small_observed_host],
            "small_global": [norm(s) for s in plan._small_global_host],
            "rows": {k: norm(v) for k, v in plan.row_groups.items()}}

cases = {
    "coin": (configs.coin(), 8), "regression": (configs.regression(512, 64), 4),
    "regression_sigma": (configs.regression(384, 64, sigma_latent=True), 64),
    "regression_ragged": (configs.regression(301, 24, sigma_latent=True), 3),
    "regression_22": (configs.regression(1000, 22), 8),
    "logistic": (configs.logistic(20000, 400, p=32), 2),
    "logistic_wide": (configs.logistic(50000, 333, p=128, intercept=True), 3),
    "logistic_256": (configs.logistic(100000, 5000, p=256), 16),
    "missing_small": (configs.missing(600), 4), "missing_big": (configs.missing(5000), 64),
    "features": (configs.feature_uncertainty(300, 32), 2), "features_p1": (configs.feature_uncertainty(9000, 1), 3),
}
out = {}
for name, (config, S) in cases.items():
    for cf in (False, True):
        out[f"{name}/{cf}"] = build(config, S, closed_form=cf)
print(json.dumps(out, sort_keys=True, default=str))
