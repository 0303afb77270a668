"""Benchmark of the ELBO + gradient hot path (BASELINE.json metric) on synthetic data.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c2|c3|c4|c5]
                    [--closed-form] [--weak] [--eager]

Default workload c2 (config[1] of BASELINE.json): Bayesian linear regression, mean-field Normal
approximation, p = 64 features, N = 1e8 observations IN TOTAL, S = 64 particles, evaluated by the
BLACK-BOX sweep: one log-density + score evaluation per (particle, observation) on the tcgen05
kernel of csrc/dense_tc.cuh. `--closed-form` lets Normal sites use their data-only sufficient
statistics instead (Gram matrix / six sums / Chebyshev moments; no per-particle work) - reported
as a clearly named secondary object in the default run.
Under torchrun (N ranks) the default is STRONG scaling: the global data set is the same at every
world size (256 row chunks, chunk c seeded seed0 + c) and rank r owns chunks
[r*256/N, (r+1)*256/N); `--weak` gives every rank the full configuration instead.
`--workload c3` (config[2]): minibatch logistic regression, p = 256, batches of 1e7 rows of a
declared N = 1e9 stream, S = 16; two resident batches alternate (one recorded step each).
`--workload c4` (config[3]): regression with feature uncertainty, N = 1e7 rows with a
per-observation latent feature vector (p = 32), S = 32 (the row-latent kernel; instruction-bound).
`--workload c5` (config[4]): masked Poisson + Normal sites over N = 1e8 elements, 30 % missing,
S = 64 (site sweeps; closed-form statistics by default for this workload, black-box beside it).
A step is one full SVI step: parameter transforms, reparameterised draws, ELBO + gradient kernels,
chain rule, Adam - `mininf_b200.nn.FusedSVIStep` (one native call, replayed from a CUDA graph);
per-observation latents (c4) are trained inside the row-latent sweep. `--torch-adam` runs
`GraphedStep` with torch's fused Adam instead. The metric is
particle-observation log-density evaluations per second: rows * S / time, whole job.

`--impl reference` times the UNMODIFIED reference (baseline/_ref, installed from /root/reference)
on the host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "elbo_grad_particle_obs_evals_per_sec"
UNIT = "evals/s"
SEED0 = 2000
N_CHUNKS = 256     # synthetic rows are generated in 256 chunks, chunk c seeded seed0 + c


class Workload:
    def __init__(self, key, name, p, particles, rows, family, declared_rows, n_batches, kernel, traffic_file,
                 event_kind="dense", cpu_sample=(4_000_000, 4), bytes_per_row=None, bound_note=None,
                 closed_form_kernel=None, closed_form_default=False, black_box_kernel=None):
        self.key, self.name, self.p, self.particles, self.rows = key, name, p, particles, rows
        self.family, self.declared_rows, self.n_batches = family, declared_rows, n_batches
        self.kernel, self.traffic_file = kernel, traffic_file
        self.event_kind, self.cpu_rows, self.cpu_particles = event_kind, cpu_sample[0], cpu_sample[1]
        # algorithmic bytes one sweep call moves per row
        self.bytes_per_row = bytes_per_row if bytes_per_row is not None else 4 * p + 4
        self.bound_note = bound_note
        # Normal / Poisson sites have closed forms in data-only statistics; which path is the headline
        self.closed_form_kernel, self.closed_form_default = closed_form_kernel, closed_form_default
        self.black_box_kernel = black_box_kernel


WORKLOADS = {
    "c2": Workload("c2", "bayesian_linear_regression_p64_N1e8_S64", 64, 64, 100_000_000, "normal", None, 1,
                   "mnf::th::dense_th_kernel<Normal> (fp16 operands; + its partial-sum reduction, ~8 us)",
                   "dense_th_traffic.json",
                   closed_form_kernel="mnf::gram::dense_gram_kernel (+ gram_reduce / gram_finish / reduce_partials)"),
    "c3": Workload("c3", "minibatch_logistic_regression_p256_batch1e7_of_N1e9_S16", 256, 16, 10_000_000,
                   "bernoulli", 1_000_000_000, 2,
                   "mnf::tcr::dense_tcr_kernel<BernoulliLogits, 16> (+ its partial-sum reduction)",
                   "dense_tcr_traffic.json", cpu_sample=(1_000_000, 4)),
    "c4": Workload("c4", "feature_uncertainty_rowlatent_p32_N1e7_S32", 32, 32, 10_000_000, "rowlatent", None, 1,
                   "mnf::rowlatent_kernel<32> (+ its partial-sum reduction)", "rowlatent_traffic.json",
                   event_kind="rowlatent", cpu_sample=(500_000, 4), bytes_per_row=32 * 20 + 4,
                   bound_note="instruction-issue bound (N*p*S = 1.0e10 Philox normal draws per step), not HBM; "
                              "see profiles/ for the pipe utilisation"),
    "c5": Workload("c5", "missing_observations_poisson_normal_N1e8_S64_30pct_masked", 1, 64, 100_000_000, "missing",
                   None, 1, "mnf::poisson_range_kernel + mnf::poisson_moment_kernel + mnf::normal_stats_kernel "
                            "(one mnf_site_sweep call)",
                   "site_sweep_traffic.json", event_kind="site", cpu_sample=(10_000_000, 4), bytes_per_row=14,
                   closed_form_default=True,
                   black_box_kernel="mnf::poisson_exp_kernel (one ex2 per live element and particle) + "
                                    "mnf::site_sweep_kernel (Normal site, per particle)",
                   bound_note="both sites are reduced to data-only sufficient statistics (33 Chebyshev moments of "
                              "the covariate for the Poisson site, six sums for the Normal site): three streaming "
                              "passes that move 23 B per element for 14 algorithmic bytes (the covariate is read by "
                              "each pass); the moment pass is bound by fp32 issue (64 FFMA/FADD per element), the "
                              "other two by HBM"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--rows", type=float, default=float(os.environ.get("MNF_BENCH_ROWS", 0)),
                    help="observations (c3: batch rows) in total; default: the BASELINE configuration")
    ap.add_argument("--eager", action="store_true", help="do not replay the step from a CUDA graph")
    ap.add_argument("--closed-form", action="store_true",
                    help="let sites with closed-form sufficient statistics skip the per-particle sweep")
    ap.add_argument("--black-box", action="store_true", help="force the per-particle sweep (c5 defaults to closed form)")
    ap.add_argument("--weak", action="store_true", help="weak scaling: the full configuration on every rank")
    ap.add_argument("--torch-adam", action="store_true",
                    help="GraphedStep with torch.optim.Adam(fused=True) instead of the native FusedSVIStep")
    ap.add_argument("--precision", default="auto", choices=["auto", "f16", "tf32", "fp32"],
                    help="tensor-core operand format of dense sites (auto: fp16 operands for p = 64 / S <= 64, else TF32)")
    ap.add_argument("--reduce", default="peer", choices=["peer", "nccl"], help="how sharded ranks combine partial sums")
    ap.add_argument("--sustain", type=float, default=1.0, help="seconds of back-to-back steps for the sustained figure")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary objects (closed form, reference on CUDA)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def dense_dtype(plan):
    """The arithmetic type of the dense sweep that ran (engine/abi.py DENSE_* modes)."""
    from mininf_b200.engine import abi
    modes = {mode for _, mode in plan.dense_sites}
    if modes == {abi.DENSE_F16}:
        return "f16 operands (11-bit significand; theta as hi + lo pairs) / f32 accumulate"
    if modes <= {abi.DENSE_TF32, abi.DENSE_TF32_CLOSED_FORM}:
        return "tf32 operands (theta as hi + lo pairs) / f32 accumulate"
    return "f32"


def dense_kernel_name(w, plan):
    from mininf_b200.engine import abi
    if w.key == "c2" and {mode for _, mode in plan.dense_sites} == {abi.DENSE_TF32}:
        return "mnf::tc::dense_tc_kernel<Normal> (TF32 operands; + its partial-sum reduction)"
    return w.kernel


def measured_peak_gbs():
    path = ROOT / "MEASURED_PEAKS.json"
    if path.exists():
        return float(json.loads(path.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------
# synthetic data: 256 row chunks, chunk c seeded seed0 + c
# ---------------------------------------------------------------------------------------------
def chunk_bounds(n, chunk):
    per = -(-n // N_CHUNKS)
    return min(chunk * per, n), min((chunk + 1) * per, n)


def rank_rows(n, rank, world):
    """Rows owned by `rank`: chunks [rank*256/W, (rank+1)*256/W) of the global data set."""
    lo, _ = chunk_bounds(n, rank * N_CHUNKS // world)
    _, hi = chunk_bounds(n, (rank + 1) * N_CHUNKS // world - 1)
    return lo, hi


def make_data(w, n, device, seed0, rows=None):
    """Rows [rows[0], rows[1]) (default: all) of the global data set of `n` rows in the chunk-seeded
    recipe: chunk c is drawn from seed0 + c, so the data are identical at every world size."""
    g = torch.Generator(device=device)
    g.manual_seed(SEED0 - 1)
    first, last = rows if rows is not None else (0, n)
    k_local = last - first
    chunks = [(chunk, *chunk_bounds(n, chunk)) for chunk in range(N_CHUNKS)]
    chunks = [(chunk, lo, hi) for chunk, lo, hi in chunks if hi > lo and lo >= first and hi <= last]
    assert sum(hi - lo for _, lo, hi in chunks) == k_local, "rank rows must be whole chunks"
    if w.family == "missing":
        x = torch.empty(k_local, device=device)
        counts = torch.empty(k_local, device=device)
        wv = torch.empty(k_local, device=device)
        m_counts = torch.empty(k_local, device=device, dtype=torch.bool)
        m_w = torch.empty(k_local, device=device, dtype=torch.bool)
        for chunk, lo, hi in chunks:
            lo, hi = lo - first, hi - first
            g.manual_seed(seed0 + chunk)
            k = hi - lo
            torch.randn(k, generator=g, device=device, out=x[lo:hi])
            counts[lo:hi] = torch.poisson(torch.exp(0.3 + 0.5 * x[lo:hi]), generator=g)
            wv[lo:hi] = -0.2 + 0.8 * x[lo:hi] + 0.7 * torch.randn(k, generator=g, device=device)
            m_counts[lo:hi] = torch.rand(k, generator=g, device=device) > 0.3
            m_w[lo:hi] = torch.rand(k, generator=g, device=device) > 0.3
        # finite fill in the holes, as in examples/missing-observations.md
        counts.mul_(m_counts)
        wv.mul_(m_w)
        return {"x": x, "counts": counts, "w": wv, "m_counts": m_counts, "m_w": m_w}
    theta_true = torch.randn(w.p, generator=g, device=device) / w.p ** 0.5
    X = torch.empty(k_local, w.p, device=device)
    y = torch.empty(k_local, device=device)
    for chunk, lo, hi in chunks:
        lo, hi = lo - first, hi - first
        g.manual_seed(seed0 + chunk)
        torch.randn(hi - lo, w.p, generator=g, device=device, out=X[lo:hi])
        if w.family == "normal":
            torch.randn(hi - lo, generator=g, device=device, out=y[lo:hi])
            y[lo:hi].addmv_(X[lo:hi], theta_true)
        elif w.family == "bernoulli":
            y[lo:hi] = torch.bernoulli(torch.sigmoid(X[lo:hi] @ theta_true), generator=g)
        else:   # rowlatent: X holds the noisy features x = z + 0.5 eps, y ~ Poisson(exp(0.5 + z . slope))
            y[lo:hi] = torch.poisson(torch.exp(0.5 + X[lo:hi] @ theta_true), generator=g)
            X[lo:hi].add_(torch.randn(hi - lo, w.p, generator=g, device=device), alpha=0.5)
    return {"X": X, "y": y}


def model_factory(m, w, n_rows, data, declared=None):
    """The workload's model over `n_rows` local rows; `declared` is the population size behind a
    minibatch (c3; a rank of a sharded run declares its share of it)."""
    from torch.distributions import Bernoulli, Gamma, Normal, Poisson
    declared = declared or w.declared_rows or n_rows

    def regression():
        theta = m.sample("theta", Normal(0, 1), w.p)
        with m.no_log_prob():
            X = m.sample("X", Normal(0, 1), (n_rows, w.p))
        m.sample("y", Normal(X @ theta, 1.0))

    def minibatch_logistic():          # examples/minibatch.md:26-35 with a Bernoulli response
        theta = m.sample("theta", Normal(0, 1), w.p)
        with m.batch(declared):
            with m.no_log_prob():
                X = m.sample("X", Normal(0, 1), (declared, w.p))
            m.sample("y", Bernoulli(logits=X @ theta))

    def feature_uncertainty():         # examples/regression-with-feature-uncertainty.md:28-38, p features
        population_scale = m.sample("population_scale", Gamma(2, 2))
        z = m.sample("z", Normal(0, population_scale), (n_rows, w.p))
        m.sample("X", Normal(z, 0.5))
        intercept = m.sample("intercept", Normal(0, 1))
        slope = m.sample("slope", Normal(0, 1), w.p)
        m.sample("y", Poisson((intercept + z @ slope).exp()))

    def missing_observations():        # examples/missing-observations.md:131 (masked conditioning)
        x = data["x"]
        a = m.sample("a", Normal(0, 1))
        b = m.sample("b", Normal(0, 1))
        c = m.sample("c", Normal(0, 1))
        d = m.sample("d", Normal(0, 1))
        sigma = m.sample("sigma", Gamma(2, 2))
        m.sample("counts", Poisson((a + b * x).exp()))
        m.sample("w", Normal(c + d * x, sigma))

    return {"normal": regression, "bernoulli": minibatch_logistic, "rowlatent": feature_uncertainty,
            "missing": missing_observations}[w.family]


def condition_on(m, model, w, data):
    """`mininf.condition` of the workload's model on one resident data set."""
    if w.family == "missing":
        return m.condition(model, counts=torch.masked.as_masked_tensor(data["counts"], data["m_counts"]),
                           w=torch.masked.as_masked_tensor(data["w"], data["m_w"]))
    return m.condition(model, X=data["X"], y=data["y"])


def make_approximation(m, w, n_rows, data, device, validate=False):
    """name -> ParameterizedDistribution. With `validate=False`, `validate_args=False` is passed
    through to torch.distributions exactly as in the reference (its constructor check is a host
    round trip per step; the kernels check scales on the device); the reference arm keeps the
    shipped default (validation on)."""
    from torch.distributions import Gamma, Normal
    extra = {} if validate else {"validate_args": False}

    def PD(cls, **parameters):
        return m.nn.ParameterizedDistribution(cls, **parameters, **extra)

    def scalar(value):
        return torch.tensor(value, device=device)

    if w.family in ("normal", "bernoulli"):
        return {"theta": PD(Normal, loc=torch.zeros(w.p, device=device), scale=0.1 * torch.ones(w.p, device=device))}
    if w.family == "rowlatent":
        return {"population_scale": PD(Gamma, concentration=scalar(2.0), rate=scalar(2.0)),
                "z": PD(Normal, loc=data["X"].clone(), scale=torch.ones(n_rows, w.p, device=device)),
                "intercept": PD(Normal, loc=scalar(0.1), scale=scalar(0.2)),
                "slope": PD(Normal, loc=torch.zeros(w.p, device=device), scale=0.2 * torch.ones(w.p, device=device))}
    approximation = {k: PD(Normal, loc=scalar(0.1), scale=scalar(0.2)) for k in "abcd"}
    approximation["sigma"] = PD(Gamma, concentration=scalar(2.0), rate=scalar(2.0))
    return approximation


# ---------------------------------------------------------------------------------------------
# clocks: sample nvidia-smi during the timed region
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """`nvidia-smi -lms` in the background from before the warm-up; only samples whose timestamp
    falls inside the timed region [t0, t1] count (B200_PROFILING.md clocks line)."""
    QUERY = ("timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.file = None

    def start(self):
        try:
            self.file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "10"], stdout=self.file, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self, t0, t1):
        import datetime
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        self.proc.wait()
        self.file.flush()
        rows = [line.strip().split(", ") for line in open(self.file.name) if line.strip()]
        os.unlink(self.file.name)
        inside, everything, reasons, sm_max = [], [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in rows:
            try:
                stamp = datetime.datetime.strptime(row[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                clock = float(row[1])
                sm_max = float(row[2])
            except (ValueError, IndexError):
                continue
            everything.append(clock)
            if t0 - 0.005 <= stamp <= t1 + 0.005:
                inside.append(clock)
                for name, flag in zip(names, row[3:7]):
                    if flag.strip().lower().startswith("active"):
                        reasons.add(name)
        chosen = sorted(inside) if inside else sorted(everything[-5:])
        median = chosen[len(chosen) // 2] if chosen else None
        return {"sm_mhz": median, "sm_max_mhz": sm_max, "reasons": sorted(reasons),
                "samples_in_timed_region": len(inside), "samples_total": len(everything)}


# ---------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the UNMODIFIED reference (baseline/_ref) on the host cores
# ---------------------------------------------------------------------------------------------
REF_DIR = ROOT / "baseline" / "_ref"


def load_reference():
    """The unmodified tillahoffmann/mininf package installed under baseline/_ref (git-ignored; it
    travels to the GPU box with the snapshot). Returns None when it is not there."""
    if not (REF_DIR / "mininf" / "__init__.py").exists():
        return None
    if str(REF_DIR) not in sys.path:
        sys.path.insert(0, str(REF_DIR))
    import importlib
    return importlib.import_module("mininf")


def reference_sample(w, rows, particles, steps, warmup, device="cpu", validate=True):
    """Time the reference's own SVI step (README.md:66-69) on a bounded sample: `rows` observations,
    `particles` sequential `mininf.nn.EvidenceLowerBoundLoss` calls per step (the reference draws
    one particle per call, mininf/nn.py:217), validation on as shipped. Falls back to the oracle
    port when baseline/_ref is missing. Returns (evals/s, seconds per step, kind)."""
    ref = load_reference()
    if ref is None:
        value, seconds = port_sample(w, rows, particles, steps, warmup)
        return value, seconds, "port"
    torch.manual_seed(0)
    device = torch.device(device)
    data = make_data(w, rows, device, SEED0)
    modules = make_approximation(ref, w, rows, data, device, validate=validate)
    parameters = [p for module in modules.values() for p in module.parameters()]
    optimizer = torch.optim.Adam(parameters, lr=0.01)
    loss_module = ref.nn.EvidenceLowerBoundLoss()
    conditioned = condition_on(ref, model_factory(ref, w, rows, data), w, data)

    def step():
        optimizer.zero_grad()
        for _ in range(particles):
            loss = loss_module(conditioned, {name: module() for name, module in modules.items()})
            (loss / particles).backward()
        optimizer.step()

    def sync():
        if device.type == "cuda":
            torch.cuda.synchronize(device)

    for _ in range(warmup):
        step()
    sync()
    begin = time.perf_counter()
    for _ in range(steps):
        step()
    sync()
    seconds = (time.perf_counter() - begin) / steps
    return rows * particles / seconds, seconds, "reference"


def port_sample(w, rows, particles, steps, warmup):
    """The oracle port of the same step (only used when baseline/_ref is absent)."""
    from oracle import configs, elbo
    torch.manual_seed(0)
    if w.family == "normal":
        config = configs.regression(rows, w.p, seed0=SEED0)
    elif w.family == "bernoulli":
        config = configs.logistic(w.declared_rows, rows, p=w.p, seed0=SEED0)
    elif w.family == "rowlatent":
        config = configs.feature_uncertainty(rows, w.p, seed0=SEED0)
    else:
        config = configs.missing(rows, seed0=SEED0)
    approx, leaves = config.approximation()
    optimizer = torch.optim.Adam(list(leaves.values()), lr=0.01)

    def step():
        optimizer.zero_grad()
        loss = elbo.neg_elbo(config.model, config.data, approx, None, particles)
        loss.backward()
        optimizer.step()

    for _ in range(warmup):
        step()
    begin = time.perf_counter()
    for _ in range(steps):
        step()
    seconds = (time.perf_counter() - begin) / steps
    return rows * particles / seconds, seconds


def cpu_sample_text(w, rows, particles, cores, kind):
    what = ("the unmodified reference (baseline/_ref: mininf.nn.EvidenceLowerBoundLoss, one call per particle, "
            "validation on as shipped)" if kind == "reference" else
            "oracle port of mininf's torch.distributions path, validation on")
    return f"{rows} rows x {particles} particles per step of the {w.name} workload; {what}, {cores} threads"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    rows, particles = w.cpu_rows, w.cpu_particles
    # torchrun exports OMP_NUM_THREADS=1 to every rank; rank 0 runs alone here, so give the CPU arm
    # every host thread this process may use
    torch.set_num_threads(max(len(os.sched_getaffinity(0)), 1))
    value, seconds, kind = reference_sample(w, rows, particles, args.steps, args.warmup)
    cores = torch.get_num_threads()
    sample = cpu_sample_text(w, rows, particles, cores, kind)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": seconds * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w.name, "rows_total": w.rows,
                   "features": w.p, "particles": w.particles, "cpu_sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# the B200 arm
# ---------------------------------------------------------------------------------------------
def run_b200(args):
    import torch.distributed as dist
    import mininf_b200 as mininf

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    distributed = world > 1
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if distributed:
        # Keep stdout to the one JSON line: with NCCL_DEBUG set, NCCL printf()s its version banner
        # to stdout while the communicator comes up, so fd 1 points at stderr until it has.
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=device)
            warm = torch.zeros(1, device=device)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    w = WORKLOADS[args.workload]
    P, S = w.p, w.particles
    n_total = int(args.rows) or w.rows                 # observations (c3: batch rows) of the whole job ...
    strong = distributed and not args.weak
    if args.weak:
        n_total *= world                               # ... unless every rank gets the full configuration
    lo, hi = rank_rows(n_total, rank, world)
    n_rows = hi - lo                                   # this rank's rows: whole chunks of the global data set
    closed_form = (args.closed_form or w.closed_form_default) and not args.black_box
    declared = (w.declared_rows * (world if args.weak else 1)) // world if w.declared_rows else None
    # one resident data set, or the resident batches of the stream (c3), 256 seeds apart
    batches = [make_data(w, n_total, device, SEED0 + b * 256, rows=(lo, hi)) for b in range(w.n_batches)]
    modules = make_approximation(mininf, w, n_rows, batches[0], device)
    streaming = w.n_batches > 1
    fused = not args.torch_adam                        # FusedSVIStep (native transforms + Adam) unless asked otherwise
    graphed = not args.eager
    model = model_factory(mininf, w, n_rows, batches[0], declared)

    def make_loss(closed):
        return mininf.nn.EvidenceLowerBoundLoss(S, dense_precision=args.precision, closed_form=closed,
                                                reduce=args.reduce, process_group=True if distributed else None)

    def approximation():
        return {name: module() for name, module in modules.items()}

    def fence():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    # ---- the step -------------------------------------------------------------------------------
    calls = [0]
    loss_modules = [make_loss(closed_form) for _ in batches]
    if fused:
        steps_by_batch = []
        for loss_module, batch in zip(loss_modules, batches):
            steps_by_batch.append(mininf.nn.FusedSVIStep(
                loss_module, condition_on(mininf, model, w, batch), modules, lr=0.01, graph=graphed,
                share_with=steps_by_batch[0] if steps_by_batch else None))
        optimizer = None
    else:
        parameters = [p for module in modules.values() for p in module.parameters()]
        # fused=True: torch's single-kernel Adam (the foreach default issues six passes over every
        # parameter, which shows at C4's 6.4e8 variational parameters)
        optimizer = torch.optim.Adam(parameters, lr=0.01, capturable=graphed, fused=True)
        if graphed:
            steps_by_batch = [mininf.nn.GraphedStep(loss_module, condition_on(mininf, model, w, batch), approximation,
                                                    optimizer) for loss_module, batch in zip(loss_modules, batches)]
        else:
            def eager(loss_module, batch):
                def run():
                    optimizer.zero_grad(set_to_none=True)
                    loss = loss_module(condition_on(mininf, model, w, batch), approximation())
                    loss.backward()
                    optimizer.step()
                    return loss
                return run
            steps_by_batch = [eager(loss_module, batch) for loss_module, batch in zip(loss_modules, batches)]

    def step():
        run = steps_by_batch[calls[0] % len(steps_by_batch)]
        calls[0] += 1
        return run()

    plan = loss_modules[0].last_plan
    step()
    kernels_per_step = plan.gpu_launches_per_step      # counted by the library during the step just enqueued

    # Roofline numerator: average duration of the workload's sweep call (the dominant kernel + its
    # partial-sum reduction, through the phase-level C-ABI with the same arguments as the recorded
    # step), from ONE CUDA-event pair around a back-to-back series of launches on the launching
    # stream. The series keeps the GPU saturated, so it runs in the power state of the timed region
    # (per-launch events in an eager loop leave idle gaps in which the SM clock recovers under the
    # power cap: that read 8 % faster than the same kernel inside the timed loop). The SM clock
    # still drifts over a run, so the series is taken right before AND right after the timed
    # region and both are averaged. Row-latent sweeps (18 ms per launch) keep per-launch events:
    # their parameter / gradient buffers are bound by the loss call.
    def measure_sweep(repeats):
        if plan.row_latents:
            plan.sweep_events.clear()
            plan.sweep_event_kinds.clear()
            plan.record_sweep_events = True
            for i in range(repeats):
                with torch.no_grad():
                    loss_modules[0](condition_on(mininf, model, w, batches[0]), approximation())
            fence()
            plan.record_sweep_events = False
            timed = [b.elapsed_time(e) for (b, e), kind in zip(plan.sweep_events, plan.sweep_event_kinds)
                     if kind == w.event_kind]
            return sum(timed) / max(len(timed), 1)
        plan.enqueue_sweeps()
        begin, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fence()
        begin.record()
        for i in range(repeats):
            plan.enqueue_sweeps()
        end.record()
        fence()
        return begin.elapsed_time(end) / repeats

    for _ in range(max(args.warmup, 3)):
        step()
    fence()
    kernel_ms_before = measure_sweep(args.steps)

    def timed_loop(n_steps, marks=None):
        begin, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fence()
        t0 = time.time()
        begin.record()
        loss = None
        for _ in range(n_steps):
            loss = step()
        end.record()
        fence()
        t1 = time.time()
        elapsed = begin.elapsed_time(end)
        by_rank = None
        if distributed:
            t = torch.tensor([elapsed], device=device, dtype=torch.float64)
            gathered = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(gathered, t)
            by_rank = [float(g) / n_steps for g in gathered]   # every rank's own clock: the step is the slowest one's
            elapsed = max(float(g) for g in gathered)
        return elapsed / n_steps, by_rank, loss, (t0, t1)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.nvtx.range_push("timed")       # lets `ncu --nvtx --nvtx-include timed/` see only these steps
    ms_per_step, by_rank, loss, (wall0, wall1) = timed_loop(args.steps)
    torch.cuda.nvtx.range_pop()
    clocks = sampler.stop(wall0, wall1) if rank == 0 else None
    for loss_module in loss_modules:
        loss_module.synchronize()
    final_loss = float(loss.detach())
    value = n_total * S / (ms_per_step * 1e-3)
    kernel_ms_after = measure_sweep(args.steps)
    kernel_ms = 0.5 * (kernel_ms_before + kernel_ms_after)

    # ---- sustained: at least `--sustain` seconds of back-to-back steps (the power cap engages) ---
    sustained = None
    if args.sustain > 0:
        n_sustain = max(args.steps, int(args.sustain * 1e3 / ms_per_step) + 1)
        s_ms, _, _, _ = timed_loop(n_sustain)
        sustained = {"steps": n_sustain, "ms_per_step": s_ms, "value": n_total * S / (s_ms * 1e-3), "unit": UNIT,
                     "seconds": s_ms * n_sustain * 1e-3}

    # ---- end to end: host buffers, H2D of the step's inputs and D2H of its result every step ----
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, w, step, batches, calls, n_rows, n_total, world, device, fence, distributed)

    # ---- secondary objects (rank 0 prints; every rank takes part in sharded evaluations) ----------
    secondary = {}
    if not args.no_secondary and fused and not plan.row_latents and (w.closed_form_kernel or w.black_box_kernel):
        other = not closed_form
        other_loss = make_loss(other)
        other_step = mininf.nn.FusedSVIStep(other_loss, condition_on(mininf, model, w, batches[0]), modules, lr=0.01,
                                            graph=graphed, share_with=steps_by_batch[0])
        saved = steps_by_batch
        steps_by_batch = [other_step]
        for _ in range(3):
            step()
        o_ms, _, _, _ = timed_loop(args.steps)
        steps_by_batch = saved
        other_loss.synchronize()
        secondary["closed_form" if other else "black_box"] = {
            "what": ("the same step with Normal / Poisson sites reduced to data-only sufficient statistics "
                     "(no per-particle work; value = rows * S / time is then free in S)" if other else
                     "the same step with one evaluation per (particle, element) for every site"),
            "kernel": w.closed_form_kernel if other else w.black_box_kernel,
            "ms_per_step": o_ms, "value": n_total * S / (o_ms * 1e-3), "unit": UNIT, "steps": args.steps}

    if rank == 0:
        peak, peak_source = measured_peak_gbs()
        algorithmic_bytes = n_rows * w.bytes_per_row
        achieved = algorithmic_bytes / (kernel_ms * 1e-3) / 1e9
        traffic = None
        traffic_file = ROOT / "profiles" / w.traffic_file
        if traffic_file.exists():
            try:
                per_row = json.loads(traffic_file.read_text())["dram_bytes_per_row"]
                traffic = per_row * n_rows
            except (KeyError, ValueError):
                traffic = None
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            rows_cpu, parts_cpu = w.cpu_rows, w.cpu_particles
            torch.set_num_threads(max(len(os.sched_getaffinity(0)), 1))
            cpu_value, _, kind = reference_sample(w, rows_cpu, parts_cpu, steps=4, warmup=1)
            cpu_baseline = {"value": cpu_value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
                            "sample": cpu_sample_text(w, rows_cpu, parts_cpu, torch.get_num_threads(), kind) +
                                      ", 4 timed steps"}
            if not args.no_secondary and load_reference() is not None:
                # SURVEY 8d's secondary comparator: the same unmodified reference with CUDA tensors
                # (stock torch eager kernels + cuBLAS), the only pre-existing Blackwell path
                try:
                    torch.cuda.empty_cache()
                    rows_gpu = min(n_rows, 10 * rows_cpu)
                    gpu_value, gpu_seconds, _ = reference_sample(w, rows_gpu, 2, steps=3, warmup=1, device=device)
                    secondary["reference_cuda"] = {
                        "what": "the unmodified reference (baseline/_ref) with CUDA tensors on this B200: stock torch "
                                "eager kernels + cuBLAS, one EvidenceLowerBoundLoss call per particle, validation on",
                        "value": gpu_value, "unit": UNIT, "ms_per_step": gpu_seconds * 1e3,
                        "sample": f"{rows_gpu} rows x 2 particles per step, 3 timed steps"}
                except Exception as error:  # noqa: BLE001  a secondary line must not fail the run
                    secondary["reference_cuda"] = {"unavailable": f"{type(error).__name__}: {error}"[:200]}
        step_kind = ("FusedSVIStep (mnf_svi_step: transforms + ELBO/grad kernels + chain rule + Adam in one native call)"
                     if fused else "GraphedStep (zero_grad + ELBO/grad kernels + backward + torch fused Adam)")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak" if args.weak else "strong", "vs_baseline": None,
            "dtype": dense_dtype(plan) if w.event_kind == "dense" else "f32 (f64 sums)",
            "data": "synthetic",
            "config": {"workload": w.name, "rows_total": n_total, "rows_per_gpu": n_rows,
                       "features": P, "particles": S,
                       "estimator": ("closed-form sufficient statistics where a site has them" if closed_form else
                                     "black-box: one log-density + score evaluation per (particle, observation)"),
                       "parallelism": (f"row shards x{world} (chunk ownership of one global data set), partial sums "
                                       f"combined by {'the engine over peer memory (NVLink)' if args.reduce == 'peer' else 'one NCCL all-reduce'}"
                                       if distributed else "single GPU"),
                       "l2": f"inputs ({n_rows * w.bytes_per_row / 1e9:.2f} GB per step and GPU) exceed the "
                             "126 MB L2; no flush needed",
                       "stream": (f"{w.n_batches} resident batches alternate; one recorded step per batch buffer, "
                                  "replayed in turn") if streaming else "one resident data set",
                       "step": step_kind + (", replayed from a CUDA graph" if graphed else ", eager launches"),
                       "final_loss": final_loss},
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": kernels_per_step * args.steps,
            "kernels_per_step": kernels_per_step,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_source,
                         "kernel": (w.closed_form_kernel if closed_form and w.closed_form_kernel else
                                    dense_kernel_name(w, plan)),
                         "kernel_ms": kernel_ms, "algorithmic_bytes": algorithmic_bytes,
                         "kernel_ms_before_after": [kernel_ms_before, kernel_ms_after],
                         "kernel_timing": f"one CUDA-event pair around {args.steps} back-to-back launches of the sweep "
                                          f"call (GPU-saturated), right before and right after the timed region "
                                          "on the same data, averaged",
                         **({"note": w.bound_note} if w.bound_note else {})},
            "sustained": sustained,
            "ms_per_step_by_rank": by_rank,
            "cpu_baseline": cpu_baseline,
            "steps_per_sec": 1e3 / ms_per_step,
            **secondary,
        }
        print(json.dumps(line))
    if distributed:
        dist.barrier()
        dist.destroy_process_group()


def run_e2e(args, w, step, batches, calls, n_rows, n_total, world, device, fence, distributed):
    """Same step, but the inputs live in pinned host memory and are copied to the device inside
    the timed region every step; the loss is read back to the host every step."""
    import psutil
    S = w.particles
    names = sorted(batches[0])
    row_bytes = sum(batches[0][k][0].numel() * batches[0][k].element_size() for k in names)
    need = n_rows * row_bytes
    available = psutil.virtual_memory().available
    rows = n_rows
    per_rank = world if distributed else 1
    if need * per_rank > 0.6 * available:
        rows = int(0.6 * available / per_rank / row_bytes)
    try:
        host = {k: torch.empty((rows,) + tuple(batches[0][k].shape[1:]), dtype=batches[0][k].dtype, pin_memory=True)
                for k in names}
    except RuntimeError:
        return {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "note": "could not pin host memory for the inputs"}
    for k in names:
        host[k].copy_(batches[0][k][:rows])
    steps = max(2, min(args.steps, 5))
    fence()
    begin, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    begin.record()
    for i in range(steps):
        # every input byte of the step crosses PCIe into the buffers the step reads; if the
        # pinned buffer is smaller than the data set it is sent repeatedly until all n_rows rows
        # have been overwritten
        target = batches[calls[0] % len(batches)]      # the batch buffer the next step reads
        for lo in range(0, n_rows, rows):
            m = min(rows, n_rows - lo)
            for k in names:
                target[k][lo:lo + m].copy_(host[k][:m], non_blocking=True)
        loss = step()
        loss_host = loss.item()        # device -> host read of the step's result
    end.record()
    fence()
    ms = begin.elapsed_time(end) / steps
    if distributed:
        import torch.distributed as dist
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    del loss_host
    return {"value": n_total * S / (ms * 1e-3), "unit": UNIT,
            "h2d_bytes_per_step": n_rows * row_bytes * (world if distributed else 1), "d2h_bytes_per_step": 4 * (world if distributed else 1),
            "ms_per_step": ms, "steps": steps,
            "note": ("inputs copied from pinned host memory every step (every rank its own shard); PCIe-bound"
                     + ("" if rows == n_rows else f"; pinned staging buffer of {rows} rows sent repeatedly"))}


def main():
    import warnings
    warnings.filterwarnings("ignore", message=".*MaskedTensors is in prototype.*")
    warnings.filterwarnings("ignore", message=".*Converting a tensor with requires_grad=True to a scalar.*")
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
