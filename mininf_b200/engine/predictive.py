"""Batched posterior predictive on the site table: ``broadcast_samples`` in one kernel launch.

The reference applies a model to a batch of posterior samples by re-running the Python model once
per sample under a ``SampleTracer`` (mininf/core.py:548-584, :192-204; example
examples/predictive.md:82-86). Here the model is traced ONCE with the samples wrapped as
:class:`~mininf_b200.engine.trace.LinkTensor` latents; every site the samples (or ``condition``)
do not provide is recorded with its distribution and the link of its parameters to the latents,
deterministic ``value`` sites with the link expression of their value. ``mnf_predictive``
(csrc/predict.cuh) then evaluates / draws all sites for all B samples in a single launch, block b
walking the sites in model order for sample b.

CUDA tensors only; the host loop of :func:`mininf_b200.core.broadcast_samples` serves CPU tensors.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from typing import Any, Callable, Dict, List, Optional, Tuple

import torch
from torch import distributions as td
from torch.distributions import Distribution

from . import abi
from .plan import LatentSpec, Plan
from .trace import Affine, Dense, LatentRef, Linear, LinkTensor
from ..core import State, TracerMixin, Value, _assert_same_batch_size
from ..util import OptionalSize, _normalize_shape, maybe_as_tensor

_FAMILIES = {td.Normal: abi.NORMAL, td.Gamma: abi.GAMMA, td.Beta: abi.BETA, td.Poisson: abi.POISSON}


def _as_latent(name: str, value: torch.Tensor) -> LinkTensor:
    ref = LatentRef(name, 0) if max(value.numel(), 1) == 1 else LatentRef(name)
    return LinkTensor.wrap(value, Affine(a_lat=ref))


class PredictiveTracer(TracerMixin):
    """One symbolic execution of the model: sites with a value pass through; everything else is
    recorded (in model order) and handed on as a new latent so later sites can depend on it."""

    def __init__(self, *args: Any, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        self.records: List[Tuple[str, str, Any, torch.Size]] = []     # (kind, name, payload, shape)

    def sample(self, state: State, name: str, distribution: Distribution,
               sample_shape: OptionalSize = None) -> torch.Tensor:
        shape = _normalize_shape(sample_shape)
        current = state.get(name)
        if current is not None:
            self._assert_valid_parameter(current, name, distribution, shape)
            return current
        if isinstance(distribution, Value):
            current = distribution.sample(shape)
            if isinstance(current, LinkTensor) and current.is_floating_point() and \
                    getattr(current, "_expr", None) is not None and current.numel() > 0:
                # a deterministic function of the samples: evaluated per sample by the kernel, and a
                # latent of its own for whatever depends on it
                self.records.append(("value", name, current._expr, current.shape))
                current = _as_latent(name, current.unwrap())
            elif isinstance(current, LinkTensor):
                raise NotImplementedError(f"value site '{name}' is not a supported link of the samples (supported: "
                                          "constants, data, `c + d*x`, `exp(a + b*x)`, `X @ theta`)")
            state[name] = current
            return current
        with torch.no_grad():
            placeholder = distribution.sample(shape)      # shapes, dtypes and validation as in the reference
        if isinstance(placeholder, LinkTensor):
            placeholder = placeholder.unwrap()
        if not placeholder.is_floating_point():
            raise NotImplementedError(f"site '{name}': only floating-point sites can be drawn on the device")
        self.records.append(("draw", name, distribution, placeholder.shape))
        current = _as_latent(name, placeholder)
        state[name] = current
        return current


class _Lowering(Plan):
    """Only the link lowering of :class:`Plan` (`_link`, `_latent_column`) over a column layout."""

    def __init__(self, specs: Dict[str, LatentSpec], device: torch.device) -> None:  # noqa: super-init-not-called
        self.by_name = specs
        self.device = device
        self.keepalive: List[torch.Tensor] = []
        self._folded_constant = False
        self.handle = None
        self.xrank = None


def _distribution_parameters(dist: Distribution, what: str) -> Tuple[int, Tuple[Any, ...]]:
    if isinstance(dist, td.Normal):
        return abi.NORMAL, (dist.loc, dist.scale)
    if isinstance(dist, td.Gamma):
        return abi.GAMMA, (dist.concentration, dist.rate)
    if isinstance(dist, td.Beta):
        return abi.BETA, (dist.concentration1, dist.concentration0)
    if isinstance(dist, td.Bernoulli):
        if "logits" in dist.__dict__:
            return abi.BERNOULLI_LOGITS, (dist.logits,)
        return abi.BERNOULLI_PROBS, (dist.probs,)
    if isinstance(dist, td.Poisson):
        return abi.POISSON, (dist.rate,)
    raise NotImplementedError(f"{what}: {type(dist).__name__} has no device sampler (supported: Normal, Gamma, "
                              "Beta, Bernoulli, Poisson)")


_calls = 0


def lower_predictive(model: Callable, given: Dict[str, torch.Tensor], B: int, device: torch.device
                     ) -> Tuple[List[abi.PredSite], Dict[str, LatentSpec], List[str], State, torch.Tensor, "_Lowering"]:
    """The host half of :func:`broadcast_samples`: trace ``model`` once and lower every site the
    samples do not provide to a ``mnf_pred_site_t``. Touches no CUDA API, so the lowering is
    testable without a GPU. Returns the site table, the column layout, the site order, the traced
    state, the sample matrix ``z`` [B, columns] with the given samples filled in, and the lowering
    object that keeps the tables' data tensors alive."""
    # ---- trace once, with sample 0 standing in for every sample --------------------------------------
    symbolic = {}
    for name, value in given.items():
        first = value[0]
        symbolic[name] = _as_latent(name, first.to(torch.float32)) if first.is_floating_point() and first.numel() > 0 \
            else first
    with State(symbolic) as traced, PredictiveTracer() as tracer:
        model()

    # ---- column layout: every latent-like name (given samples, value sites, drawn sites) ---------------
    specs: Dict[str, LatentSpec] = {}
    offset = 0
    order = list(traced.keys())
    for name in order:
        value = traced[name]
        if isinstance(value, LinkTensor):
            numel = max(value.numel(), 1)
            specs[name] = LatentSpec(name, abi.NORMAL, value.shape, numel, offset)
            offset += numel
    n_columns = max(offset, 1)
    z = torch.zeros(B, n_columns, device=device, dtype=torch.float32)
    for name, value in given.items():
        if name in specs:
            spec = specs[name]
            z[:, spec.offset:spec.offset + spec.numel] = value.to(torch.float32).reshape(B, spec.numel)

    # ---- lower the recorded sites ----------------------------------------------------------------------
    lower = _Lowering(specs, device)
    sites: List[abi.PredSite] = []
    for kind, name, payload, shape in tracer.records:
        spec = specs[name]
        what = f"site '{name}'"
        site = abi.PredSite(kind=abi.PRED_VALUE if kind == "value" else abi.PRED_DRAW, family=abi.NORMAL,
                            numel=spec.numel, out_col=spec.offset, transform=abi.T_ID, X=None, ldx=0, p=0,
                            theta_lat=-1, icpt_lat=-1, icpt_const=0.0)
        site.param[0] = abi.const_link(0.0)
        site.param[1] = abi.const_link(1.0)
        if kind == "value":
            first: Any = LinkTensor.wrap(torch.empty(shape, device=device), payload)
            params: Tuple[Any, ...] = (first,)
        else:
            site.family, params = _distribution_parameters(payload, what)
            if payload.event_shape:
                raise NotImplementedError(f"{what}: event-shaped distributions are not supported")
            params = tuple(p if isinstance(p, LinkTensor) else torch.as_tensor(p, device=device).expand(shape)
                           for p in params)
            first = params[0]
        expr = getattr(first, "_expr", None) if isinstance(first, LinkTensor) else None
        if kind == "draw" and site.family == abi.BERNOULLI_PROBS and getattr(expr, "transform", None) == "sigmoid":
            # Bernoulli(probs=sigmoid(eta)) draws as Bernoulli(logits=eta), like Plan._lower scores it
            site.family = abi.BERNOULLI_LOGITS
            expr = dataclasses.replace(expr, transform="id")
            first = LinkTensor.wrap(first.unwrap(), expr)
        if getattr(expr, "transform", "id") not in ("id", "exp"):
            raise NotImplementedError(f"{what}: `{expr.transform}` of a sample-dependent tensor is not a supported "
                                      "link here (sigmoid only as Bernoulli(probs=sigmoid(...)))")
        if isinstance(expr, (Dense, Linear)):
            if isinstance(expr, Linear):               # b1*x1 + b2*x2 + ...: a design matrix built here, as in Plan
                X, theta_offset = lower._linear_design(expr, spec.numel, what)
            else:
                X = expr.X.detach().to(device=device, dtype=torch.float32).contiguous()
                theta = specs.get(expr.theta)
                if theta is None or theta.numel != X.shape[1] or X.shape[0] != spec.numel:
                    raise NotImplementedError(f"{what}: `X @ theta` does not match the site")
                theta_offset = theta.offset
            lower.keepalive.append(X)
            site.X, site.ldx, site.p, site.theta_lat = X.data_ptr(), X.stride(0), X.shape[1], theta_offset
            site.icpt_const = expr.icpt_const
            if expr.icpt_lat is not None:
                site.icpt_lat, _ = lower._latent_column(expr.icpt_lat, 1, what)
            site.transform = abi.T_EXP if expr.transform == "exp" else abi.T_ID
        else:
            site.param[0] = lower._link(first, torch.Size([spec.numel]) if len(shape) == 0 else shape, what)
        if len(params) > 1:
            site.param[1] = lower._link(params[1], torch.Size([spec.numel]) if len(shape) == 0 else shape, what)
        sites.append(site)

    return sites, specs, order, traced, z, lower


def broadcast_samples(model: Callable, states: Dict[str, torch.Tensor]) -> State:
    """``broadcast_samples`` for CUDA samples: one trace, one launch, B samples."""
    global _calls
    lib = abi.load()
    B = _assert_same_batch_size(states)
    device = next(iter(states.values())).device
    given = {name: maybe_as_tensor(value) for name, value in states.items()}
    for name, value in given.items():
        if value.device != device:
            raise ValueError(f"all samples must live on {device}; '{name}' is on {value.device}")
    sites, specs, order, traced, z, lower = lower_predictive(model, given, B, device)
    n_columns = z.shape[1]

    # ---- one launch ------------------------------------------------------------------------------------
    if sites:
        table = (abi.PredSite * len(sites))(*sites)
        table_dev = torch.frombuffer(bytearray(bytes(table)), dtype=torch.uint8).to(device)
        status = torch.zeros(1, device=device, dtype=torch.int32)
        generator = torch.cuda.default_generators[device.index or 0]
        _calls += 1
        with torch.cuda.device(device):
            lib.call("mnf_predictive", table_dev.data_ptr(), len(sites), B, n_columns, z.data_ptr(),
                     int(generator.initial_seed()) & (2 ** 63 - 1), (1 << 45) + _calls, status.data_ptr(),
                     torch.cuda.current_stream(device).cuda_stream)
        bits = int(status.item())
        if bits:
            from .plan import status_message
            raise ValueError("the predictive sampler flagged invalid values: " + status_message(bits))

    # ---- the broadcast state: a leading batch dimension on every site ----------------------------------
    result = State()
    for name in order:
        if name in specs:
            spec = specs[name]
            block = z[:, spec.offset:spec.offset + spec.numel].reshape((B,) + tuple(spec.shape))
            result[name] = given[name] if name in given else block
        elif name in given:
            result[name] = given[name]
        else:
            constant = maybe_as_tensor(traced[name])
            result[name] = constant[None].expand((B,) + tuple(constant.shape)).clone()
    return result
