"""Lower a traced site table to the C-ABI tables and run ELBO steps on the CUDA engine.

A :class:`Plan` is built once per (model, conditioned tensors, approximation structure, particle
count) and then only enqueues kernels: ``mnf_rsample`` -> sweeps over the observed data
(``mnf_dense_sweep`` / ``mnf_site_sweep`` / ``mnf_small_sites``) -> optional all-reduce of the
[S][1+D] accumulator across ranks -> global (latent-valued) sites -> ``mnf_finalize``.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import torch
from torch import distributions as td

from . import abi
from .trace import Affine, Dense, LatentRef, Linear, LinkTensor, RowDot, SiteRecord
from ..util import is_masked

BIG_SITE = 2048          # observed sites at least this long go through the fused site sweep
ROW_LATENT = 8192        # Normal latent sites with more elements are per-observation ("row") latents

_LATENT_FAMILIES = {td.Normal: abi.NORMAL, td.Gamma: abi.GAMMA, td.Beta: abi.BETA}


def latent_parameters(dist: td.Distribution) -> Tuple[int, torch.Tensor, torch.Tensor]:
    """Family id and the two constrained parameter tensors of an approximation factor."""
    family = _LATENT_FAMILIES.get(type(dist))
    if family is None:
        raise NotImplementedError(
            f"the CUDA engine supports Normal, Gamma and Beta approximations, not {type(dist)}")
    if family == abi.NORMAL:
        return family, dist.loc, dist.scale
    if family == abi.GAMMA:
        return family, dist.concentration, dist.rate
    return family, dist.concentration1, dist.concentration0


@dataclasses.dataclass
class LatentSpec:
    name: str
    family: int
    shape: torch.Size
    numel: int
    offset: int          # first column in the packed z / noise; -1 for row latents
    row_latent: bool = False


def row_latent_names(sites: Sequence[SiteRecord]) -> set:
    """Latents a model uses as the per-observation matrix of a ``Z @ beta`` / ``z * slope`` link:
    whatever their size, they are swept by the row-latent kernel (the packed-latent sweeps have no
    link that is bilinear in two latents)."""
    names = set()
    for record in sites:
        for value in vars(record.distribution).values():
            if isinstance(value, LinkTensor) and isinstance(value._expr, RowDot):
                names.add(value._expr.Z)
    return names


def slope_groups(sites: Sequence[SiteRecord]) -> List[List[str]]:
    """Names of the latents whose columns a several-covariate link (``trace.Linear``) reads as ONE
    run of the packed latents, per site, in term order."""
    groups = []
    for record in sites:
        for value in vars(record.distribution).values():
            if isinstance(value, LinkTensor) and isinstance(value._expr, Linear):
                names: List[str] = []
                for ref, _ in value._expr.terms:
                    if ref.name not in names:
                        names.append(ref.name)
                groups.append(names)
    return groups


def _packing_order(names: List[str], together: Sequence[Sequence[str]]) -> List[str]:
    """``names`` reordered so that the members of every group in ``together`` are neighbours (groups
    that share a member are chained in order of appearance). The caller's order is kept otherwise;
    a layout no order satisfies is caught when the site is lowered (`Plan._linear_design`)."""
    clusters: List[List[str]] = []
    for group in together:
        touching = [c for c in clusters if any(name in c for name in group)]
        merged = [name for c in touching for name in c]
        merged += [name for name in group if name not in merged]
        clusters = [c for c in clusters if c not in touching] + [merged]
    placed: List[str] = []
    for name in names:
        if name in placed:
            continue
        cluster = next((c for c in clusters if name in c), None)
        placed.extend([n for n in cluster if n in names] if cluster else [name])
    return placed


def assign_offsets(entries: Sequence[Tuple[str, int, torch.Size]], rows: Sequence[str] = (),
                   together: Sequence[Sequence[str]] = ()) -> List[LatentSpec]:
    """Latent specs for (name, family, shape) triples: small sites are packed into z[S][D] in
    order; large Normal sites - and those named in ``rows`` (:func:`row_latent_names`) - become row
    latents that are never materialised. ``together`` (:func:`slope_groups`) lists latents that must
    sit in neighbouring columns; the specs come back in packing order (the caller's, unless a group moved)."""
    by_name = {name: (name, family, shape) for name, family, shape in entries}
    order = _packing_order([name for name, _, _ in entries], together) if together else list(by_name)
    packed: Dict[str, LatentSpec] = {}
    offset = 0
    for name in order:
        _, family, shape = by_name[name]
        numel = max(shape.numel(), 1)
        if (numel > ROW_LATENT or name in rows) and family == abi.NORMAL and len(shape) in (1, 2):
            packed[name] = LatentSpec(name, family, shape, numel, -1, True)
        else:
            packed[name] = LatentSpec(name, family, shape, numel, offset)
            offset += numel
    return [packed[name] for name in order]          # packing order: the latent table stays sorted by column


def _bytes_to_device(struct_array: Any, device: torch.device) -> torch.Tensor:
    return torch.frombuffer(bytearray(bytes(struct_array)), dtype=torch.uint8).to(device)


def _f32(tensor: torch.Tensor, device: torch.device) -> torch.Tensor:
    if isinstance(tensor, LinkTensor):
        tensor = tensor.unwrap()
    if tensor.device != device:
        raise ValueError(f"all tensors of an ELBO evaluation must live on {device}; found one on "
                         f"{tensor.device}")
    return tensor.detach().to(torch.float32).contiguous()


def _fits_f16_operands(X: torch.Tensor) -> bool:
    """Whether a design matrix may be fed to the tensor cores as IEEE half precision: no entry at
    or above 2^15 (the kernel re-checks that on every step, MNF_ST_RANGE), and no non-zero column
    whose root-mean-square is below 2^-8 - below 2^-14 fp16 spacing is an absolute 2^-24, which
    would no longer be small against such a column's entries (TF32 has the fp32 exponent range)."""
    if X.numel() == 0:
        return True
    lo, hi = torch.aminmax(X)
    if not (max(abs(float(lo)), abs(float(hi))) < 2.0 ** 15):      # also False for NaN
        return False
    rms = torch.linalg.vector_norm(X, dim=0) / X.shape[0] ** 0.5
    return bool(((rms == 0) | (rms >= 2.0 ** -8)).all())


def _has_fast_sweep(site: abi.Site) -> bool:
    first, second = site.param[0], site.param[1]
    return (site.family == abi.POISSON and first.transform == abi.T_EXP) or \
        (site.family == abi.NORMAL and first.transform == abi.T_ID and not second.x)


def _has_moment_path(site: abi.Site) -> bool:
    """Poisson(exp) site whose layout qualifies for the Chebyshev-moment kernels
    (csrc/site.cu::launch_poisson_moments): unit-stride covariate, 16-byte aligned data."""
    first = site.param[0]
    if not (site.family == abi.POISSON and first.transform == abi.T_EXP and first.x and first.x_stride == 1):
        return False
    return first.x % 16 == 0 and (site.value or 0) % 16 == 0 and (site.mask or 0) % 4 == 0


class Plan:
    """Everything one ELBO step needs, resolved to device pointers."""

    def __init__(self, sites: Sequence[SiteRecord], latents: Sequence[LatentSpec],
                 n_particles: int, device: torch.device, dense_mode: str = "auto",
                 dry_run: bool = False, closed_form: bool = False) -> None:
        """``dry_run`` lowers the site table without touching the GPU (host-logic tests).
        ``closed_form`` lets sites whose log-density is a closed form of data-only sufficient
        statistics skip the per-(particle, observation) evaluation (Gram statistics of a Normal
        dense site with p <= 64, six sums / Chebyshev moments of scalar-link Normal / Poisson
        sites); off, every site is swept once per particle and observation."""
        if device.type != "cuda" and not dry_run:
            raise RuntimeError("the mininf_b200 ELBO engine runs on CUDA tensors only; move the "
                               "approximation parameters and the conditioned data to the GPU")
        self.lib = abi.load()
        self.device = device
        self.S = int(n_particles)
        self.all_latents = list(latents)
        self.latents = [spec for spec in latents if not spec.row_latent]     # packed into z[S][D]
        self.row_latents = [spec for spec in latents if spec.row_latent]
        self.by_name = {spec.name: spec for spec in self.all_latents}
        self.D = sum(spec.numel for spec in self.latents)
        if self.S < 1 or self.D < 1:
            raise ValueError("need at least one particle and one latent element")
        self.dense_mode = dense_mode
        self.closed_form = bool(closed_form)
        self.keepalive: List[torch.Tensor] = []
        self.all_normal = all(spec.family == abi.NORMAL for spec in self.latents)

        S, D = self.S, self.D
        f32 = dict(device=device, dtype=torch.float32)
        # constrained parameters of the packed latents, [p0 columns | p1 columns] in one buffer
        # (the fused SVI step writes transform(raw) here, include/mininf_b200.h::mnf_adam_t)
        self.P = torch.cat([torch.zeros(D, **f32), torch.ones(D, **f32)])
        self.P0, self.P1 = self.P[:D], self.P[D:]
        self.z = torch.empty(S, D, **f32)
        self.noise = torch.empty(S, D, **f32)
        self.noise_in = torch.empty(S, D, **f32)
        self.acc = torch.empty(S, D + 1, device=device, dtype=torch.float64)
        self.status = torch.zeros(1, device=device, dtype=torch.int32)
        self.step_counter = torch.zeros(1, device=device, dtype=torch.int64)   # graph replays (see step)
        self.out = torch.empty(1 + 2 * D, **f32)
        self.dry_run = dry_run
        self.workspace_bytes = 0
        self.workspace = torch.empty(0, device=device, dtype=torch.uint8)

        table = (abi.Latent * len(self.latents))()
        for i, spec in enumerate(self.latents):
            table[i] = abi.Latent(family=spec.family, numel=spec.numel, offset=spec.offset, reserved=0,
                                  p0=self.P0.data_ptr() + 4 * spec.offset,
                                  p1=self.P1.data_ptr() + 4 * spec.offset)
        self._latent_host = table
        self.latent_table = _bytes_to_device(table, device)
        self.handle: Optional[C.c_void_p] = None      # mnf_plan_t*, created after lowering
        self.xrank: Optional[Any] = None              # peer-memory exchange (sharded evaluations)
        self.launches_last_step = 0

        # optional instrumentation: CUDA event pairs around every dense sweep (bench.py roofline)
        self.record_sweep_events = False
        self.sweep_events: List[Tuple[torch.cuda.Event, torch.cuda.Event]] = []
        self.sweep_event_kinds: List[str] = []

        # one descriptor per row latent, filled in by the sites that touch it
        self.row_groups: Dict[str, abi.RowLatent] = {}
        for spec in self.row_latents:
            n_rows = spec.shape[0]
            desc = abi.RowLatent(n_rows=n_rows, p=spec.numel // n_rows, resp_family=-1, beta_lat=-1,
                                 icpt_lat=-1, icpt_const=0.0, resp_transform=abi.T_ID)
            desc.prior_loc = abi.const_link(0.0)
            desc.prior_scale = abi.const_link(1.0)
            desc.feat_scale = abi.const_link(1.0)
            desc.resp_scale = abi.const_link(1.0)
            desc._has_prior = False
            self.row_groups[spec.name] = desc

        self.dense_sites: List[Tuple[abi.DenseSite, int]] = []
        self.sweep_groups: List[Any] = []
        self._host_tables: List[Tuple[Any, torch.Tensor]] = []   # (host struct array, its device copy)
        self._folded_constant = False     # a tensor parameter was folded into a constant link
        self._sources: List[Tuple[int, int]] = []
        self.rebindable = False
        small_observed: List[abi.Site] = []
        small_global: List[abi.Site] = []
        big: Dict[int, List[abi.Site]] = {}
        for record in sites:
            self._lower(record, small_observed, small_global, big)
        for numel, group in big.items():
            for start in range(0, len(group), abi.MAX_FUSED_SITES):
                chunk = group[start:start + abi.MAX_FUSED_SITES]
                array = (abi.Site * len(chunk))(*chunk)
                self.sweep_groups.append(array)
        self.small_observed = self._site_table(small_observed)
        self.small_global = self._site_table(small_global)
        for name, desc in self.row_groups.items():
            if not desc._has_prior:
                raise NotImplementedError(f"row latent '{name}' has no prior site in the model")
        self._small_observed_host = small_observed
        self._small_global_host = small_global
        self._create_native()

    # ------------------------------------------------------------------------------------------
    # the native plan (include/mininf_b200.h::mnf_plan_create): one C call per step
    # ------------------------------------------------------------------------------------------
    def _describe(self) -> abi.PlanDesc:
        """The flat tables as one ``mnf_plan_desc_t`` (host arrays are kept alive on ``self``)."""
        dense = (abi.DenseSite * max(len(self.dense_sites), 1))(*[site for site, _ in self.dense_sites])
        modes = (C.c_int32 * max(len(self.dense_sites), 1))(*[mode for _, mode in self.dense_sites])
        grouped = [group[i] for group in self.sweep_groups for i in range(len(group))]
        group_sites = (abi.Site * max(len(grouped), 1))(*grouped)
        group_sizes = (C.c_int32 * max(len(self.sweep_groups), 1))(*[len(group) for group in self.sweep_groups])
        observed = self.small_observed[3] if self.small_observed is not None else (abi.Site * 1)()
        priors = self.small_global[3] if self.small_global is not None else (abi.Site * 1)()
        rows = (abi.RowLatent * max(len(self.row_groups), 1))(*self.row_groups.values())
        self._desc_arrays = (dense, modes, group_sites, group_sizes, observed, priors, rows)
        return abi.PlanDesc(
            n_particles=self.S, n_latent_total=self.D, n_latents=len(self.latents),
            n_dense=len(self.dense_sites), n_groups=len(self.sweep_groups),
            n_small_observed=self.small_observed[1] if self.small_observed is not None else 0,
            n_small_global=self.small_global[1] if self.small_global is not None else 0,
            n_rowlatent=len(self.row_groups), latents=self._latent_host, dense=dense, dense_mode=modes,
            group_sites=group_sites, group_sizes=group_sizes, small_observed=observed, small_global=priors,
            rowlatent=rows, flags=abi.SWEEP_CLOSED_FORM if self.closed_form else 0, device=-1)

    def _create_native(self) -> None:
        if not self.dry_run:
            with torch.cuda.device(self.device):
                handle = C.c_void_p()
                desc = self._describe()
                self.lib.call("mnf_plan_create", C.byref(desc), C.byref(handle))
                self.handle = handle
                size = C.c_size_t()
                self.lib.call("mnf_plan_workspace_bytes", handle, C.byref(size))
            self.workspace_bytes = int(size.value)
            self.workspace = torch.empty(self.workspace_bytes, device=self.device, dtype=torch.uint8)
        self._row_buffers = (abi.RowBuffers * max(len(self.row_groups), 1))()
        self.buffers = abi.Buffers(
            z=self.z.data_ptr(), noise=self.noise.data_ptr(), acc=self.acc.data_ptr(), out=self.out.data_ptr(),
            workspace=self.workspace.data_ptr(), workspace_bytes=self.workspace_bytes,
            status=self.status.data_ptr(), noise_in=None, step_counter=None, rows=self._row_buffers, xrank=None)

    def __del__(self) -> None:
        try:
            if getattr(self, "handle", None):
                self.lib.raw("mnf_plan_destroy")(self.handle)
                self.handle = None
            if getattr(self, "xrank", None):
                self.lib.raw("mnf_xrank_destroy")(self.xrank)
                self.xrank = None
        except Exception:  # noqa: BLE001  interpreter shutdown
            pass

    # ------------------------------------------------------------------------------------------
    # lowering
    # ------------------------------------------------------------------------------------------
    def _site_table(self, sites: List[abi.Site]) -> Optional[Tuple[torch.Tensor, int, int, Any]]:
        if not sites:
            return None
        array = (abi.Site * len(sites))(*sites)
        table = _bytes_to_device(array, self.device)
        self._host_tables.append((array, table))
        return table, len(sites), max(s.numel for s in sites), array

    def _latent_column(self, ref: LatentRef, numel: int, what: str) -> Tuple[int, int]:
        """(column, stride) of a latent reference inside a site of ``numel`` elements."""
        spec = self.by_name.get(ref.name)
        if spec is None:
            raise NotImplementedError(f"{what} depends on '{ref.name}', which the approximation "
                                      "does not provide")
        if spec.row_latent:
            raise NotImplementedError(f"{what}: per-observation latent '{ref.name}' is only supported "
                                      "as a site value, as the location of an observed Normal site or "
                                      "in `Z @ beta`")
        if ref.is_scalar:
            return spec.offset + ref.index, 0
        if spec.numel == 1:
            return spec.offset, 0
        if spec.numel != numel:
            raise NotImplementedError(f"{what}: latent '{ref.name}' with {spec.numel} elements is "
                                      f"not aligned with a site of {numel} elements")
        return spec.offset, 1

    def _link(self, param: Any, shape: torch.Size, what: str) -> abi.Link:
        """A distribution parameter (already broadcast by the distribution) as a scalar link."""
        numel = shape.numel()
        if isinstance(param, LinkTensor):
            expr = param._expr
            if not isinstance(expr, Affine):
                raise NotImplementedError(
                    f"{what} is not a supported link of the latent variables (supported: a latent "
                    "itself, constants, data, `c + d*x` and what reduces to it, `exp(a + b*x)`, `X @ theta`, "
                    "`a + b1*x1 + b2*x2` with scalar latents; see README.md)")
            if expr.transform not in ("id", "exp"):
                raise NotImplementedError(
                    f"{what}: `{expr.transform}` of a latent-dependent tensor is not a supported link "
                    "here (sigmoid only as Bernoulli(probs=sigmoid(...)))")
            link = abi.Link(a_const=expr.a_const, b_const=expr.b_const, a_lat=-1, b_lat=-1,
                            a_stride=0, b_stride=0, x=None, x_stride=0,
                            transform=abi.T_EXP if expr.transform == "exp" else abi.T_ID)
            if expr.a_lat is not None:
                link.a_lat, link.a_stride = self._latent_column(expr.a_lat, numel, what)
            if expr.b_lat is not None:
                link.b_lat, link.b_stride = self._latent_column(expr.b_lat, numel, what)
            if expr.x is not None:
                x = _f32(expr.x.expand(shape), self.device).reshape(-1)
                self.keepalive.append(x)
                link.x, link.x_stride = x.data_ptr(), 1
            elif expr.has_x_term:
                raise NotImplementedError(f"{what}: slope term without a covariate")
            return link
        tensor = torch.as_tensor(param)
        if tensor.numel() == 1 or bool((tensor == tensor.reshape(-1)[0]).all()):
            # an expanded scalar (all strides 0, e.g. `Normal(eta, 1.0)` after broadcast_all) IS a
            # constant; a materialised tensor that merely holds equal values could differ in the next batch
            expanded = all(stride == 0 or size == 1 for stride, size in zip(tensor.stride(), tensor.shape))
            self._folded_constant = self._folded_constant or (tensor.numel() > 1 and not expanded)
            return abi.const_link(float(tensor.reshape(-1)[0]))
        data = _f32(tensor.expand(shape), self.device).reshape(-1)
        self.keepalive.append(data)
        link = abi.const_link(0.0)
        link.b_const, link.x, link.x_stride = 1.0, data.data_ptr(), 1
        return link

    def _lower(self, record: SiteRecord, small_observed: List[abi.Site],
               small_global: List[abi.Site], big: Dict[int, List[abi.Site]]) -> None:
        dist, value, name = record.distribution, record.value, record.name
        what = f"site '{name}'"
        # family and raw parameters
        if isinstance(dist, td.Normal):
            family, params = abi.NORMAL, (dist.loc, dist.scale)
        elif isinstance(dist, td.Gamma):
            family, params = abi.GAMMA, (dist.concentration, dist.rate)
        elif isinstance(dist, td.Beta):
            family, params = abi.BETA, (dist.concentration1, dist.concentration0)
        elif isinstance(dist, td.Bernoulli):
            if "logits" in dist.__dict__:
                family, params = abi.BERNOULLI_LOGITS, (dist.logits,)
            else:
                family, params = abi.BERNOULLI_PROBS, (dist.probs,)
        elif isinstance(dist, td.Poisson):
            family, params = abi.POISSON, (dist.rate,)
        else:
            raise NotImplementedError(f"{what}: {type(dist).__name__} has no CUDA log-density "
                                      "(supported: Normal, Gamma, Beta, Bernoulli, Poisson)")
        if dist.event_shape:
            raise NotImplementedError(f"{what}: event-shaped distributions are not supported")
        if family == abi.BERNOULLI_PROBS and isinstance(params[0], LinkTensor) and \
                getattr(params[0]._expr, "transform", None) == "sigmoid":
            # Bernoulli(probs=sigmoid(eta)) is Bernoulli(logits=eta). The reference goes through
            # probs_to_logits (TORCH distributions/utils.py: clamp to [eps, 1 - eps], log p - log1p(-p)),
            # which agrees with the logits form to fp32 rounding for |eta| < 15 and saturates beyond.
            family = abi.BERNOULLI_LOGITS
            params = (LinkTensor.wrap(params[0].unwrap(),
                                      dataclasses.replace(params[0]._expr, transform="id")),)

        if self._lower_row_latent(record, family, params, what):
            return

        # value: a latent itself (prior site) or observed data (possibly masked)
        mask = None
        value_lat = -1
        if isinstance(value, LinkTensor):
            expr = value._expr
            if not (isinstance(expr, Affine) and expr.is_pure_latent):
                raise NotImplementedError(f"{what}: its value is a function of latent variables")
            shape, numel, data_ptr = value.shape, value.numel(), None
            value_lat, stride = self._latent_column(expr.a_lat, numel, what)
            if stride == 0 and numel != 1:
                raise NotImplementedError(f"{what}: a scalar latent broadcast over the site")
        else:
            if is_masked(value):
                mask = value.get_mask().contiguous()
                if mask.device != self.device:
                    raise ValueError(f"{what}: mask lives on {mask.device}, expected {self.device}")
                self.keepalive.append(mask)
                value = value.get_data()
            data = _f32(value, self.device)
            self.keepalive.append(data)
            shape, numel, data_ptr = data.shape, data.numel(), data.data_ptr()
        if numel == 0:
            return

        # dense linear predictor in the first parameter?
        first = params[0]
        if isinstance(first, LinkTensor) and isinstance(first._expr, (Dense, Linear)):
            self._lower_dense(record, family, params, data_ptr, mask, numel, what)
            return

        site = abi.Site(family=family, value_lat=value_lat, value=data_ptr,
                        mask=mask.data_ptr() if mask is not None else None, numel=numel,
                        scale=float(record.scale))
        site.param[0] = self._link(params[0], shape, what)
        site.param[1] = self._link(params[1], shape, what) if len(params) > 1 else abi.const_link(1.0)
        scalar_links = all(not ((link.a_lat >= 0 and link.a_stride) or (link.b_lat >= 0 and link.b_stride))
                           for link in (site.param[0], site.param[1]))
        if value_lat >= 0:
            small_global.append(site)
        elif numel >= BIG_SITE and scalar_links:
            big.setdefault(numel, []).append(site)
        else:
            small_observed.append(site)

    def _scalar_link(self, param: Any, what: str) -> abi.Link:
        link = self._link(param, torch.Size([1]), what)
        if link.x or link.b_lat >= 0 or link.a_stride:
            raise NotImplementedError(f"{what}: expected a constant or a scalar latent")
        return link

    def _lower_row_latent(self, record: SiteRecord, family: int, params: Tuple[Any, ...], what: str) -> bool:
        """Sites touching a per-observation latent go into that latent's row-latent descriptor."""
        value = record.value

        def row_spec(tensor: Any) -> Optional[LatentSpec]:
            expr = getattr(tensor, "_expr", None) if isinstance(tensor, LinkTensor) else None
            if isinstance(expr, Affine) and expr.is_pure_latent and not expr.a_lat.is_scalar:
                spec = self.by_name.get(expr.a_lat.name)
                if spec is not None and spec.row_latent:
                    return spec
            return None

        spec = row_spec(value)
        if spec is not None:                                   # prior: Z ~ Normal(const, const | latent)
            if family != abi.NORMAL or record.scale != 1.0:
                raise NotImplementedError(f"{what}: row latents need an unbatched Normal prior")
            desc = self.row_groups[spec.name]
            desc.prior_loc = self._scalar_link(params[0], what)
            desc.prior_scale = self._scalar_link(params[1], what)
            if desc.prior_loc.a_lat >= 0:
                raise NotImplementedError(f"{what}: the prior location of a row latent must be constant")
            desc._has_prior = True
            return True
        first = params[0]
        spec = row_spec(first)
        if spec is not None:                                   # features: X ~ Normal(Z, const)
            if family != abi.NORMAL or isinstance(value, LinkTensor) or is_masked(value) or \
                    record.scale != 1.0 or value.numel() != spec.numel:
                raise NotImplementedError(f"{what}: only an unmasked Normal(loc=Z, scale=const) site can "
                                          "observe a row latent element-wise")
            desc = self.row_groups[spec.name]
            data = _f32(value, self.device)
            self.keepalive.append(data)
            desc.feat = data.data_ptr()
            desc.feat_scale = self._scalar_link(params[1], what)
            if desc.feat_scale.a_lat >= 0:
                raise NotImplementedError(f"{what}: the feature scale must be a constant")
            return True
        expr = getattr(first, "_expr", None) if isinstance(first, LinkTensor) else None
        if isinstance(expr, RowDot):                           # response: y ~ F(T(icpt + Z @ beta))
            spec = self.by_name[expr.Z]
            beta = self.by_name.get(expr.beta)
            if not spec.row_latent or beta is None or beta.row_latent or isinstance(value, LinkTensor) or \
                    is_masked(value) or record.scale != 1.0:
                raise NotImplementedError(f"{what}: unsupported `Z @ beta` site")
            desc = self.row_groups[spec.name]
            if beta.numel != desc.p or value.numel() != desc.n_rows:
                raise NotImplementedError(f"{what}: shapes of Z, beta and the response do not match")
            ok = (family == abi.POISSON and expr.transform in ("id", "exp")) or \
                (family in (abi.NORMAL, abi.BERNOULLI_LOGITS) and expr.transform == "id")
            if not ok:
                raise NotImplementedError(f"{what}: `Z @ beta` supports Poisson, Normal(loc) and Bernoulli(logits)")
            data = _f32(value, self.device)
            self.keepalive.append(data)
            desc.resp, desc.resp_family = data.data_ptr(), family
            desc.resp_transform = abi.T_EXP if expr.transform == "exp" else abi.T_ID
            desc.beta_lat, desc.icpt_const = beta.offset, expr.icpt_const
            if expr.icpt_lat is not None:
                desc.icpt_lat, _ = self._latent_column(expr.icpt_lat, 1, what)
            if family == abi.NORMAL:
                desc.resp_scale = self._scalar_link(params[1], what)
            return True
        return False

    def _linear_design(self, expr: Linear, numel: int, what: str) -> Tuple[torch.Tensor, int]:
        """Design matrix and first latent column of a several-covariate link ``sum_k z[c_k] * x_k``:
        one column per distinct latent (covariates of a repeated latent add up), ordered like the packed
        latents, which must be adjacent there - the dense kernels read theta as one run of z."""
        columns: Dict[int, torch.Tensor] = {}
        names: Dict[int, str] = {}
        for ref, x in expr.terms:
            spec = self.by_name.get(ref.name)
            if spec is not None and not ref.is_scalar and spec.numel > 1:
                # a whole coefficient vector with its block X[n, G] (`X @ theta`, `alpha[group]`)
                first, _ = self._latent_column(LatentRef(ref.name, 0), 1, what)
                block = _f32(x, self.device)
                if block.ndim != 2 or tuple(block.shape) != (numel, spec.numel):
                    raise NotImplementedError(f"{what}: the block of '{ref.name}' does not match the site")
                for g in range(spec.numel):
                    columns[first + g] = columns[first + g] + block[:, g] if first + g in columns else block[:, g]
                    names[first + g] = f"{ref.name}[{g}]"
                continue
            column, _ = self._latent_column(ref, 1, what)
            values = _f32(x, self.device).expand(numel) if x.numel() == 1 else _f32(x, self.device).reshape(-1)
            if values.numel() != numel:
                raise NotImplementedError(f"{what}: a covariate with {values.numel()} elements in a site of {numel}")
            columns[column] = columns[column] + values if column in columns else values
            names[column] = ref.name if ref.index in (None, 0) and self.by_name[ref.name].numel == 1 \
                else f"{ref.name}[{ref.index}]"
        order = sorted(columns)
        if order != list(range(order[0], order[0] + len(order))):
            raise NotImplementedError(
                f"{what}: the slopes of a several-covariate link must be adjacent in the approximation "
                f"(packed columns {order} for {[names[c] for c in order]}); list these latents next to each other")
        return torch.stack([columns[c] for c in order], dim=1).contiguous(), order[0]

    def _lower_dense(self, record: SiteRecord, family: int, params: Tuple[Any, ...],
                     data_ptr: Optional[int], mask: Optional[torch.Tensor], numel: int,
                     what: str) -> None:
        expr: Dense = params[0]._expr
        if data_ptr is None:
            raise NotImplementedError(f"{what}: a dense-link site must be observed")
        if family == abi.NORMAL and expr.transform == "id":
            dense_family = abi.NORMAL
        elif family == abi.BERNOULLI_LOGITS and expr.transform == "id":
            dense_family = abi.BERNOULLI_LOGITS
        elif family == abi.POISSON and expr.transform == "exp":
            dense_family = abi.POISSON
        else:
            raise NotImplementedError(f"{what}: dense links are supported for Normal(loc=X@theta), "
                                      "Bernoulli(logits=X@theta) and Poisson(rate=exp(X@theta))")
        if isinstance(expr, Linear):
            X, theta_offset = self._linear_design(expr, numel, what)
        else:
            X = _f32(expr.X, self.device)
            theta = self.by_name.get(expr.theta)
            if theta is None or theta.numel != X.shape[1]:
                raise NotImplementedError(f"{what}: coefficient vector '{expr.theta}' does not match X")
            theta_offset = theta.offset
        self.keepalive.append(X)
        n, p = X.shape
        if n != numel:
            raise NotImplementedError(f"{what}: X has {n} rows but the site has {numel} elements")
        icpt_lat = -1
        if expr.icpt_lat is not None:
            icpt_lat, _ = self._latent_column(expr.icpt_lat, 1, what)
        scale_link = abi.const_link(1.0)
        if dense_family == abi.NORMAL:
            scale_link = self._link(params[1], torch.Size([numel]), what)
            if scale_link.x:
                raise NotImplementedError(f"{what}: per-observation scales are not supported for "
                                          "dense-link sites")
            if scale_link.b_lat >= 0 or scale_link.a_stride:
                raise NotImplementedError(f"{what}: the scale must be a constant or a scalar latent")
        site = abi.DenseSite(family=dense_family, p=p, n_rows=n, ldx=X.stride(0), X=X.data_ptr(),
                             y=data_ptr, mask=mask.data_ptr() if mask is not None else None,
                             theta_lat=theta_offset, icpt_lat=icpt_lat, icpt_const=expr.icpt_const,
                             reserved=0, scale=scale_link, weight=float(record.scale))
        has_icpt = icpt_lat >= 0 or expr.icpt_const != 0.0
        kernel = self.lib.raw("mnf_dense_tf32_kernel")(dense_family, p, self.S)
        if kernel == 3 and not self.closed_form:
            kernel = 0          # only the Gram closed form covers this shape
        tf32_ok = kernel != 0 and X.data_ptr() % 16 == 0 and X.stride(0) % 4 == 0 and n < 2 ** 31
        if self.dense_mode == "tf32" and not tf32_ok:
            raise NotImplementedError(f"{what}: the tcgen05 TF32 kernels need p == 64 with at most 64 "
                                      "particles, or p a multiple of 4 with at most 128 particles, "
                                      "and 16-byte aligned rows")
        tf32 = abi.DENSE_TF32_CLOSED_FORM if self.closed_form else abi.DENSE_TF32
        # fp16 operands (csrc/dense_th.cuh: the same 11-bit significand as TF32 at half the tensor-pipe
        # cost) for the p = 64 / S <= 64 black-box shape, when the design matrix fits fp16's range
        f16_shape = p == 64 and self.S <= 64 and not self.closed_form and X.data_ptr() % 16 == 0 and \
            X.stride(0) % 4 == 0 and n < 2 ** 31
        f16_ok = f16_shape and self.dense_mode in ("auto", "f16") and _fits_f16_operands(X)
        if self.dense_mode == "f16" and not f16_ok:
            raise NotImplementedError(f"{what}: the fp16-operand kernel needs p == 64, at most 64 particles, the "
                                      "black-box estimator, 16-byte aligned rows and a design matrix inside fp16's "
                                      "range (|x| < 2^15, no column with a root-mean-square below 2^-8)")
        if f16_ok:
            mode = abi.DENSE_F16
        else:
            mode = tf32 if (self.dense_mode in ("auto", "tf32") and tf32_ok) else abi.DENSE_FP32
        self.dense_sites.append((site, mode))

    # ------------------------------------------------------------------------------------------
    # rebinding to new conditioned tensors of the same layout (minibatch streams)
    # ------------------------------------------------------------------------------------------
    _PER_STEP_FIELDS = ("loc", "scale", "grad_loc", "grad_scale", "eps")   # RowLatent, set by every step

    def _host_structs(self) -> List[Any]:
        structs: List[Any] = [site for site, _ in self.dense_sites]
        for group in self.sweep_groups:
            structs.extend(group[i] for i in range(len(group)))
        for array, _ in self._host_tables:
            structs.extend(array[i] for i in range(len(array)))
        structs.extend(self.row_groups.values())
        return structs

    def _walk_pointers(self, struct: Any, visit: Callable[[Any, str, int], None]) -> None:
        per_step = self._PER_STEP_FIELDS if isinstance(struct, abi.RowLatent) else ()
        for name, ctype in struct._fields_:
            if ctype is C.c_void_p:
                value = getattr(struct, name)
                if value and name not in per_step:
                    visit(struct, name, value)
            elif isinstance(ctype, type) and issubclass(ctype, C.Structure):
                self._walk_pointers(getattr(struct, name), visit)
            elif isinstance(ctype, type) and issubclass(ctype, C.Array):
                array = getattr(struct, name)
                for i in range(len(array)):
                    if isinstance(array[i], C.Structure):
                        self._walk_pointers(array[i], visit)

    @staticmethod
    def _ranges(leaves: Sequence[torch.Tensor]) -> List[Tuple[int, int]]:
        return [(t.data_ptr(), t.data_ptr() + t.numel() * t.element_size()) for t in leaves]

    def bind_sources(self, leaves: Sequence[torch.Tensor]) -> None:
        """Declare the conditioned tensors this plan was lowered from. The plan becomes
        *rebindable* when every data pointer it stores lies inside one of them (no converted or
        re-laid-out copies, no tensor folded into a constant, no overlapping sources): then a new
        batch of the same shapes, dtypes and strides only needs its pointers patched."""
        ranges = self._ranges(leaves)
        self._sources = ranges
        contiguous = all(t.is_contiguous() for t in leaves)
        ordered = sorted(r for r in ranges if r[1] > r[0])
        disjoint = all(a[1] <= b[0] for a, b in zip(ordered, ordered[1:]))
        resolved = [True]

        def visit(struct: Any, name: str, value: int) -> None:
            if not any(lo <= value < hi for lo, hi in ranges):
                resolved[0] = False

        for struct in self._host_structs():
            self._walk_pointers(struct, visit)
        # a one-element conditioned tensor may have been folded into a link constant by the tracer
        # (`a + scale0`, `Normal(eta, scale0)`): its VALUE is in the tables, not its address, so new
        # contents - in place or behind a new pointer - need a fresh trace
        scalar_leaf = any(t.numel() == 1 and t.is_floating_point() for t in leaves)
        self.rebindable = contiguous and disjoint and resolved[0] and not self._folded_constant and \
            not scalar_leaf

    def rebind(self, leaves: Sequence[torch.Tensor]) -> bool:
        """Point the plan at new conditioned tensors (same order and layout as the ones given to
        :meth:`bind_sources`). Returns ``False`` if the new tensors cannot be used as they are
        (the caller then lowers a fresh plan). The support checks the reference repeats on every
        call (mininf/core.py:142-189) are left to the kernels' status word for the new data."""
        new = self._ranges(leaves)
        old = self._sources
        if not self.rebindable or len(new) != len(old) or \
                any(a[1] - a[0] != b[1] - b[0] for a, b in zip(old, new)) or \
                not all(t.is_contiguous() and t.device == self.device for t in leaves):
            return False
        if new == old:
            return True
        ordered = sorted(r for r in new if r[1] > r[0])
        if not all(a[1] <= b[0] for a, b in zip(ordered, ordered[1:])):
            return False

        def translate(value: int) -> int:
            for (lo, hi), (new_lo, _) in zip(old, new):
                if lo <= value < hi:
                    return new_lo + (value - lo)
            raise AssertionError("unresolved data pointer in a rebindable plan")

        # the tcgen05 kernels need 16-byte aligned rows; check before anything is modified
        for site, mode in self.dense_sites:
            if mode != abi.DENSE_FP32 and translate(site.X) % 16:
                return False
        before = [bytes(array) for array, _ in self._host_tables]
        for struct in self._host_structs():
            self._walk_pointers(struct, lambda st, name, value: setattr(st, name, translate(value)))
        for (array, table), old_bytes in zip(self._host_tables, before):
            if bytes(array) != old_bytes:       # tables of latent-valued sites hold no data pointers
                table.copy_(torch.frombuffer(bytearray(bytes(array)), dtype=torch.uint8))
        self._sources = new
        if self.handle:
            desc = self._describe()
            stream = torch.cuda.current_stream(self.device).cuda_stream
            self.lib.call("mnf_plan_update", self.handle, C.byref(desc), stream)
        return True

    # ------------------------------------------------------------------------------------------
    # execution
    # ------------------------------------------------------------------------------------------
    @property
    def gpu_launches_per_step(self) -> int:
        """Kernels of this library the most recent :meth:`step` enqueued (counted by the library
        itself, include/mininf_b200.h::mnf_plan_launches)."""
        return self.launches_last_step

    def enable_peer_exchange(self, group: Any = None) -> None:
        """Sharded evaluations: every observed site is this rank's row shard, and the partial
        [S][1+D] accumulators of all ranks are combined by the engine itself over peer memory
        (``mnf_xrank_*``: push to every peer's inbox over NVLink, fixed-order sum in the tail
        kernel) - no NCCL call on the step path, so a sharded step is CUDA-graph capturable. The
        CUDA-IPC handles are exchanged once, here, through ``torch.distributed``."""
        import torch.distributed as dist
        if self.xrank is not None:
            return
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if world < 2:
            return
        with torch.cuda.device(self.device):
            handle = C.c_void_p()
            mine = C.create_string_buffer(abi.XRANK_HANDLE_BYTES)
            self.lib.call("mnf_xrank_create", world, rank, self.S * (self.D + 1), C.byref(handle), mine)
            gathered: List[Any] = [None] * world
            dist.all_gather_object(gathered, bytes(mine.raw), group=group)
            everyone = C.create_string_buffer(b"".join(gathered), world * abi.XRANK_HANDLE_BYTES)
            self.lib.call("mnf_xrank_connect", handle, everyone)
            dist.barrier(group)          # every rank has mapped every inbox before the first push
        self.xrank = handle
        self.buffers.xrank = handle

    def step(self, noise: Optional[torch.Tensor], seed: int, offset: int, with_entropy: bool = True,
             reduce_fn: Optional[Callable[[torch.Tensor], None]] = None,
             device_counter: bool = False) -> torch.Tensor:
        """Enqueue one ELBO evaluation on the current stream; returns the [1 + 2D] output buffer
        (loss, d loss / d p0, d loss / d p1). Parameters must already be in ``P0`` / ``P1``.
        ``device_counter``: the Philox call index is ``offset`` plus a device-resident counter that
        the last kernel of the step increments - for steps recorded into a CUDA graph, whose
        arguments are frozen at capture time. ``reduce_fn`` (e.g. an NCCL all-reduce) combines the
        accumulators of row-sharded ranks between the two halves of the step; with
        :meth:`enable_peer_exchange` the engine does that itself and the step is a single call."""
        lib, S, D = self.lib, self.S, self.D
        stream = 0 if self.dry_run else torch.cuda.current_stream(self.device).cuda_stream
        buffers = self.buffers
        buffers.step_counter = self.step_counter.data_ptr() if device_counter else None
        buffers.noise_in = None
        if noise is not None:
            self.noise_in.copy_(noise.reshape(S, D))
            buffers.noise_in = self.noise_in.data_ptr()
        for i, desc in enumerate(self.row_groups.values()):
            row = self._row_buffers[i]
            row.loc, row.scale, row.grad_loc, row.grad_scale, row.eps = \
                desc.loc, desc.scale, desc.grad_loc, desc.grad_scale, desc.eps
        flags = abi.STEP_ENTROPY if with_entropy else 0
        if self.record_sweep_events:
            # instrumented: the observed-site half through the phase-level entry points with CUDA
            # events around every sweep call (bench.py's roofline numerator), then the second half
            self._step_phases(seed, offset, with_entropy, stream)
            if reduce_fn is not None:
                reduce_fn(self.acc)
            lib.call("mnf_elbo_fwd_bwd", self.handle, C.byref(buffers), seed, offset, flags | abi.STEP_POST, stream)
        elif reduce_fn is not None:
            lib.call("mnf_elbo_fwd_bwd", self.handle, C.byref(buffers), seed, offset, flags | abi.STEP_PRE, stream)
            launched = self._launches()
            reduce_fn(self.acc)   # observed sites are row shards: sum the partial accumulators
            lib.call("mnf_elbo_fwd_bwd", self.handle, C.byref(buffers), seed, offset, flags | abi.STEP_POST, stream)
            self.launches_last_step = launched + self._launches()
            return self.out
        else:
            lib.call("mnf_elbo_fwd_bwd", self.handle, C.byref(buffers), seed, offset, flags | abi.STEP_ALL, stream)
        self.launches_last_step = self._launches()
        return self.out

    def _launches(self) -> int:
        count = C.c_int()
        self.lib.call("mnf_plan_launches", self.handle, C.byref(count))
        return int(count.value)

    def enqueue_sweeps(self) -> None:
        """Only the sweeps over the dense and scalar-link observed sites, with the particles of the
        last step (bench.py times a back-to-back series of these: a GPU-saturated loop sees the same
        power state as the timed region, which per-launch events in a gappy eager loop do not)."""
        lib, S, D = self.lib, self.S, self.D
        stream = torch.cuda.current_stream(self.device).cuda_stream
        status = self.status.data_ptr()
        for site, mode in self.dense_sites:
            lib.call("mnf_dense_sweep", C.byref(site), mode, self.z.data_ptr(), S, D, self.acc.data_ptr(),
                     self.workspace.data_ptr(), self.workspace_bytes, status, stream)
        for group in self.sweep_groups:
            lib.call("mnf_site_sweep", group, len(group), self.z.data_ptr(), S, D, self.acc.data_ptr(),
                     self.workspace.data_ptr(), self.workspace_bytes,
                     abi.SWEEP_CLOSED_FORM if self.closed_form else 0, status, stream)

    def _step_phases(self, seed: int, offset: int, with_entropy: bool, stream: int) -> None:
        """rsample and every sweep over observed sites through the phase-level C-ABI."""
        lib, S, D = self.lib, self.S, self.D
        counter = self.buffers.step_counter
        status = self.status.data_ptr()
        lib.call("mnf_rsample", self.latent_table.data_ptr(), len(self.latents), S, D, self.buffers.noise_in,
                 seed, offset, counter, self.z.data_ptr(), self.noise.data_ptr(), self.acc.data_ptr(),
                 status, stream)

        def timed(kind: str, call: Callable[[], None]) -> None:
            begin, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            begin.record(torch.cuda.current_stream(self.device))
            call()
            end.record(torch.cuda.current_stream(self.device))
            self.sweep_events.append((begin, end))
            self.sweep_event_kinds.append(kind)

        for site, mode in self.dense_sites:
            timed("dense", lambda: lib.call(
                "mnf_dense_sweep", C.byref(site), mode, self.z.data_ptr(), S, D, self.acc.data_ptr(),
                self.workspace.data_ptr(), self.workspace_bytes, status, stream))
        for group in self.sweep_groups:
            timed("site", lambda: lib.call(
                "mnf_site_sweep", group, len(group), self.z.data_ptr(), S, D, self.acc.data_ptr(),
                self.workspace.data_ptr(), self.workspace_bytes,
                abi.SWEEP_CLOSED_FORM if self.closed_form else 0, status, stream))
        if self.small_observed is not None:
            table, count, longest, _ = self.small_observed
            lib.call("mnf_small_sites", table.data_ptr(), count, longest, self.z.data_ptr(), S, D,
                     self.acc.data_ptr(), status, stream)
        for name, desc in self.row_groups.items():
            timed("rowlatent", lambda: lib.call(
                "mnf_rowlatent_sweep", C.byref(desc), self.z.data_ptr(), S, D, seed, offset, counter,
                int(with_entropy), self.acc.data_ptr(), self.workspace.data_ptr(), self.workspace_bytes,
                status, stream))


def status_message(bits: int) -> str:
    parts = []
    if bits & abi.ST_BAD_PARAM:
        parts.append("a distribution parameter left its constraint (e.g. a non-positive scale)")
    if bits & abi.ST_BAD_VALUE:
        parts.append("an observed value is not in the support of its distribution")
    if bits & abi.ST_NONFINITE:
        parts.append("the loss or a gradient is not finite")
    if bits & abi.ST_RANGE:
        parts.append("a design-matrix entry is NaN/Inf or at least 2^15: outside the range of the fp16-operand "
                     "kernel (use dense_precision='tf32')")
    if bits & abi.ST_XRANK_TIMEOUT:
        parts.append("a peer rank did not deliver its partial sums in time (sharded evaluation)")
    return "; ".join(parts)
