"""Variational-inference layer behind the reference's ``mininf.nn`` interface.

``ParameterizedDistribution`` / ``FactorizedDistribution`` / ``ParameterizedFactorizedDistribution``
keep the reference's semantics (mininf/nn.py:29-187). ``EvidenceLowerBoundLoss`` keeps its call
signature and return contract (a 0-dim tensor with ``grad_fn`` whose ``backward`` fills the
unconstrained parameters, mininf/nn.py:190-228, tests/test_nn.py:50-66) but computes the Monte
Carlo ELBO and its gradient with the CUDA engine:

* the conditioned model is traced ONCE into a site table (``engine.trace``), cached per plan;
* every step enqueues the fused forward+backward kernels (``engine.plan``) for ``n_particles``
  reparameterised draws (the reference uses exactly one, mininf/nn.py:217; S particles are the
  mean of S reference evaluations);
* gradients come back for the CONSTRAINED parameters of the approximation, so stock autograd
  still chains through ``ParameterizedDistribution``'s transforms (mininf/nn.py:88-96).

There is no CPU or eager fallback: CPU tensors, unsupported distributions or link functions
raise.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Set, Tuple, Type

import torch
from torch import distributions, nn

from .core import batch, condition, no_log_prob
from .util import OptionalSize, TensorDict, _normalize_shape, maybe_as_tensor

DistributionDict = Dict[str, torch.distributions.Distribution]


def _is_identity_transform(transform: distributions.Transform) -> bool:
    return isinstance(transform, distributions.ComposeTransform) and not transform.parts


class ParameterizedDistribution(nn.Module):
    """A distribution type with trainable parameters stored in unconstrained space
    (``transform_to(constraint).inv``); calling the module rebuilds the distribution.

    ``_const`` names parameters to keep fixed, ``_clone`` protects caller-owned tensors of
    identity-constrained parameters from in-place optimiser updates.
    """

    def __init__(self, cls: Type[distributions.Distribution], *, _const: Set[str] | None = None,
                 _clone: bool = True, **parameters: Any) -> None:
        super().__init__()
        self.distribution_cls = cls
        fixed = _const or set()
        constraints = cls.arg_constraints
        self.distribution_constants: Dict[str, Any] = {}
        trainable: Dict[str, nn.Parameter] = {}
        for name, given in parameters.items():
            if name in fixed or name not in constraints:
                self.distribution_constants[name] = given
                continue
            given = maybe_as_tensor(given)
            transform = distributions.transform_to(constraints[name])
            if _is_identity_transform(transform) and _clone:
                raw = 1 * given
            else:
                raw = transform.inv(given)
            trainable[name] = nn.Parameter(raw)
        self.distribution_parameters = nn.ParameterDict(trainable)

    def forward(self) -> distributions.Distribution:
        constraints = self.distribution_cls.arg_constraints
        constrained = {}
        for name, raw in self.distribution_parameters.items():
            transform = distributions.transform_to(constraints[name])
            # `1 * raw` keeps the nn.Parameter itself out of the distribution
            constrained[name] = 1 * raw if _is_identity_transform(transform) else transform(raw)
        return self.distribution_cls(**constrained, **self.distribution_constants)


class FactorizedDistribution(DistributionDict):
    """Independent named factors: entropy adds up, sampling maps over the factors in order."""

    def entropy(self) -> torch.Tensor:
        return sum(factor.entropy().sum() for factor in self.values())  # type: ignore

    def rsample(self, sample_shape: OptionalSize = None) -> TensorDict:
        shape = _normalize_shape(sample_shape)
        return {name: factor.rsample(shape) for name, factor in self.items()}

    def sample(self, sample_shape: OptionalSize = None) -> TensorDict:
        shape = _normalize_shape(sample_shape)
        return {name: factor.sample(shape) for name, factor in self.items()}


class ParameterizedFactorizedDistribution(nn.ModuleDict):
    """``nn.ModuleDict`` of :class:`ParameterizedDistribution`; calling it yields a
    :class:`FactorizedDistribution`."""

    def __init__(self, arg: Dict[str, ParameterizedDistribution] | None = None,
                 **kwargs: ParameterizedDistribution) -> None:
        modules = dict(arg or {})
        modules.update(kwargs)
        super().__init__(modules)

    def forward(self) -> FactorizedDistribution:
        return FactorizedDistribution({name: module() for name, module in self.items()})


# -------------------------------------------------------------------------------------------------
# the engine-backed loss
# -------------------------------------------------------------------------------------------------
def _unwrap_conditioning(model: Callable) -> Tuple[Callable, List[Tuple[str, Any]]]:
    """Base callable and every conditioned (name, value) pair of a ``condition`` chain."""
    pinned: List[Tuple[str, Any]] = []
    while hasattr(model, "_mininf_values"):
        pinned.extend(model._mininf_values.items())
        model = model._mininf_model
    return model, pinned


def _constant_key(value: Any) -> Tuple:
    """Fingerprint of something a model function captured: plain Python constants by value (a
    hyper-parameter changed between steps must retrace - the reference re-runs the model every
    step, ``mininf/nn.py:223-225``), everything else by identity (tensors are read in place)."""
    if value is None or isinstance(value, (bool, int, float, complex, str, bytes)):
        return ("v", type(value).__name__, value)
    if isinstance(value, (tuple, list)) and len(value) <= 16 and \
            all(v is None or isinstance(v, (bool, int, float, str)) for v in value):
        return ("s", type(value).__name__, tuple(value))
    return ("o", id(value))


def _captured_key(function: Any) -> Tuple:
    """Closure cells, defaults and the plain-constant globals a function names: what the trace of
    a model depends on besides its conditioned values."""
    code = getattr(function, "__code__", None)
    if code is None:
        return ()
    captured = []
    for cell in getattr(function, "__closure__", None) or ():
        try:
            captured.append(_constant_key(cell.cell_contents))
        except ValueError:            # empty cell
            captured.append(("e",))
    for default in getattr(function, "__defaults__", None) or ():
        captured.append(_constant_key(default))
    for name, default in sorted((getattr(function, "__kwdefaults__", None) or {}).items()):
        captured.append((name,) + _constant_key(default))
    namespace = getattr(function, "__globals__", {})
    for name in code.co_names:
        value = namespace.get(name, _callable_key)      # sentinel: not a module-level name
        if value is None or isinstance(value, (bool, int, float, complex, str)):
            captured.append((name,) + _constant_key(value))
    return tuple(captured)


def _callable_key(model: Callable) -> Tuple:
    """Identity of a model callable that survives re-creation of thin wrappers: a bound method
    (``obj.model``) is a new object on every attribute access and a ``functools.partial`` is often
    rebuilt per step, so keying the plan cache on ``id(model)`` would retrace every step. Python
    constants the function captured (closure cells, defaults, module-level scalars it names) are
    part of the key by VALUE, so changing one between steps traces the model again."""
    import functools
    if isinstance(model, functools.partial):
        frozen = tuple(_constant_key(a) for a in model.args) + \
            tuple((k,) + _constant_key(v) for k, v in sorted(model.keywords.items()))
        return ("partial", _callable_key(model.func), frozen)
    if hasattr(model, "__func__") and hasattr(model, "__self__"):
        return ("method", id(model.__func__), id(model.__self__), _captured_key(model.__func__))
    code = getattr(model, "__code__", None)
    if code is not None:
        # a plain function or lambda: the code object plus what it captured. A lambda re-created in every
        # loop iteration (`condition(lambda: model(x), ...)`) keeps its code object, so it does not retrace.
        return ("function", id(code), _captured_key(model))
    return ("callable", id(model))


def _layout_key(value: Any) -> Tuple:
    """What a plan depends on structurally: the layout of a conditioned value, not its address."""
    if isinstance(value, torch.masked.MaskedTensor):
        return ("masked",) + _layout_key(value.get_data()) + _layout_key(value.get_mask())
    if isinstance(value, torch.Tensor):
        return (tuple(value.shape), tuple(value.stride()), str(value.dtype), str(value.device))
    return (repr(value),)


def _leaves(values: List[Any]) -> List[torch.Tensor]:
    """Tensors behind the conditioned values, in a fixed order (data then mask for masked ones)."""
    out: List[torch.Tensor] = []
    for value in values:
        if isinstance(value, torch.masked.MaskedTensor):
            out.extend([value.get_data(), value.get_mask()])
        elif isinstance(value, torch.Tensor):
            out.append(value)
    return out


class _EngineFunction(torch.autograd.Function):
    """One fused forward+backward evaluation; inputs are the constrained parameter tensors."""

    @staticmethod
    def forward(ctx, loss_module, plan, noise, row_noise, seed, offset, reduce_fn, with_entropy, capturing,
                *params):
        with torch.no_grad():
            big_grads = []
            for spec, p0, p1 in zip(plan.all_latents, params[0::2], params[1::2]):
                if spec.row_latent:
                    # row latents are read in place; their gradients are written by the kernel
                    loc, scale = p0.contiguous(), p1.contiguous()
                    g_loc, g_scale = torch.empty_like(loc), torch.empty_like(scale)
                    desc = plan.row_groups[spec.name]
                    desc.loc, desc.scale = loc.data_ptr(), scale.data_ptr()
                    desc.grad_loc, desc.grad_scale = g_loc.data_ptr(), g_scale.data_ptr()
                    given = row_noise.get(spec.name) if row_noise else None
                    desc.eps = given.data_ptr() if given is not None else None
                    big_grads += [loc, scale, g_loc, g_scale]
                else:
                    lo, hi = spec.offset, spec.offset + spec.numel
                    plan.P0[lo:hi].copy_(p0.reshape(-1))
                    plan.P1[lo:hi].copy_(p1.reshape(-1))
            out = plan.step(noise, seed, offset, with_entropy=with_entropy, reduce_fn=reduce_fn,
                            device_counter=capturing)
            saved = out.clone()
        ctx.plan = plan
        ctx.big_grads = big_grads
        ctx.save_for_backward(saved)
        if capturing:
            loss_module._graphed_plans[id(plan)] = plan     # status is read by synchronize()
        else:
            loss_module._schedule_status_check(plan)
        return saved[0].clone()

    @staticmethod
    def backward(ctx, grad_output):
        (saved,) = ctx.saved_tensors
        plan = ctx.plan
        grads: List[Optional[torch.Tensor]] = [None] * 9
        big = iter(ctx.big_grads)
        for spec in plan.all_latents:
            if spec.row_latent:
                _, _, g_loc, g_scale = next(big), next(big), next(big), next(big)
                grads.append(g_loc.reshape(spec.shape) * grad_output)
                grads.append(g_scale.reshape(spec.shape) * grad_output)
                continue
            lo, hi = 1 + spec.offset, 1 + spec.offset + spec.numel
            grads.append((saved[lo:hi] * grad_output).reshape(spec.shape))
            grads.append((saved[plan.D + lo:plan.D + hi] * grad_output).reshape(spec.shape))
        return tuple(grads)


class EvidenceLowerBoundLoss(nn.Module):
    """Negative Monte Carlo ELBO of ``model`` under ``approximation``.

    Args (all optional, all beyond the reference's zero-argument constructor):
        n_particles: reparameterised draws averaged per evaluation (default 1 = the reference).
        dense_precision: operand format of dense-link sites (``X @ theta``) on the tensor cores.
            ``"auto"``: fp16 operands (11-bit significand like TF32, half the tensor-pipe time)
            for p = 64 / S <= 64 when the design matrix fits fp16's range, else TF32 operands
            where a tcgen05 kernel covers the shape, else fp32 SIMT; ``"f16"`` / ``"tf32"`` require
            that format, ``"fp32"`` forces the exact SIMT kernel. In every tensor-core mode theta
            enters as hi + lo pairs (22 bits) and accumulation is fp32.
        cache: reuse the traced plan while the model, conditioned tensors and approximation
            structure are unchanged (the reference re-runs the model every step).
        process_group: ``True`` / a ``torch.distributed`` group to treat every observed site as
            this rank's row shard and combine the partial sums of all ranks once per step.
        reduce: how sharded ranks combine their partial sums. ``"peer"`` (default): the engine
            pushes its [S][1+D] accumulator into every peer's inbox over NVLink and each rank adds
            the inboxes in rank order inside the step's last kernel - no NCCL call, bit-identical
            totals on every rank, CUDA-graph capturable. ``"nccl"``: one ``all_reduce`` between
            the two halves of the step. Every rank must evaluate the same sequence of losses; the
            Philox seed is taken from rank 0 when the exchange is set up.
        closed_form: let sites whose log-density is a closed form of data-only sufficient
            statistics skip the per-(particle, observation) sweep (Gram statistics of a Normal
            dense site with p <= 64; six sums / Chebyshev moments for scalar-link Normal / Poisson
            sites). Off (default), every site is evaluated once per particle and observation -
            the black-box estimator the reference implements.
        check: ``"lazy"`` validates the device status word of the previous step at the next call
            (no extra synchronisation), ``"sync"`` synchronises every call, ``"off"`` never.
    """

    def __init__(self, n_particles: int = 1, *, dense_precision: str = "auto", cache: bool = True,
                 process_group: Any = None, check: str = "lazy", closed_form: bool = False,
                 reduce: str = "peer") -> None:
        super().__init__()
        if dense_precision not in ("auto", "f16", "tf32", "fp32"):
            raise ValueError("dense_precision must be 'auto', 'f16', 'tf32' or 'fp32'")
        if check not in ("lazy", "sync", "off"):
            raise ValueError("check must be 'lazy', 'sync' or 'off'")
        if reduce not in ("peer", "nccl"):
            raise ValueError("reduce must be 'peer' or 'nccl'")
        self.n_particles = int(n_particles)
        self.dense_precision = dense_precision
        self.cache = cache
        self.process_group = process_group
        self.check = check
        self.closed_form = bool(closed_form)
        self.reduce = reduce
        self._shared_seed: Optional[int] = None
        self._plans: Dict[Tuple, Any] = {}
        self._pending: List[Tuple[Any, torch.Tensor, torch.cuda.Event]] = []
        self._free_slots: List[torch.Tensor] = []
        self._graphed_plans: Dict[int, Any] = {}
        self._calls = 0
        self.last_plan = None

    # -- status word ------------------------------------------------------------------------
    def _schedule_status_check(self, plan: Any) -> None:
        if self.check == "off":
            return
        if self.check == "sync":
            bits = int(plan.status.item())
            if bits:
                plan.status.zero_()          # report once; the plan stays usable for the next batch
            self._raise_for_status(bits)
            return
        if len(self._pending) >= 8:          # bounded backlog: settle the oldest check first
            _, host, event = self._pending.pop(0)
            event.synchronize()
            self._raise_for_status(int(host.item()))
        else:
            # pinned slots are recycled: allocating pinned memory every step is slow and may
            # synchronise the device
            host = self._free_slots.pop() if self._free_slots else \
                torch.empty(1, dtype=torch.int32, pin_memory=True)
            event = torch.cuda.Event()
        host.copy_(plan.status, non_blocking=True)
        event.record(torch.cuda.current_stream(plan.device))
        self._pending.append((plan, host, event))

    def _drain_status(self, block: bool = False) -> None:
        remaining, flagged = [], 0
        for plan, host, event in self._pending:
            if block:
                event.synchronize()
            if event.query():
                bits = int(host.item())
                if bits:
                    plan.status.zero_()      # report once
                flagged |= bits
                self._free_slots.append(host)
            else:
                remaining.append((plan, host, event))
        self._pending = remaining
        self._raise_for_status(flagged)

    @staticmethod
    def _raise_for_status(bits: int) -> None:
        if bits:
            from .engine.plan import status_message
            raise ValueError("the ELBO engine flagged invalid values: " + status_message(bits))

    def synchronize(self) -> None:
        """Wait for outstanding evaluations and raise if any of them flagged invalid values
        (including steps replayed from a CUDA graph, whose status word is only read here)."""
        self._drain_status(block=True)
        flagged = 0
        for plan in self._graphed_plans.values():
            bits = int(plan.status.item())
            if bits:
                plan.status.zero_()
            flagged |= bits
        self._raise_for_status(flagged)

    # -- tracing ----------------------------------------------------------------------------
    def _build_plan(self, model: Callable, approximation: DistributionDict) -> Any:
        from .engine.plan import Plan, assign_offsets, latent_parameters, row_latent_names, slope_groups
        from .engine.trace import Affine, LatentRef, LinkTensor, SiteTableTracer

        entries = []
        draws: Dict[str, torch.Tensor] = {}
        device = None
        for name, factor in approximation.items():
            family, p0, _ = latent_parameters(factor)
            if factor.event_shape:
                raise NotImplementedError("event-shaped approximations are not supported")
            shape = factor.batch_shape
            numel = max(shape.numel(), 1)
            entries.append((name, family, shape))
            device = device or p0.device
            with torch.no_grad():
                draw = factor.sample()
            # one-element latents are referenced as scalars so they may broadcast over any site
            ref = LatentRef(name, 0) if numel == 1 else LatentRef(name)
            draws[name] = LinkTensor.wrap(draw, Affine(a_lat=ref))
        if device is None:
            raise ValueError("the approximation has no factors")
        if device.type != "cuda":
            raise RuntimeError("the mininf_b200 ELBO engine runs on CUDA tensors only (there is "
                               "no CPU fallback); got approximation parameters on " + str(device))
        with SiteTableTracer() as tracer:
            condition(model, **draws)()
        # latent draws the model never scores still take part in the entropy term; that mirrors
        # the reference, where `condition` silently accepts unused names
        specs = assign_offsets(entries, row_latent_names(tracer.sites), slope_groups(tracer.sites))
        return Plan(tracer.sites, specs, self.n_particles, device,
                    dense_mode=self.dense_precision, closed_form=self.closed_form)

    def _plan_for(self, model: Callable, approximation: DistributionDict) -> Any:
        """Cached plan for this model, data layout and approximation structure. The same tensors
        (address and version counter) reuse the plan as it is; new tensors of the same layout -
        the next minibatch - only have their pointers patched (:meth:`Plan.rebind`) when the plan
        reads them in place; anything else is traced and lowered again."""
        if not self.cache:
            return self._build_plan(model, approximation)
        base, pinned = _unwrap_conditioning(model)
        values = [val for _, val in pinned]
        leaves = _leaves(values)
        key = (_callable_key(base), tuple((name, _layout_key(val)) for name, val in pinned),
               tuple((name, type(factor).__name__, tuple(factor.batch_shape))
                     for name, factor in approximation.items()),
               self.n_particles, self.dense_precision, self.closed_form,
               # contexts opened AROUND the loss call shape the trace like the ones inside the model
               tuple(batch.get_shape()), no_log_prob.get_instance() is not None)
        bound = tuple((leaf.data_ptr(), leaf._version) for leaf in leaves)
        plan = self._plans.get(key)
        if plan is not None:
            if plan._bound == bound:
                return plan
            if plan.rebind(leaves):
                plan._bound, plan._pin = bound, (base, values)
                return plan
            del self._plans[key]
        if len(self._plans) >= 8:
            self._plans.pop(next(iter(self._plans)))
        plan = self._build_plan(model, approximation)
        plan.bind_sources(leaves)
        plan._bound, plan._pin = bound, (base, values)   # keep ids / data pointers valid
        self._plans[key] = plan
        return plan

    # -- noise ------------------------------------------------------------------------------
    def _noise(self, plan: Any, given: Optional[Dict[str, torch.Tensor]]) -> Optional[torch.Tensor]:
        """External reparameterisation noise [S, D] for the packed latents, or ``None`` to let
        the kernel draw everything with Philox. Convention per family: eps for Normal, the
        standard-gamma draw for Gamma, the drawn value for Beta."""
        if given is None:
            return None
        missing = [spec.name for spec in plan.latents if spec.name not in given]
        if missing:
            raise ValueError(f"_noise must cover every packed latent site; missing {missing}")
        S = plan.S
        noise = torch.empty(S, plan.D, device=plan.device, dtype=torch.float32)
        with torch.no_grad():
            for spec in plan.latents:
                block = noise[:, spec.offset:spec.offset + spec.numel]
                block.copy_(given[spec.name].to(plan.device, torch.float32).reshape(S, spec.numel))
        return noise

    # -- forward ----------------------------------------------------------------------------
    def forward(self, model: Callable,
                approximation: torch.distributions.Distribution | DistributionDict, *,
                _noise: Optional[Dict[str, torch.Tensor]] = None,
                _with_entropy: bool = True) -> torch.Tensor:
        if not isinstance(approximation, dict):
            raise TypeError("Expected a distribution which samples dictionaries of tensors but got "
                            f"a sample of type {type(approximation)}")
        # inside torch.cuda.graph(...) nothing may query events or touch pinned memory: the step is
        # recorded with a device-side Philox call index and its status word is read by synchronize()
        capturing = torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()
        if not capturing:
            self._drain_status()
        plan = self._plan_for(model, approximation)
        self.last_plan = plan
        from .engine.plan import latent_parameters
        params: List[torch.Tensor] = []
        for spec in plan.all_latents:
            _, p0, p1 = latent_parameters(approximation[spec.name])
            shape = spec.shape if len(spec.shape) else torch.Size([])
            params.append(p0.to(torch.float32).expand(shape))
            params.append(p1.to(torch.float32).expand(shape))
        noise = self._noise(plan, _noise)
        # external noise of row latents [S, n, p] (parity tests); otherwise in-kernel Philox
        row_noise = {spec.name: _noise[spec.name].to(plan.device, torch.float32).contiguous()
                     for spec in plan.row_latents if _noise is not None and spec.name in _noise}
        seed = torch.cuda.default_generators[plan.device.index or 0].initial_seed() \
            if plan.device.index is not None else torch.cuda.initial_seed()
        self._calls += 1
        reduce_fn = None
        if self.process_group is not None:
            import torch.distributed as dist
            group = None if self.process_group is True else self.process_group
            if self._shared_seed is None:
                # every rank must draw the same z: the partial sums of all ranks are added, and the
                # priors / entropy / optimiser update are computed redundantly on every rank. Rank
                # 0's generator seed and call count become everyone's (once, outside any capture).
                if capturing:
                    raise RuntimeError("evaluate a sharded loss once eagerly before capturing it in a CUDA graph")
                shared = [int(seed), self._calls]
                dist.broadcast_object_list(shared, src=dist.get_global_rank(group, 0) if group is not None else 0,
                                           group=group)
                self._shared_seed, self._calls = int(shared[0]), int(shared[1])
            seed = self._shared_seed
            if self.reduce == "peer":
                if plan.xrank is None:
                    if capturing:
                        raise RuntimeError("evaluate a sharded loss once eagerly before capturing it in a CUDA graph")
                    plan.enable_peer_exchange(group)
            else:
                reduce_fn = lambda acc: dist.all_reduce(acc, group=group)  # noqa: E731
        offset = self._calls + (1 << 40 if capturing else 0)    # replays count on from here on the device
        return _EngineFunction.apply(self, plan, noise, row_noise, int(seed) & (2 ** 63 - 1), offset,
                                     reduce_fn, bool(_with_entropy), capturing, *params)


class GraphedStep:
    """One SVI step - ``zero_grad``, loss, ``backward``, ``optimizer.step`` - recorded once into a
    CUDA graph and replayed with a single launch (``torch.cuda.graph``; the reference's loop of
    README.md:66-69 issues ~20 small kernels around the sweeps, which costs 5-15 % of a step).

    ``approximation`` is a callable returning the approximation dictionary, e.g.
    ``lambda: {"theta": q()}``; the optimizer must be capturable (``Adam(..., capturable=True)``).
    Conditioned tensors are read in place on every replay: refill them with ``copy_`` to feed the
    next minibatch. Every replay draws fresh Philox noise (the call index lives on the device).
    ``loss`` holds the value of the most recent replay; ``loss_module.synchronize()`` reports
    invalid values flagged by the kernels. A sharded loss (``process_group=...``, peer exchange)
    is capturable too: evaluate it once eagerly first (the warm-up steps here do).
    :class:`FusedSVIStep` is the fully native variant (transforms and Adam inside the engine).
    """

    def __init__(self, loss_module: "EvidenceLowerBoundLoss", model: Callable, approximation: Callable[[], Any],
                 optimizer: torch.optim.Optimizer, warmup: int = 3) -> None:
        self.loss_module, self.model, self.approximation, self.optimizer = loss_module, model, approximation, optimizer
        for group in optimizer.param_groups:
            if "capturable" in group and not group["capturable"]:
                raise ValueError("GraphedStep needs a capturable optimizer, e.g. torch.optim.Adam(..., capturable=True)")
        # eager warm-up on a side stream (builds the plan, optimizer state and kernel attributes)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(int(warmup), 1)):
                self._step()
        torch.cuda.current_stream().wait_stream(side)
        loss_module.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        optimizer.zero_grad(set_to_none=True)
        # torch.distributions validates constructor arguments with a host round trip
        # (`if not valid.all()`), which a capture cannot contain; the warm-up steps above ran
        # with validation on, and the kernels keep checking scales and supports on the device
        validate = torch.distributions.Distribution._validate_args
        torch.distributions.Distribution.set_default_validate_args(False)
        try:
            with torch.cuda.graph(self.graph):
                self.loss = self._step(zero_grad=False).detach()
        finally:
            torch.distributions.Distribution.set_default_validate_args(validate)

    def _step(self, zero_grad: bool = True) -> torch.Tensor:
        if zero_grad:
            self.optimizer.zero_grad(set_to_none=True)
        loss = self.loss_module(self.model, self.approximation())
        loss.backward()
        self.optimizer.step()
        return loss

    def __call__(self) -> torch.Tensor:
        self.graph.replay()
        return self.loss


_PARAMETER_NAMES = {distributions.Normal: ("loc", "scale"), distributions.Gamma: ("concentration", "rate"),
                    distributions.Beta: ("concentration1", "concentration0")}


def _transform_code(transform: distributions.Transform) -> Optional[int]:
    """MNF_T_* code of a constraint transform, recognised by what it computes (torch wraps the
    transform of `positive` as ComposeTransform([ExpTransform(), AffineTransform(0, 1)]))."""
    from .engine import abi
    probe = torch.tensor([-1.5, 0.0, 0.75, 2.0], dtype=torch.float64)
    try:
        image = transform(probe)
    except Exception:  # noqa: BLE001  shape-changing transforms (simplex, cholesky, ...)
        return None
    if image.shape != probe.shape:
        return None
    if torch.equal(image, probe):
        return abi.T_ID
    if torch.allclose(image, probe.exp(), rtol=1e-12, atol=0.0):
        return abi.T_EXP
    return None


class FusedSVIStep:
    """The whole SVI step of README.md:63-69 - ``zero_grad``, loss, ``backward``,
    ``optimizer.step()`` - as ONE native call (``mnf_svi_step``, include/mininf_b200.h) that
    enqueues 4-6 kernels: parameter transforms + reparameterised draws, the sweeps over the
    observed sites, [the peer exchange of a sharded run], and a single tail kernel with the prior
    sites, the entropy, the pathwise gradients, the chain rule through
    ``ParameterizedDistribution``'s transforms (mininf/nn.py:88-96) and Adam
    (``torch.optim.Adam`` semantics: no weight decay, no amsgrad). With ``graph=True`` the call is
    recorded once into a CUDA graph and replayed.

    ``approximation`` maps latent names to :class:`ParameterizedDistribution` modules (or is a
    :class:`ParameterizedFactorizedDistribution`). Their ``nn.Parameter`` objects are re-pointed at
    views of one packed buffer that the kernels update in place, so the modules keep working as
    usual (``module()``, ``state_dict()``). Supported: Normal / Gamma / Beta factors whose
    parameters are real (identity transform) or positive (exp transform); per-observation
    ("row") latents are not part of the fused step.
    """

    def __init__(self, loss_module: "EvidenceLowerBoundLoss", model: Callable,
                 approximation: Dict[str, ParameterizedDistribution] | ParameterizedFactorizedDistribution,
                 lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 graph: bool = True, share_with: Optional["FusedSVIStep"] = None) -> None:
        """``share_with``: another step over the same approximation whose parameters and Adam state
        this one updates too - one step object per resident minibatch buffer (conditioned tensors
        are baked into a recorded step by address), all training the same parameters."""
        import ctypes as C
        from .engine import abi
        self._C, self._abi = C, abi
        self.loss_module, self.model = loss_module, model
        self.modules: Dict[str, ParameterizedDistribution] = dict(approximation.items())
        with torch.no_grad():
            factors = {name: module() for name, module in self.modules.items()}
        # one eager evaluation builds the plan (and, for a sharded loss, the peer exchange)
        loss_module(model, factors)
        loss_module.synchronize()
        plan = self.plan = loss_module.last_plan
        if plan.row_latents and (plan.S > 32 or share_with is not None):
            raise NotImplementedError("FusedSVIStep trains per-observation latents inside the row-latent sweep for "
                                      "single-pass sweeps only (at most 32 particles, one resident data set); use "
                                      "GraphedStep with a torch optimizer otherwise")
        D, device = plan.D, plan.device
        raw = torch.zeros(2 * D, device=device)
        codes = torch.zeros(2 * D, dtype=torch.uint8)
        if share_with is not None and share_with.raw.numel() != 2 * D:
            raise ValueError("share_with: the two steps train different approximations")
        for spec in (plan.latents if share_with is None else []):
            module = self.modules[spec.name]
            names = _PARAMETER_NAMES.get(module.distribution_cls)
            if names is None:
                raise NotImplementedError(f"FusedSVIStep: unsupported approximation family for '{spec.name}'")
            for half, name in enumerate(names):
                lo = half * D + spec.offset
                hi = lo + spec.numel
                if name in module.distribution_parameters:
                    parameter = module.distribution_parameters[name]
                    transform = distributions.transform_to(module.distribution_cls.arg_constraints[name])
                    code = _transform_code(transform)
                    if code is None:
                        raise NotImplementedError(f"FusedSVIStep: transform {transform} of '{spec.name}.{name}'")
                    if parameter.numel() != spec.numel:
                        raise NotImplementedError(f"FusedSVIStep: '{spec.name}.{name}' is broadcast over the site; "
                                                  "give it the site's shape")
                    raw[lo:hi].copy_(parameter.detach().reshape(-1))
                    # the packed buffer becomes the parameter's storage: in-place kernel updates are
                    # visible through the module
                    parameter.data = raw[lo:hi].view(parameter.shape)
                    codes[lo:hi] = code
                else:
                    constant = torch.as_tensor(module.distribution_constants[name], dtype=torch.float32)
                    raw[lo:hi].copy_(constant.to(device).expand(spec.shape if len(spec.shape) else (1,)).reshape(-1))
                    codes[lo:hi] = abi.T_ID | abi.T_FROZEN
        if share_with is None:
            self.raw, self.codes = raw, codes.to(device)
            self.m, self.v = torch.zeros_like(raw), torch.zeros_like(raw)
            self.steps = torch.zeros(1, dtype=torch.int64, device=device)
        else:
            self.raw, self.codes, self.m, self.v, self.steps = \
                share_with.raw, share_with.codes, share_with.m, share_with.v, share_with.steps
        raw = self.raw
        self.adam = abi.Adam(lr=lr, beta1=betas[0], beta2=betas[1], eps=eps, raw=raw.data_ptr(),
                             transform=self.codes.data_ptr(), m=self.m.data_ptr(), v=self.v.data_ptr(),
                             constrained=plan.P.data_ptr(), step=self.steps.data_ptr())
        # per-observation ("row") latents: the N*p location / log-scale parameters are read, updated
        # and written back by the row-latent sweep itself (csrc/rowlatent.cuh, RowAdam)
        self._row_state: List[torch.Tensor] = []
        for index, spec in enumerate(plan.row_latents):
            module = self.modules[spec.name]
            if module.distribution_cls is not distributions.Normal or \
                    set(module.distribution_parameters) != {"loc", "scale"}:
                raise NotImplementedError(f"FusedSVIStep: row latent '{spec.name}' needs a Normal approximation with "
                                          "trainable loc and scale")
            loc, raw_scale = module.distribution_parameters["loc"], module.distribution_parameters["scale"]
            for parameter in (loc, raw_scale):
                if tuple(parameter.shape) != tuple(spec.shape) or not parameter.is_contiguous():
                    raise NotImplementedError(f"FusedSVIStep: parameters of '{spec.name}' must be contiguous and have "
                                              "the site's shape")
            moments = [torch.zeros_like(loc) for _ in range(4)]
            self._row_state += moments
            row = plan._row_buffers[index]
            row.loc, row.loc_rw, row.raw_scale = loc.data_ptr(), loc.data_ptr(), raw_scale.data_ptr()
            row.scale, row.grad_loc, row.grad_scale, row.eps = None, None, None, None
            row.m_loc, row.v_loc, row.m_scale, row.v_scale = (t.data_ptr() for t in moments)
        self.loss = plan.out[0]
        generator = torch.cuda.default_generators[device.index or 0]
        self._seed = (loss_module._shared_seed if loss_module._shared_seed is not None
                      else generator.initial_seed()) & (2 ** 63 - 1)
        FusedSVIStep._instances += 1       # every step object draws from its own Philox call range
        self._offset = (1 << 41) + (FusedSVIStep._instances << 34) + loss_module._calls
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        if graph:
            side = torch.cuda.Stream(device)
            side.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(side):
                for _ in range(2):
                    self._enqueue()
            torch.cuda.current_stream(device).wait_stream(side)
            torch.cuda.synchronize(device)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._enqueue()
        loss_module._graphed_plans[id(plan)] = plan       # status word is read by synchronize()

    _instances = 0

    def _enqueue(self) -> None:
        plan, C = self.plan, self._C
        stream = torch.cuda.current_stream(plan.device).cuda_stream
        plan.buffers.noise_in = None                       # an eager loss call on the same plan resets these
        plan.buffers.step_counter = plan.step_counter.data_ptr()
        plan.lib.call("mnf_svi_step", plan.handle, C.byref(plan.buffers), C.byref(self.adam), self._seed,
                      self._offset, self._abi.STEP_ENTROPY | self._abi.STEP_ALL, stream)
        plan.launches_last_step = plan._launches()

    def __call__(self) -> torch.Tensor:
        """Run one step; returns the (device-resident) loss of the parameters BEFORE the update."""
        if self.graph is not None:
            self.graph.replay()
        else:
            self._enqueue()
        return self.loss

    @property
    def kernels_per_step(self) -> int:
        return self.plan.launches_last_step


class LogLikelihoodLoss(nn.Module):
    """Negative joint log-density at fixed parameter values (mininf/nn.py:231-257).

    CUDA tensors run on the engine: the same site table and sweeps as the ELBO with one
    "particle" pinned to the given values (zero noise, no entropy term), so the gradient with
    respect to ``parameters`` comes from the fused kernels. CPU tensors raise, exactly like
    :class:`EvidenceLowerBoundLoss`: there is no CPU fallback.
    """

    def __init__(self, *, dense_precision: str = "auto") -> None:
        super().__init__()
        self._engine = EvidenceLowerBoundLoss(1, dense_precision=dense_precision)

    def forward(self, model: Callable, parameters: TensorDict) -> torch.Tensor:
        values = {name: maybe_as_tensor(value) for name, value in parameters.items()}
        if not values or not all(isinstance(v, torch.Tensor) and v.is_cuda for v in values.values()):
            raise RuntimeError("the mininf_b200 engine runs on CUDA tensors only (there is no CPU fallback); "
                               "LogLikelihoodLoss got parameter values that are not CUDA tensors")
        # a unit-scale Normal around each value with zero noise draws exactly the value, and
        # d loss / d loc is the gradient with respect to the value
        point = {name: distributions.Normal(value, torch.ones_like(value)) for name, value in values.items()}
        zeros = {name: torch.zeros((1,) + tuple(value.shape), device=value.device) for name, value in values.items()}
        return self._engine(model, point, _noise=zeros, _with_entropy=False)
