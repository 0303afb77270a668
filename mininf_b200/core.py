"""Host-side mirror of the reference's effect-handler core (``mininf/core.py``).

A model is a plain function that calls :func:`sample` for every random variable; what a call does
is decided by the active *tracer*. The names, signatures, exception types and message fragments
follow the reference so user models, ``condition``-ed data and tests carry over unchanged:

===========================  =====================================================
this module                  reference
===========================  =====================================================
``SingletonContextMixin``    mininf/core.py:19-83
``State``                    mininf/core.py:86-125
``TracerMixin``              mininf/core.py:128-189 (validation rules :142-189)
``SampleTracer``             mininf/core.py:192-204
``LogProbTracer``            mininf/core.py:207-277
``sample`` / ``condition``   mininf/core.py:300-328 / :331-387
``Value`` / ``value``        mininf/core.py:390-492
``batch`` / ``no_log_prob``  mininf/core.py:587-622 / :625-642
===========================  =====================================================

The ELBO hot path does NOT evaluate densities here: ``mininf_b200.nn.EvidenceLowerBoundLoss``
replays the model once under ``engine.trace.SiteTableTracer`` (a ``TracerMixin`` subclass) to
obtain a flat site table and hands it to the CUDA engine. ``LogProbTracer`` below is the
inspection tool of the reference API (per-site log-density tensors via ``torch.distributions``)
and is never used by the loss.
"""
from __future__ import annotations

import functools
import logging
from typing import Any, Callable, ClassVar, Dict, List, Tuple, Type, TypeVar
from unittest import mock

import torch
from torch.distributions import Distribution
from torch.distributions.constraints import Constraint

from .util import (OptionalSize, TensorDict, _format_dict_compact, _normalize_shape,
                   check_constraint, get_masked_data_with_dense_grad, is_masked, maybe_as_tensor)

LOGGER = logging.getLogger(__name__)
C = TypeVar("C", bound="SingletonContextMixin")


class SingletonContextMixin:
    """Context managers of which at most one per ``SINGLETON_KEY`` can be active at a time.

    The registry is process-global and contexts are not re-entrant, exactly like the reference.
    """
    INSTANCES: ClassVar[Dict[str, "SingletonContextMixin"]] = {}
    SINGLETON_KEY: ClassVar[str | None] = None

    @classmethod
    def _assert_singleton_key(cls) -> str:
        if not cls.SINGLETON_KEY:
            raise RuntimeError("Your class must define a singleton key.")
        return cls.SINGLETON_KEY

    def __enter__(self):
        key = self._assert_singleton_key()
        current = self.INSTANCES.get(key)
        if current is self:
            raise RuntimeError(f"Cannot reactivate {self} because it is already active.")
        if current is not None:
            raise RuntimeError(f"Cannot activate {self} with singleton key '{key}'; {current} is "
                               "already active.")
        self.INSTANCES[key] = self
        LOGGER.info("Activated %s as context for singleton key '%s'.", self, key)
        return self

    def __exit__(self, *_: Any) -> None:
        key = self._assert_singleton_key()
        current = self.INSTANCES.get(key)
        if current is None:
            raise RuntimeError(f"Cannot deactivate {self} with singleton key '{key}'; no context "
                               "is active.")
        if current is not self:
            raise RuntimeError(f"Cannot deactivate {self} with singleton key '{key}'; {current} is "
                               "active.")
        del self.INSTANCES[key]
        LOGGER.info("Deactivated %s as context for singleton key '%s'.", self, key)

    @classmethod
    def get_instance(cls: Type[C], strict: bool = False) -> C | None:
        """The active context of this class' key, ``None`` (or ``KeyError`` if ``strict``)."""
        key = cls._assert_singleton_key()
        current = cls.INSTANCES.get(key)
        if current is None:
            if strict:
                raise KeyError(f"No '{key}' context is active.")
            return None
        if not isinstance(current, cls):
            raise TypeError(f"Active context {current} is not an instance of {cls}.")
        return current


class State(Dict[str, Any], SingletonContextMixin):
    """The value table of one model execution: site name -> tensor."""
    SINGLETON_KEY = "state"

    def __repr__(self) -> str:
        return _format_dict_compact(self)

    def subset(self, *names: str) -> "State":
        """A new state holding (references to) the named entries only."""
        return State((name, self[name]) for name in names)


class batch(SingletonContextMixin):
    """Declare the full leading shape of the sites sampled inside the context; sites whose
    conditioned value is smaller along those dimensions get their log-density rescaled."""
    SINGLETON_KEY = "batch"

    def __init__(self, shape: torch.Size | int) -> None:
        self.shape = _normalize_shape(shape)

    @classmethod
    def get_shape(cls) -> torch.Size:
        active = cls.get_instance()
        return active.shape if active is not None else torch.Size()


class no_log_prob(SingletonContextMixin):
    """Sites sampled inside the context do not contribute to the joint log-density."""
    SINGLETON_KEY = "no_log_prob"


def batch_scale(value_shape: torch.Size, batch_shape: torch.Size) -> float:
    """Declared-over-actual element ratio of the batched leading dimensions
    (mininf/core.py:267-271)."""
    if not batch_shape:
        return 1.0
    return batch_shape.numel() / value_shape[:len(batch_shape)].numel()


class TracerMixin(SingletonContextMixin):
    """Base of the contexts that decide what :func:`sample` does."""
    SINGLETON_KEY = "tracer"

    def __init__(self, *args: Any, _validate_parameters: bool = True, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        self._validate_parameters = _validate_parameters

    def sample(self, state: State, name: str, distribution: Distribution,
               sample_shape: OptionalSize = None) -> torch.Tensor:
        raise NotImplementedError

    def _assert_valid_parameter(self, value: Any, name: str, distribution: Distribution,
                                sample_shape: OptionalSize) -> Any:
        """Type, rank, shape and support checks of a site value (mininf/core.py:142-189)."""
        if not self._validate_parameters:
            return value
        value = maybe_as_tensor(value)
        if not isinstance(value, torch.Tensor):
            raise TypeError(f"Expected a tensor for parameter '{name}' but got {type(value)}.")

        declared = batch.get_shape()
        sample_shape = _normalize_shape(sample_shape)
        iid_shape = sample_shape + distribution.batch_shape
        if len(declared) > len(iid_shape):
            raise ValueError(f"Declared batch shape {declared} for parameter '{name}' has more "
                             f"dimensions than the actual batch shape {iid_shape}.")

        expected = iid_shape + distribution.event_shape
        actual = value.shape
        mismatch = len(expected) != len(actual)
        if not mismatch:
            for dim, (want, got) in enumerate(zip(expected, actual)):
                batched = dim < len(declared)
                if got != want and not batched:
                    mismatch = True
                    break
                if got > want:
                    LOGGER.warning("Actual batch shape %s for parameter '%s' exceeds expected "
                                   "batch shape %s along dimension %d.", iid_shape, name, declared,
                                   dim)
        if mismatch:
            sizes = [f"{size}*" for size in declared] + \
                [str(size) for size in expected[len(declared):]]
            shown = ", ".join(sizes) + ("," if len(sizes) == 1 else "")
            raise ValueError(f"Expected shape ({shown}) for parameter '{name}' but got "
                             f"{tuple(actual)}.")

        in_support = check_constraint(distribution.support, value)
        if is_masked(in_support):
            # only unmasked entries count; an all-masked tensor is vacuously valid (reducing a
            # fully masked MaskedTensor to a bool raises inside torch.masked)
            in_support = in_support.get_data()[in_support.get_mask()]
        if not in_support.all():
            raise ValueError(f"Parameter '{name}' is not in the support of {distribution}.")
        return value


class SampleTracer(TracerMixin):
    """Draw every site that has no value yet (prior-predictive sampling)."""

    def sample(self, state: State, name: str, distribution: Distribution,
               sample_shape: OptionalSize = None) -> torch.Tensor:
        sample_shape = _normalize_shape(sample_shape)
        if state.get(name) is None:
            state[name] = distribution.sample(sample_shape)
        value = state[name]
        self._assert_valid_parameter(value, name, distribution, sample_shape)
        return value


class LogProbTracer(TracerMixin, Dict[str, Tuple[torch.Tensor, torch.Size]]):
    """Record ``(log_prob tensor, declared batch shape)`` per site with ``torch.distributions``.

    This is the reference's inspection API, kept for parity of the interface
    (tests/test_core.py:60-72, :203-231, :356-400). The ELBO loss does not go through it.
    """

    def sample(self, state: State, name: str, distribution: Distribution,
               sample_shape: OptionalSize = None) -> torch.Tensor:
        if isinstance(distribution, Value):
            current = state.get(name, distribution.value)
            self._assert_valid_parameter(current, name, distribution, sample_shape)
            return current
        if name in self:
            raise RuntimeError(f"Log probability has already been evaluated for '{name}'. Did you "
                               "call `sample` twice with the same variable name?")
        current = state.get(name)
        if current is None:
            raise ValueError(f"Cannot evaluate log probability; variable '{name}' is missing. Did "
                             "you forget to condition on observed data?")
        self._assert_valid_parameter(current, name, distribution, sample_shape)
        if no_log_prob.get_instance():
            return current

        if is_masked(current):
            if distribution._validate_args and \
                    not check_constraint(distribution.support, current).all():
                raise ValueError(f"Sample {current} is not in the support {distribution.support} "
                                 f"of distribution {distribution}.")
            with mock.patch.object(distribution, "_validate_args", False):
                dense = distribution.log_prob(get_masked_data_with_dense_grad(current))
            log_prob = torch.masked.as_masked_tensor(dense, current.get_mask())
        else:
            log_prob = distribution.log_prob(current)
        self[name] = (log_prob, batch.get_shape())
        return current

    def contribution(self, name: str) -> torch.Tensor:
        """The site's (masked, batch-rescaled) share of the joint log-density."""
        log_prob, declared = self[name]
        if is_masked(log_prob):
            if declared:
                raise ValueError("Batch dimensions are not supported for masked data.")
            return get_masked_data_with_dense_grad(log_prob)[log_prob.get_mask()].sum()
        if declared:
            return log_prob.sum() * declared.numel() / log_prob.shape[:len(declared)].numel()
        return log_prob.sum()

    @property
    def total(self) -> torch.Tensor:
        return sum(self.contribution(name) for name in self)  # type: ignore[return-value]

    def __repr__(self) -> str:
        return _format_dict_compact({name: entry[0] for name, entry in self.items()}, id(self),
                                    type(self).__name__)


def with_active_state(func: Callable) -> Callable:
    """Pass the active :class:`State` (or a fresh one for the duration of the call) as the first
    argument of ``func``."""
    @functools.wraps(func)
    def _wrapper(*args: Any, **kwargs: Any) -> Any:
        active = State.get_instance()
        if active is not None:
            return func(active, *args, **kwargs)
        with State() as fresh:
            return func(fresh, *args, **kwargs)
    return _wrapper


@with_active_state
def sample(state: State, name: str, distribution: Distribution,
           sample_shape: OptionalSize = None) -> torch.Tensor:
    """Declare random variable ``name`` with the given distribution and iid ``sample_shape``;
    returns its value of shape ``sample_shape + batch_shape + event_shape``."""
    tracer = TracerMixin.get_instance()
    if tracer is None:      # not `or`: an empty LogProbTracer is a falsy dict
        tracer = SampleTracer()
    return tracer.sample(state, name, distribution, sample_shape)


def condition(model: Callable, values: TensorDict | None = None, *, _strict: bool = True,
              **kwargs: torch.Tensor) -> Callable:
    """Pin sites of ``model`` to values. Keyword arguments override the dictionary; tensors are
    held by reference. With ``_strict`` a site may be conditioned only once, otherwise the first
    (innermost) conditioning wins."""
    merged: Dict[str, Any] = dict(values or {})
    merged.update(kwargs)
    pinned = {name: maybe_as_tensor(val) for name, val in merged.items()}

    @with_active_state
    @functools.wraps(model)
    def _wrapper(state: State, *args: Any, **kw: Any) -> Any:
        if _strict:
            conflict = set(state) & set(pinned)
            if conflict:
                raise ValueError(f"Cannot update state {state} because it already has parameters "
                                 f"{conflict}.")
        state.update(pinned)
        return model(*args, **kw)

    # introspection hooks for the engine's plan cache (not part of the reference API)
    _wrapper._mininf_model = model  # type: ignore[attr-defined]
    _wrapper._mininf_values = pinned  # type: ignore[attr-defined]
    _wrapper._mininf_strict = _strict  # type: ignore[attr-defined]
    return _wrapper


class Value(Distribution):
    """A constant or deterministic quantity: carries data, never contributes a log-density."""
    arg_constraints: Dict[str, Constraint] = {}

    def __init__(self, value: torch.Tensor | None = None, support: Constraint | None = None,
                 validate_args: bool | None = None) -> None:
        value = maybe_as_tensor(value)
        super().__init__(torch.Size(), torch.Size(), validate_args)
        self.value = value
        self._support = support or torch.distributions.constraints.real
        if value is not None and not check_constraint(self.support, value).all():
            raise ValueError(f"Default value is not in the specified support {self.support}.")

    @property
    def support(self) -> Constraint:  # type: ignore[override]
        return self._support

    def sample(self, sample_shape: Any = None) -> torch.Tensor:  # type: ignore[override]
        if self.value is None:
            raise ValueError("No default value given. Did you mean to specify one value by "
                             "conditioning?")
        return self.value

    def log_prob(self, value: Any) -> torch.Tensor:
        raise NotImplementedError("Values do not implement `log_prob` by design.")

    def __repr__(self) -> str:
        shown = {"value": self.value, "support": self.support}
        return "Value(" + ", ".join(f"{k}={v}" for k, v in shown.items() if v is not None) + ")"


def value(name: str, value: torch.Tensor | None = None, shape: torch.Size | None = None,
          support: Constraint | None = None, validate_args: bool | None = None) -> torch.Tensor:
    """Declare a deterministic site; without a default its shape must be given and the value is
    supplied later through :func:`condition`."""
    if shape is None and value is not None:
        value = torch.as_tensor(value)
        shape = value.shape
    return sample(name, Value(value, support, validate_args), shape)


def _assert_same_batch_size(state: Dict[str, torch.Tensor]) -> int:
    if not state:
        raise ValueError("Cannot check batch sizes because the state is empty.")
    by_size: Dict[int, List[str]] = {}
    for key, element in state.items():
        by_size.setdefault(element.shape[0], []).append(key)
    if len(by_size) != 1:
        raise ValueError(f"Inconsistent batch sizes: {by_size}")
    return next(iter(by_size))


def transpose_states(states: Any) -> Any:
    """State of batched tensors -> list of per-sample states, or the reverse."""
    if isinstance(states, dict):
        size = _assert_same_batch_size(states)
        return [State((key, element[i]) for key, element in states.items()) for i in range(size)]
    stacked: Dict[str, List[torch.Tensor]] = {}
    for single in states:
        for key, element in single.items():
            stacked.setdefault(key, []).append(maybe_as_tensor(element)[None])
    return State((key, torch.concatenate(parts)) for key, parts in stacked.items())


def broadcast_samples(model: Callable, states: State | None = None,
                      **params: torch.Tensor) -> State:
    """Apply ``model`` to every leading-dimension slice of the given samples and stack the results
    (posterior-predictive helper, mininf/core.py:548-584). CUDA samples are broadcast by the
    engine - the model is traced once and ``mnf_predictive`` evaluates / draws every site for all
    samples in one launch (engine/predictive.py); CPU samples take the reference's per-sample
    loop."""
    states = states or State()
    states.update(params)
    if states and any(isinstance(v, torch.Tensor) and v.is_cuda for v in states.values()):
        from .engine.predictive import broadcast_samples as _broadcast_on_device
        return _broadcast_on_device(model, states)
    results = []
    for single in transpose_states(states):
        with single:
            model()
        results.append(single)
    return transpose_states(results)
