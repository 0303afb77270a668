"""ORACLE - test infrastructure, not product code.

A CPU restatement of the reference's effect handlers, just enough to evaluate the joint
log-density of a model the way ``mininf.core.LogProbTracer`` does. Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this package; the
product (``mininf_b200``) never does.

Models used with the oracle take the API namespace as their first argument::

    def model(m):
        theta = m.sample("theta", Normal(0, 1), 3)

so the very same function runs under the reference (``m = mininf``), the product
(``m = mininf_b200``) and this module (``m = oracle.handlers``).

Restated from /root/reference (tillahoffmann/mininf):
  sample dispatch + LogProbTracer.sample      mininf/core.py:211-245, 300-328
  support / shape validation                  mininf/core.py:142-189, mininf/util.py:43-66
  contribution / total (mask, batch scaling)  mininf/core.py:247-273
  condition (kwargs win, strict conflicts)    mininf/core.py:331-387
  batch / no_log_prob / value                 mininf/core.py:390-492, 587-642
Parity pinned by tests/test_oracle_golden.py against fixtures generated from the reference itself
(tests/golden/make_golden.py).
"""
from __future__ import annotations

import contextlib
from typing import Any, Callable, Dict, Iterator, List, Optional, Tuple

import torch
from torch.distributions import Distribution, constraints


class _Frame:
    """One evaluation: the value table plus what has been scored so far."""

    def __init__(self, values: Dict[str, Any], validate: bool) -> None:
        self.values = dict(values)
        self.validate = validate
        self.scored: Dict[str, torch.Tensor] = {}     # name -> scaled scalar contribution
        self.batch_shape: Tuple[int, ...] = ()
        self.muted = False
        self.drawing = False                          # prior-predictive mode


_STACK: List[_Frame] = []


def _frame() -> _Frame:
    if not _STACK:
        raise RuntimeError("oracle.handlers.sample called outside evaluate() / draw()")
    return _STACK[-1]


def _shape(shape: Any) -> torch.Size:
    if shape is None:
        return torch.Size()
    if isinstance(shape, torch.Size):
        return shape
    if isinstance(shape, int) or (torch.is_tensor(shape) and shape.ndim == 0):
        return torch.Size([int(shape)])
    return torch.Size(shape)


def _is_masked(value: Any) -> bool:
    return isinstance(value, torch.masked.MaskedTensor)


class Value(Distribution):
    """Deterministic site (mininf/core.py:390-448)."""
    arg_constraints: Dict[str, Any] = {}

    def __init__(self, value: Any = None, support: Any = None) -> None:
        super().__init__(torch.Size(), torch.Size(), validate_args=False)
        self.value = torch.as_tensor(value) if isinstance(value, (int, float)) else value
        self._support = support or constraints.real

    @property
    def support(self) -> Any:  # type: ignore[override]
        return self._support


def _check(value: Any, name: str, dist: Distribution, sample_shape: torch.Size,
           declared: Tuple[int, ...]) -> None:
    """Shape and support validation (mininf/core.py:142-189)."""
    expected = tuple(sample_shape) + tuple(dist.batch_shape) + tuple(dist.event_shape)
    actual = tuple(value.shape)
    bad = len(expected) != len(actual) or any(
        got != want and dim >= len(declared) for dim, (want, got) in enumerate(zip(expected, actual)))
    if bad:
        raise ValueError(f"Expected shape {expected} for parameter '{name}' but got {actual}.")
    if _is_masked(value):
        mask = value.get_mask()
        for _ in range(dist.support.event_dim):
            mask = mask.all(dim=-1)
        ok = bool(dist.support.check(value.get_data())[mask].all())
    else:
        ok = bool(dist.support.check(value).all())
    if not ok:
        raise ValueError(f"Parameter '{name}' is not in the support of {dist}.")


def sample(name: str, distribution: Distribution, sample_shape: Any = None) -> torch.Tensor:
    frame = _frame()
    shape = _shape(sample_shape)
    if isinstance(distribution, Value):
        return frame.values.get(name, distribution.value)
    if frame.drawing:
        if frame.values.get(name) is None:
            frame.values[name] = distribution.sample(shape)
        return frame.values[name]
    if name in frame.scored:
        raise RuntimeError(f"'{name}' was sampled twice")
    value = frame.values.get(name)
    if value is None:
        raise ValueError(f"variable '{name}' is missing")
    if frame.validate:
        _check(value, name, distribution, shape, frame.batch_shape)
    if frame.muted:
        return value
    # LogProbTracer.sample: masked data is evaluated densely, then only unmasked entries count
    if _is_masked(value):
        if frame.batch_shape:
            raise ValueError("Batch dimensions are not supported for masked data.")
        data, mask = value.get_data(), value.get_mask()
        saved = distribution._validate_args
        distribution._validate_args = False
        try:
            log_prob = distribution.log_prob(data)
        finally:
            distribution._validate_args = saved
        contribution = log_prob[mask].sum()
    else:
        log_prob = distribution.log_prob(value)
        contribution = log_prob.sum()
        if frame.batch_shape:
            declared = torch.Size(frame.batch_shape)
            contribution = contribution * declared.numel() / \
                log_prob.shape[:len(declared)].numel()
    frame.scored[name] = contribution
    return value


def value(name: str, value: Any = None, shape: Any = None, support: Any = None) -> torch.Tensor:
    return sample(name, Value(value, support), shape)


@contextlib.contextmanager
def batch(shape: Any) -> Iterator[None]:
    frame = _frame()
    previous = frame.batch_shape
    frame.batch_shape = tuple(_shape(shape))
    try:
        yield
    finally:
        frame.batch_shape = previous


@contextlib.contextmanager
def no_log_prob() -> Iterator[None]:
    frame = _frame()
    previous = frame.muted
    frame.muted = True
    try:
        yield
    finally:
        frame.muted = previous


def condition(model: Callable, values: Optional[Dict[str, Any]] = None, **kwargs: Any) -> Callable:
    pinned = dict(values or {})
    pinned.update(kwargs)

    def conditioned(*args: Any, **kw: Any) -> Any:
        frame = _frame()
        conflict = set(frame.values) & set(pinned)
        if conflict:
            raise ValueError(f"Cannot update state; it already has parameters {conflict}.")
        frame.values.update(pinned)
        return model(*args, **kw)

    return conditioned


def evaluate(model: Callable, values: Dict[str, Any], validate: bool = True) -> Dict[str, torch.Tensor]:
    """Run ``model`` with every site pinned; returns name -> scaled log-density contribution."""
    frame = _Frame(values, validate)
    _STACK.append(frame)
    try:
        model()
    finally:
        _STACK.pop()
    return frame.scored


def draw(model: Callable, values: Optional[Dict[str, Any]] = None) -> Dict[str, Any]:
    """Prior-predictive draw (SampleTracer, mininf/core.py:192-204)."""
    frame = _Frame(values or {}, validate=False)
    frame.drawing = True
    _STACK.append(frame)
    try:
        model()
    finally:
        _STACK.pop()
    return frame.values
