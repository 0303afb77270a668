"""Trace a model ONCE into a flat site table.

The reference re-executes the Python model on every step and lets autograd differentiate whatever
torch code links latent draws to distribution parameters (mininf/nn.py:223-225). Here the model
is replayed a single time under :class:`SiteTableTracer` with the latent draws wrapped in
:class:`LinkTensor` - a ``torch.Tensor`` subclass that carries real values (so every shape check
and argument validation of ``torch.distributions`` runs as usual) plus a symbolic *link
expression* describing how the value was built from latents and data. The closed set of link
forms is what the CUDA kernels implement (include/mininf_b200.h):

* ``Affine``:  ``T(a_const + a_lat + (b_const + b_lat) * x)`` - constants, data tensors, a latent
  itself, ``c + d*x``, ``exp(a + b*x)``, and what reduces to it: ``-z``, ``c*z``, ``z/c``, ``z/x``,
  ``a - b*x``, ``c - x``, ``(c + z)*x`` (a scaled latent moves to the slope, the constant or the
  product of data tensors becomes its covariate);
* ``Dense``:   ``T(icpt + X @ theta)`` - ``X @ theta`` with an optional scalar intercept;
* ``RowDot``:  ``T(icpt + Z @ beta)`` - a per-observation latent matrix times a latent vector
  (``z * slope`` for one latent feature per row);
* ``Linear``:  ``T(icpt + b1*x1 + b2*x2 + ...)`` - one scalar latent per covariate; lowered like
  ``Dense`` over the design matrix ``[x1 x2 ...]`` built once at trace time.

Anything else makes the tensor *opaque*; an opaque tensor reaching a distribution parameter of a
site that contributes to the log-density raises ``NotImplementedError`` (no fallback).
"""
from __future__ import annotations

import dataclasses
import numbers
from typing import Any, Dict, List, Optional, Tuple

import torch
from torch.distributions import Distribution

from ..core import State, TracerMixin, Value, batch, batch_scale, no_log_prob
from ..util import OptionalSize, is_masked


@dataclasses.dataclass(frozen=True)
class LatentRef:
    """A reference to latent site ``name``: the whole tensor (``index is None``, aligned
    element-wise with the expression) or one flat element of it."""
    name: str
    index: Optional[int] = None

    @property
    def is_scalar(self) -> bool:
        return self.index is not None


@dataclasses.dataclass
class Affine:
    a_const: float = 0.0
    a_lat: Optional[LatentRef] = None
    b_const: float = 0.0
    b_lat: Optional[LatentRef] = None
    x: Optional[torch.Tensor] = None      # data covariate, shaped like the expression
    transform: str = "id"

    @property
    def has_x_term(self) -> bool:
        return self.b_const != 0.0 or self.b_lat is not None

    @property
    def is_pure_latent(self) -> bool:
        return self.a_lat is not None and self.a_const == 0.0 and not self.has_x_term and \
            self.transform == "id"

    @property
    def only_scalars(self) -> bool:
        return all(ref is None or ref.is_scalar for ref in (self.a_lat, self.b_lat))


@dataclasses.dataclass
class Dense:
    X: torch.Tensor                       # [n, p] data
    theta: str                            # latent site holding the p coefficients
    icpt_const: float = 0.0
    icpt_lat: Optional[LatentRef] = None
    transform: str = "id"


@dataclasses.dataclass
class RowDot:
    """``T(icpt + Z @ beta)`` with Z a per-observation latent matrix [n, p] and beta a latent
    vector [p] (examples/regression-with-feature-uncertainty.md:38 widened to p features)."""
    Z: str
    beta: str
    icpt_const: float = 0.0
    icpt_lat: Optional[LatentRef] = None
    transform: str = "id"


@dataclasses.dataclass
class Linear:
    """``T(icpt + sum_k z[ref_k] * x_k)``: several scalar latents, each with a covariate of its own
    (``a + b1 * x1 + b2 * x2``). Lowered as a dense-link site over the design matrix ``[x_1 ... x_K]``
    built once at trace time; the slopes must be adjacent columns of the packed latents."""
    terms: List[Tuple[LatentRef, torch.Tensor]]     # (scalar reference, x[n]) or (whole vector latent, X[n, G])
    icpt_const: float = 0.0
    icpt_lat: Optional[LatentRef] = None
    transform: str = "id"


Expr = Any  # Affine | Dense | RowDot | Linear | None (opaque)


def _as_const(value: Any) -> Optional[float]:
    """A Python number or a one-element plain tensor as a float constant."""
    if isinstance(value, numbers.Number):
        return float(value)
    if isinstance(value, torch.Tensor) and not isinstance(value, LinkTensor) and value.numel() == 1:
        return float(value)
    return None


def _plain(value: Any) -> bool:
    return isinstance(value, torch.Tensor) and not isinstance(value, LinkTensor)


class LinkTensor(torch.Tensor):
    """Real values + the symbolic link expression that produced them (``None`` = opaque)."""
    _expr: Expr

    @staticmethod
    def wrap(data: torch.Tensor, expr: Expr) -> "LinkTensor":
        out = torch.Tensor._make_subclass(LinkTensor, data.detach(), False)
        out._expr = expr
        return out

    def unwrap(self) -> torch.Tensor:
        with torch._C.DisableTorchFunctionSubclass():
            return self.as_subclass(torch.Tensor)

    def __repr__(self, *args: Any, **kwargs: Any) -> str:  # pragma: no cover - debugging aid
        return f"LinkTensor(shape={tuple(self.shape)}, expr={getattr(self, '_expr', None)})"

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):  # type: ignore[override]
        kwargs = kwargs or {}
        name = getattr(func, "__name__", "")
        if name in _HOST_READS and any(isinstance(a, LinkTensor) and a.is_floating_point()
                                       for a in _flatten_args(args)):
            # float(z), z.item(), z.tolist(): the trace-time draw would be frozen into the plan as a
            # constant and the site's dependence on the latent silently lost
            raise NotImplementedError(
                f"`{name}` reads the value of a latent-dependent tensor on the host; the model is "
                "traced once, so the value would be baked into the plan as a constant")
        with torch._C.DisableTorchFunctionSubclass():
            raw = func(*args, **kwargs)
        if _is_inplace(name) or kwargs.get("out") is not None:
            # values were modified behind an existing expression: every tensor involved is opaque
            for target in _flatten_args(args)[:1] + [kwargs.get("out")]:
                if isinstance(target, LinkTensor):
                    target._expr = None
            return _wrap_opaque(raw)
        rule = _RULES.get(name)
        if isinstance(raw, (tuple, list)):
            if name == "broadcast_tensors":
                return type(raw)(_rewrap(r, _shape_rule(a, r)) if isinstance(a, LinkTensor)
                                 else _strip(r) for a, r in zip(_flatten_args(args), raw))
            if name == "unbind":
                return type(raw)(_rule_unbind(args, kwargs, raw))
            # split / chunk / max(dim) / ...: pieces of a latent-dependent tensor stay marked (opaque),
            # so a site that uses one raises instead of freezing the trace-time draw
            return _retuple(raw, [_wrap_opaque(r) for r in raw])
        if not isinstance(raw, torch.Tensor) or not raw.is_floating_point():
            return _strip(raw)
        expr = rule(args, kwargs, raw) if rule is not None else None
        return _rewrap(raw, expr)


_HOST_READS = {"item", "__float__", "__int__", "__index__", "__complex__", "tolist", "numpy", "__array__"}


def _is_inplace(name: str) -> bool:
    return (name.endswith("_") and not name.endswith("__")) or (name.startswith("__i") and name.endswith("__")
                                                                   and name not in ("__iter__", "__index__",
                                                                                    "__int__", "__invert__"))


def _wrap_opaque(value: Any) -> Any:
    """Floating tensors derived from a LinkTensor without a rule: opaque, but still marked."""
    if isinstance(value, torch.Tensor) and value.is_floating_point():
        return _rewrap(value, None)
    return _strip(value)


def _retuple(raw: Any, items: List[Any]) -> Any:
    try:
        return type(raw)(items)
    except TypeError:          # structseq / namedtuple variants that want positional fields
        return tuple(items)


def _strip(value: Any) -> Any:
    if isinstance(value, LinkTensor):
        return value.unwrap()
    return value


def _rewrap(raw: torch.Tensor, expr: Expr) -> LinkTensor:
    raw = _strip(raw)
    return LinkTensor.wrap(raw, expr)


def _flatten_args(args: Tuple[Any, ...]) -> List[Any]:
    if len(args) == 1 and isinstance(args[0], (tuple, list)):
        return list(args[0])
    return list(args)


def _expr_of(value: Any) -> Expr:
    return getattr(value, "_expr", None) if isinstance(value, LinkTensor) else None


# ---------------------------------------------------------------------------------------------
# rules: (args, kwargs, raw result) -> expression of the result or None
# ---------------------------------------------------------------------------------------------
def _shape_rule(source: LinkTensor, raw: torch.Tensor, reshaper=None) -> Expr:
    """Shape-only operations keep the expression if element alignment survives."""
    expr = _expr_of(source)
    if expr is None:
        return None
    same_numel = raw.numel() == source.numel()
    if isinstance(expr, (Dense, RowDot)):
        return expr if same_numel and raw.ndim == 1 else None
    if isinstance(expr, Linear):
        return expr if raw.shape == source.shape else None
    whole = [ref for ref in (expr.a_lat, expr.b_lat) if ref is not None and not ref.is_scalar]
    if whole and not same_numel:
        return None                     # an element-wise latent cannot be broadcast periodically
    x = expr.x
    if x is not None:
        with torch._C.DisableTorchFunctionSubclass():
            x = x.expand(raw.shape) if reshaper is None else reshaper(x)
        if x.shape != raw.shape:
            return None
    return dataclasses.replace(expr, x=x)


def _rule_identity(args, kwargs, raw):
    return _shape_rule(args[0], raw) if isinstance(args[0], LinkTensor) else None


def _rule_reshape(args, kwargs, raw):
    source = args[0]
    if not isinstance(source, LinkTensor):
        return None
    return _shape_rule(source, raw, reshaper=lambda x: x.expand(source.shape).reshape(raw.shape))


def _rule_to(args, kwargs, raw):
    if raw.dtype != args[0].dtype:
        return None
    return _rule_identity(args, kwargs, raw)


def _slope_terms(expr: Expr, shape: torch.Size) -> Optional[Tuple[List[Tuple[LatentRef, torch.Tensor]], float, Optional[LatentRef]]]:
    """(terms, constant, intercept latent) of an expression that is a sum of scalar-latent slopes."""
    if isinstance(expr, Linear):
        return (list(expr.terms), expr.icpt_const, expr.icpt_lat) if expr.transform == "id" else None
    if isinstance(expr, Dense):
        # `X @ theta` / `alpha[group]` joining other covariates: its block is copied into the design matrix
        if expr.transform != "id" or expr.X.numel() > _MAX_INDICATOR_ELEMENTS:
            return None
        return [(LatentRef(expr.theta), expr.X)], expr.icpt_const, expr.icpt_lat
    if not isinstance(expr, Affine) or expr.transform != "id":
        return None
    if expr.a_lat is not None and not expr.a_lat.is_scalar:
        return None
    if not expr.has_x_term:
        return [], expr.a_const, expr.a_lat
    if expr.b_const != 0.0 or expr.b_lat is None or not expr.b_lat.is_scalar or expr.x is None:
        return None                       # a data-only term has no latent column to ride on
    with torch._C.DisableTorchFunctionSubclass():
        return [(expr.b_lat, expr.x.expand(shape))], expr.a_const, expr.a_lat


def _combine_linear(left: Expr, right: Expr, shape: torch.Size) -> Expr:
    """Sum of two expressions that both carry covariates: a several-covariate link."""
    a, b = _slope_terms(left, shape), _slope_terms(right, shape)
    if a is None or b is None or len(shape) != 1:
        return None
    if a[2] is not None and b[2] is not None:
        return None                       # two latent intercepts
    return Linear(terms=a[0] + b[0], icpt_const=a[1] + b[1], icpt_lat=a[2] or b[2])


def _combine_add(left: Expr, right: Expr, shape: torch.Size) -> Expr:
    """left + right for two expressions."""
    if left is None or right is None:
        return None
    if isinstance(left, Linear) or isinstance(right, Linear):
        return _combine_linear(left, right, shape)
    if isinstance(right, (Dense, RowDot)):
        left, right = right, left
    if isinstance(left, Dense) and (isinstance(right, Dense) or (isinstance(right, Affine) and right.has_x_term)):
        return _combine_linear(left, right, shape)       # mixed model: a block of coefficients plus named slopes
    if isinstance(left, (Dense, RowDot)):
        if not isinstance(right, Affine) or right.has_x_term or right.transform != "id" or \
                left.transform != "id":
            return None
        if right.a_lat is not None and (left.icpt_lat is not None or not right.a_lat.is_scalar):
            return None
        return dataclasses.replace(left, icpt_const=left.icpt_const + right.a_const,
                                   icpt_lat=left.icpt_lat or right.a_lat)
    if left.transform != "id" or right.transform != "id":
        return None
    if left.has_x_term and right.has_x_term:
        return _combine_linear(left, right, shape)
    if left.a_lat is not None and right.a_lat is not None:
        return None
    carrier = left if left.has_x_term else right
    return Affine(a_const=left.a_const + right.a_const, a_lat=left.a_lat or right.a_lat,
                  b_const=carrier.b_const, b_lat=carrier.b_lat, x=carrier.x)


def _lift(value: Any, like: torch.Tensor) -> Expr:
    """Constants and plain data tensors as expressions."""
    const = _as_const(value)
    if const is not None:
        return Affine(a_const=const)
    if _plain(value):
        with torch._C.DisableTorchFunctionSubclass():
            x = value.to(like.dtype).expand(like.shape) if value.shape != like.shape else value
        return Affine(b_const=1.0, x=x)
    return _expr_of(value)


def _rule_add(args, kwargs, raw):
    if kwargs.get("alpha", 1) != 1:
        return None
    return _combine_add(_lift(args[0], raw), _lift(args[1], raw), raw.shape)


def _scale_const(expr: Expr, c: float, raw: torch.Tensor) -> Expr:
    """c * expr for a Python constant. The link form has no coefficient on ``a_lat``, so a scaled
    latent moves to the slope with the constant as its covariate: ``c * z = z * full(c)``."""
    if isinstance(expr, Linear) and expr.transform == "id" and (expr.icpt_lat is None or c == 1.0):
        with torch._C.DisableTorchFunctionSubclass():
            return Linear(terms=[(ref, x * c) for ref, x in expr.terms], icpt_const=expr.icpt_const * c + 0.0,
                          icpt_lat=expr.icpt_lat)
    if not isinstance(expr, Affine) or expr.transform != "id":
        return None
    if c == 1.0:
        return expr
    with torch._C.DisableTorchFunctionSubclass():
        if expr.a_lat is not None:
            if expr.has_x_term:
                return None
            x = torch.full(raw.shape, c, dtype=raw.dtype, device=raw.device)
            return Affine(a_const=expr.a_const * c + 0.0, b_lat=expr.a_lat, x=x)
        if expr.b_lat is not None:
            return Affine(a_const=expr.a_const * c + 0.0, b_const=expr.b_const, b_lat=expr.b_lat,
                          x=expr.x.expand(raw.shape) * c)
    return Affine(a_const=expr.a_const * c + 0.0, b_const=expr.b_const * c + 0.0, x=expr.x)


def _scale_data(expr: Expr, weight: torch.Tensor, raw: torch.Tensor) -> Expr:
    """expr * w for a plain data tensor w (element-wise, broadcast to the result)."""
    if not isinstance(expr, Affine) or expr.transform != "id":
        return None
    with torch._C.DisableTorchFunctionSubclass():
        w = weight.to(raw.dtype).expand(raw.shape)
        if not expr.has_x_term:
            # (a_const + z) * w: intercept and latent become the slope of the covariate w
            return Affine(b_const=expr.a_const, b_lat=expr.a_lat, x=w)
        if expr.a_const == 0.0 and expr.a_lat is None:
            return Affine(b_const=expr.b_const, b_lat=expr.b_lat, x=expr.x.expand(raw.shape) * w)
    return None


def _rule_sub(args, kwargs, raw):
    if kwargs.get("alpha", 1) != 1:
        return None
    return _combine_add(_lift(args[0], raw), _scale_const(_lift(args[1], raw), -1.0, raw), raw.shape)


def _rule_rsub(args, kwargs, raw):
    """``other - self`` (``Tensor.__rsub__`` / ``torch.rsub``)."""
    if kwargs.get("alpha", 1) != 1:
        return None
    return _combine_add(_lift(args[1], raw), _scale_const(_lift(args[0], raw), -1.0, raw), raw.shape)


def _rule_neg(args, kwargs, raw):
    return _scale_const(_expr_of(args[0]), -1.0, raw)


def _rule_mul(args, kwargs, raw):
    left, right = args[0], args[1]
    if isinstance(right, LinkTensor) and not isinstance(left, LinkTensor):
        left, right = right, left
    expr = _expr_of(left)
    if isinstance(right, LinkTensor):
        return _rule_latent_product(left, right, raw)
    if expr is None:
        return None
    if isinstance(right, numbers.Number):
        if float(right) == 1.0:
            # `1 * x` is how ParameterizedDistribution hides raw parameters (mininf/nn.py:92-94)
            return _shape_rule(left, raw)
        return _scale_const(expr, float(right), raw)
    if _plain(right) and isinstance(expr, Affine):
        whole = [ref for ref in (expr.a_lat, expr.b_lat) if ref is not None and not ref.is_scalar]
        if whole and left.numel() != raw.numel():
            return None
        return _scale_data(expr, right, raw)
    return None


def _rule_latent_product(left: "LinkTensor", right: "LinkTensor", raw: torch.Tensor) -> Expr:
    """``z * slope``: an element-wise latent vector times a one-element latent is ``Z @ beta`` with
    one feature per row (examples/regression-with-feature-uncertainty.md:38). Which latent the
    scalar reference belongs to is checked when the site is lowered (``beta`` must hold exactly p = 1
    element); every other product of two latent-dependent tensors is outside the link forms."""
    for vector, scalar in ((left, right), (right, left)):
        ve, se = _expr_of(vector), _expr_of(scalar)
        if isinstance(ve, Affine) and ve.is_pure_latent and not ve.a_lat.is_scalar and vector.ndim == 1 and \
                isinstance(se, Affine) and se.is_pure_latent and se.a_lat.is_scalar and se.a_lat.index == 0 and \
                scalar.numel() == 1 and raw.shape == vector.shape:
            return RowDot(Z=ve.a_lat.name, beta=se.a_lat.name)
    return None


def _rule_div(args, kwargs, raw):
    """``expr / c`` and ``expr / data``; a latent in the denominator is outside the link forms."""
    left, right = args[0], args[1]
    if kwargs.get("rounding_mode") is not None or not isinstance(left, LinkTensor) or \
            isinstance(right, LinkTensor):
        return None
    expr = _expr_of(left)
    if isinstance(right, numbers.Number):
        return _scale_const(expr, 1.0 / float(right), raw) if float(right) != 0.0 else None
    if _plain(right) and isinstance(expr, Affine):
        whole = [ref for ref in (expr.a_lat, expr.b_lat) if ref is not None and not ref.is_scalar]
        if whole and left.numel() != raw.numel():
            return None
        with torch._C.DisableTorchFunctionSubclass():
            return _scale_data(expr, right.to(raw.dtype).reciprocal(), raw)
    return None


def _rule_exp(args, kwargs, raw):
    expr = _expr_of(args[0])
    if expr is None or expr.transform != "id":
        return None
    return dataclasses.replace(expr, transform="exp")


def _rule_sigmoid(args, kwargs, raw):
    """Only meaningful as ``Bernoulli(probs=sigmoid(eta))``, which the plan lowers as
    ``Bernoulli(logits=eta)``; anywhere else the lowering raises."""
    expr = _expr_of(args[0])
    if expr is None or expr.transform != "id":
        return None
    return dataclasses.replace(expr, transform="sigmoid")


def _rule_matmul(args, kwargs, raw):
    X, theta = args[0], args[1]
    expr = _expr_of(theta)
    if not isinstance(expr, Affine) or not expr.is_pure_latent or theta.ndim != 1:
        return None
    if expr.a_lat.is_scalar and not (theta.numel() == 1 and expr.a_lat.index == 0):
        return None                       # one-element latents are referenced as scalars
    if _plain(X) and X.ndim == 2:
        return Dense(X=X, theta=expr.a_lat.name)
    left = _expr_of(X)
    if isinstance(X, LinkTensor) and X.ndim == 2 and isinstance(left, Affine) and left.is_pure_latent \
            and not left.a_lat.is_scalar:
        return RowDot(Z=left.a_lat.name, beta=expr.a_lat.name)
    return None


_MAX_INDICATOR_ELEMENTS = 1 << 28     # rows x groups of an `alpha[group]` indicator matrix (1 GiB of fp32)


def _rule_getitem(args, kwargs, raw):
    source, index = args[0], args[1]
    expr = _expr_of(source)
    if not isinstance(expr, Affine) or not expr.is_pure_latent or expr.a_lat.is_scalar:
        return None
    if isinstance(index, torch.Tensor) and not isinstance(index, LinkTensor) and index.ndim == 1 and \
            index.numel() > 1 and index.dtype in (torch.int64, torch.int32) and source.ndim == 1 and raw.ndim == 1:
        # `alpha[group]` - a coefficient per group picked by an integer data vector (random effects) - is
        # one_hot(group) @ alpha: a dense link over an indicator design matrix built once at trace time
        groups = source.numel()
        if index.numel() * groups > _MAX_INDICATOR_ELEMENTS:
            return None
        with torch._C.DisableTorchFunctionSubclass():
            wrapped = torch.where(index < 0, index + groups, index).long()
            indicator = torch.nn.functional.one_hot(wrapped, groups).to(source.dtype)
        return Dense(X=indicator, theta=expr.a_lat.name)
    if raw.numel() != 1:
        return None
    if isinstance(index, torch.Tensor):
        if index.numel() != 1:
            return None
        index = int(index)
    if isinstance(index, int):
        index = (index,)
    if not (isinstance(index, tuple) and len(index) == source.ndim and
            all(isinstance(i, int) for i in index)):
        return None
    flat = 0
    for i, size in zip(index, source.shape):
        flat = flat * size + (i % size)
    return Affine(a_lat=LatentRef(expr.a_lat.name, flat))


def _rule_unbind(args, kwargs, raw):
    """Pieces of `unbind` (and of `Tensor.__iter__`, which calls it): the elements of a 1-D
    element-wise latent are scalar references; anything else is opaque."""
    source = args[0]
    expr = _expr_of(source)
    scalar_pieces = isinstance(expr, Affine) and expr.is_pure_latent and not expr.a_lat.is_scalar and \
        source.ndim == 1
    out = []
    for i, piece in enumerate(raw):
        if not (isinstance(piece, torch.Tensor) and piece.is_floating_point()):
            out.append(_strip(piece))
        elif scalar_pieces:
            out.append(_rewrap(piece, Affine(a_lat=LatentRef(expr.a_lat.name, i))))
        else:
            out.append(_rewrap(piece, None))
    return out


def _rule_select(args, kwargs, raw):
    source = args[0]
    if source.ndim != 1:
        return None
    index = args[2] if len(args) > 2 else kwargs.get("index")
    return _rule_getitem((source, int(index)), {}, raw)


_RULES: Dict[str, Any] = {
    "add": _rule_add, "__add__": _rule_add, "__radd__": _rule_add,
    "sub": _rule_sub, "__sub__": _rule_sub, "subtract": _rule_sub,
    "rsub": _rule_rsub, "__rsub__": _rule_rsub,
    "neg": _rule_neg, "__neg__": _rule_neg, "negative": _rule_neg,
    "mul": _rule_mul, "__mul__": _rule_mul, "__rmul__": _rule_mul, "multiply": _rule_mul,
    "div": _rule_div, "__truediv__": _rule_div, "true_divide": _rule_div, "divide": _rule_div,
    "exp": _rule_exp, "sigmoid": _rule_sigmoid,
    "matmul": _rule_matmul, "__matmul__": _rule_matmul, "mv": _rule_matmul,
    "__getitem__": _rule_getitem, "select": _rule_select,
    "expand": _rule_identity, "expand_as": _rule_identity, "broadcast_to": _rule_identity,
    "contiguous": _rule_identity, "clone": _rule_identity, "detach": _rule_identity,
    "reshape": _rule_reshape, "view": _rule_reshape, "squeeze": _rule_reshape,
    "unsqueeze": _rule_reshape, "flatten": _rule_reshape,
    "to": _rule_to, "float": _rule_to, "type_as": _rule_to,
}


# ---------------------------------------------------------------------------------------------
# the tracer
# ---------------------------------------------------------------------------------------------
@dataclasses.dataclass
class SiteRecord:
    name: str
    distribution: Distribution
    value: torch.Tensor          # LinkTensor (latent), plain tensor or MaskedTensor (observed)
    scale: float                 # declared-over-actual batch ratio (mininf/core.py:267-271)


class SiteTableTracer(TracerMixin):
    """Record every log-density site of one model execution (the plan-time counterpart of
    ``LogProbTracer.sample``, mininf/core.py:211-245): same ``Value`` handling, duplicate /
    missing-value errors, validation and ``no_log_prob`` skip - but nothing is evaluated."""

    def __init__(self, *args: Any, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        self.sites: List[SiteRecord] = []
        self._names: set = set()

    def sample(self, state: State, name: str, distribution: Distribution,
               sample_shape: OptionalSize = None) -> torch.Tensor:
        if isinstance(distribution, Value):
            current = state.get(name, distribution.value)
            self._assert_valid_parameter(current, name, distribution, sample_shape)
            return current
        if name in self._names:
            raise RuntimeError(f"Log probability has already been evaluated for '{name}'. Did you "
                               "call `sample` twice with the same variable name?")
        current = state.get(name)
        if current is None:
            raise ValueError(f"Cannot evaluate log probability; variable '{name}' is missing. Did "
                             "you forget to condition on observed data?")
        self._assert_valid_parameter(current, name, distribution, sample_shape)
        if no_log_prob.get_instance():
            return current
        declared = batch.get_shape()
        if is_masked(current) and declared:
            raise ValueError("Batch dimensions are not supported for masked data.")
        self._names.add(name)
        self.sites.append(SiteRecord(name, distribution, current,
                                     batch_scale(current.shape, declared)))
        return current
