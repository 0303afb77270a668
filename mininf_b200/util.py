"""Small helpers shared by the host-side mirror of the mininf interface.

Mirrors the behaviour of the reference's ``mininf/util.py`` (shape normalisation :15-25,
masked-aware constraint checks :43-66, dense gradients for masked data :69-92, number-to-tensor
coercion :95-107) with its names kept so user code and tests port unchanged.
"""
from __future__ import annotations

import numbers
import os
from typing import Any, Dict, Union

import torch
from torch.distributions.constraints import Constraint

IN_CI = "CI" in os.environ

OptionalSize = Union[torch.Size, torch.Tensor, int, None]
TensorDict = Dict[str, torch.Tensor]


def is_masked(value: Any) -> bool:
    return isinstance(value, torch.masked.MaskedTensor)


def _normalize_shape(shape: OptionalSize) -> torch.Size:
    """``None`` -> ``()``, an int or 0-dim tensor -> ``(n,)``, anything else through
    ``torch.Size`` (reference table: tests/test_util.py:92-102)."""
    if shape is None:
        return torch.Size()
    if isinstance(shape, torch.Size):
        return shape
    scalar_tensor = torch.is_tensor(shape) and shape.ndim == 0
    if isinstance(shape, int) or scalar_tensor:
        return torch.Size([int(shape)])
    return torch.Size(shape)


def _format_dict_compact(value: Dict[str, Any], id_: int | None = None,
                         name: str | None = None) -> str:
    """One-line description of a mapping: tensors by shape, everything else by type."""
    parts = []
    for key, element in value.items():
        if isinstance(element, torch.Tensor):
            shown = f"{type(element).__name__}(shape={tuple(element.shape)})"
        else:
            shown = str(type(element))
        parts.append(f"'{key}': {shown}")
    label = name or type(value).__name__
    return f"<{label} at {hex(id_ or id(value))} comprising {{{', '.join(parts)}}}>"


def check_constraint(constraint: Constraint, value: torch.Tensor) -> torch.Tensor:
    """Element-wise support check that understands ``MaskedTensor``: the mask is all-reduced over
    the constraint's event dimensions and the raw data is checked without gradients."""
    if not is_masked(value):
        return constraint.check(value)
    mask = value.get_mask()
    for _ in range(constraint.event_dim):
        mask = mask.all(dim=-1)
    with torch.no_grad():
        return torch.masked.as_masked_tensor(constraint.check(value.get_data()), mask)


class _MaskedData(torch.autograd.Function):
    """Identity on the data of a masked tensor whose backward zeroes the masked-out slots."""

    @staticmethod
    def forward(ctx, value):  # type: ignore[override]
        ctx.mask = value._masked_mask
        return value._masked_data

    @staticmethod
    def backward(ctx, grad_output):  # type: ignore[override]
        if torch.masked.is_masked_tensor(grad_output):
            grad_output = grad_output._masked_data
        return torch.where(ctx.mask, grad_output, 0)


def get_masked_data_with_dense_grad(value: torch.masked.MaskedTensor) -> torch.Tensor:
    """Data of a masked tensor with dense (zero-filled) gradients."""
    return _MaskedData.apply(value)


def maybe_as_tensor(value: Any) -> Any:
    """Plain numbers become tensors; everything else is returned untouched."""
    if value is not None and isinstance(value, numbers.Number):
        return torch.as_tensor(value)
    return value
