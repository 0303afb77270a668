"""mininf_b200: the mininf interface with a B200-native ELBO engine underneath.

``import mininf_b200 as mininf`` is the intended drop-in: ``sample``, ``condition``, ``batch``,
``no_log_prob``, ``value``, ``State`` and ``nn`` keep the reference's names and semantics
(mininf/__init__.py:1-14); ``nn.EvidenceLowerBoundLoss`` runs on hand-written sm_100a kernels
(see DESIGN.md).
"""
from . import core, distributions, nn, stream, util  # noqa: F401
from .core import State, batch, broadcast_samples, condition, no_log_prob, sample, value  # noqa: F401

__all__ = ["batch", "broadcast_samples", "condition", "nn", "no_log_prob", "value", "sample",
           "State"]
