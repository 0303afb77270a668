"""Out-of-core minibatch feed for data sets that do not fit in HBM.

The reference's minibatch loop (``examples/minibatch.md:68-88``) iterates a
``torch.utils.data.DataLoader`` over host tensors and conditions the model on every batch::

    for X, y in loader:
        conditioned = mininf.condition(model, X=X, y=y)
        ...

``HostBatchStream`` is the feed for the same loop when the tensors live in (pinned) host memory
and the model runs on a B200: batches are copied host -> device on a dedicated copy stream into a
ring of ``depth`` staging buffers, so the copy of batch ``i + 1`` overlaps the ELBO sweep of batch
``i``. The staging buffers keep their layout, so ``EvidenceLowerBoundLoss`` rebinds its cached
plan to each batch (``Plan.rebind``) instead of tracing the model again. The feed is PCIe-bound
(≈ 55 GB/s measured): it hides the sweep behind the copy, not the other way round.
"""
from __future__ import annotations

from typing import Dict, Iterator, List, Optional

import torch


class HostBatchStream:
    """Iterate ``{name: device tensor}`` batches of ``batch_rows`` leading rows over host tensors.

    Args:
        tensors: host tensors with a common leading dimension (pinned on construction unless
            they already are; pass ``pin=False`` to stream from pageable memory).
        batch_rows: leading rows per batch; a ragged last batch is yielded as shorter views.
        device: CUDA device of the staging buffers.
        depth: staging buffers per tensor (2 = double buffering).
        order: optional 1-D index tensor of batch starts' order (e.g. a permutation of
            ``range(n_batches)``) for shuffled epochs; rows inside a batch stay contiguous.

    A yielded batch stays valid until ``depth - 1`` further batches have been requested: the
    stream records an event on the consumer's stream when the next batch is asked for and the
    copy that reuses a buffer waits for the event of its previous consumer.
    """

    def __init__(self, tensors: Dict[str, torch.Tensor], batch_rows: int, device: torch.device | str = "cuda",
                 depth: int = 2, order: Optional[torch.Tensor] = None, pin: bool = True) -> None:
        if not tensors:
            raise ValueError("HostBatchStream needs at least one tensor")
        rows = {int(t.shape[0]) for t in tensors.values()}
        if len(rows) != 1:
            raise ValueError(f"all tensors must share the leading dimension; got {sorted(rows)}")
        if batch_rows < 1 or depth < 2:
            raise ValueError("batch_rows must be positive and depth at least 2")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("HostBatchStream stages batches on a CUDA device (there is no CPU fallback)")
        self.n_rows = rows.pop()
        self.batch_rows = int(batch_rows)
        self.depth = int(depth)
        self.n_batches = -(-self.n_rows // self.batch_rows)
        self.order = None if order is None else [int(i) for i in order]
        if self.order is not None and sorted(self.order) != list(range(self.n_batches)):
            raise ValueError("order must be a permutation of range(n_batches)")
        self.host = {name: (t if (t.is_pinned() or not pin) else t.contiguous().pin_memory())
                     for name, t in tensors.items()}
        for name, t in self.host.items():
            if t.device.type != "cpu":
                raise ValueError(f"'{name}' must be a host tensor")
        self.staging: List[Dict[str, torch.Tensor]] = [
            {name: torch.empty((self.batch_rows,) + tuple(t.shape[1:]), dtype=t.dtype, device=self.device)
             for name, t in self.host.items()} for _ in range(self.depth)]
        self.copy_stream = torch.cuda.Stream(self.device)
        self.bytes_per_batch = sum(self.batch_rows * (t[0].numel() if t.dim() > 1 else 1) * t.element_size()
                                   for t in self.host.values())

    def __len__(self) -> int:
        return self.n_batches

    def _rows(self, position: int) -> tuple:
        index = position if self.order is None else self.order[position]
        lo = index * self.batch_rows
        return lo, min(lo + self.batch_rows, self.n_rows)

    def _enqueue(self, position: int, slot: int, free: Optional[torch.cuda.Event]) -> torch.cuda.Event:
        lo, hi = self._rows(position)
        with torch.cuda.stream(self.copy_stream):
            if free is not None:
                self.copy_stream.wait_event(free)        # the slot's previous consumer has finished
            for name, source in self.host.items():
                self.staging[slot][name][:hi - lo].copy_(source[lo:hi], non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self.copy_stream)
        return ready

    def __iter__(self) -> Iterator[Dict[str, torch.Tensor]]:
        ready: List[Optional[torch.cuda.Event]] = [None] * self.depth
        free: List[Optional[torch.cuda.Event]] = [None] * self.depth
        # A new epoch (or a new iterator after an early `break`) refills slots whose last
        # consumers - the asynchronous sweeps of the previous epoch - may still be running on the
        # caller's stream: nothing of this epoch may be copied before they have finished.
        self.copy_stream.wait_stream(torch.cuda.current_stream(self.device))
        for position in range(min(self.depth - 1, self.n_batches)):
            ready[position % self.depth] = self._enqueue(position, position % self.depth, None)
        for position in range(self.n_batches):
            slot = position % self.depth
            ahead = position + self.depth - 1
            if ahead < self.n_batches:
                # reuses the slot of batch `position - 1`, whose consumer work is on the current stream
                ready[ahead % self.depth] = self._enqueue(ahead, ahead % self.depth, free[ahead % self.depth])
            torch.cuda.current_stream(self.device).wait_event(ready[slot])
            lo, hi = self._rows(position)
            yield {name: buffer[:hi - lo] for name, buffer in self.staging[slot].items()}
            done = torch.cuda.Event()
            done.record(torch.cuda.current_stream(self.device))
            free[slot] = done
