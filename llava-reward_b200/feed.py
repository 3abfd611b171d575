"""Host -> device input feeding for the batch-eval loops (`batch_eval.score_pairs / score_single / best_of_n`).

The reference moves every batch with blocking `.to('cuda')` calls inside its loop (eval/batch_inference_rm_phi.py:
82-90), so the GPU idles during each copy - 1.47 GB of fp32 pixels per 64 samples at config 2.
`DevicePrefetcher` wraps the same iterable of CPU batches and yields them on the device, with batch k+1 staged through
reusable pinned buffers and copied on a side stream while batch k is being scored. No arithmetic happens here.
"""
from __future__ import annotations

from collections.abc import Mapping
from typing import Any, Dict, Iterable, Iterator, List, Tuple

import torch


def _map_tensors(obj: Any, fn):
    """Apply fn to every tensor of a nested batch structure (dict / BatchFeature / tuple / list); leaves the rest."""
    if torch.is_tensor(obj):
        return fn(obj)
    if isinstance(obj, Mapping):
        mapped = {k: _map_tensors(v, fn) for k, v in obj.items()}
        try:   # keep BatchFeature-like containers (transformers) as they are
            return type(obj)(mapped)
        except Exception:
            return mapped
    if isinstance(obj, tuple):
        return tuple(_map_tensors(v, fn) for v in obj)
    if isinstance(obj, list):
        return [_map_tensors(v, fn) for v in obj]
    return obj


class DevicePrefetcher:
    """Iterate `batches` (CPU tensors in any nesting) as device tensors, one batch ahead.

    depth pinned staging sets are reused round-robin; a set is rewritten only after the copy that read it has finished
    (event per set). Device tensors are allocated on the copy stream and handed to the consumer's stream with
    `wait_stream` + `record_stream`, so they are safe to use and to free in the usual way.
    """

    def __init__(self, batches: Iterable, device="cuda", depth: int = 2):
        self.batches, self.device, self.depth = batches, torch.device(device), max(2, depth)
        if self.device.type != "cuda":
            raise RuntimeError("DevicePrefetcher feeds a CUDA device (there is no CPU path in this package)")
        self.stream = torch.cuda.Stream(device=self.device)
        # one pinned byte arena per staging set, grown to the largest batch seen (ragged batches - Qwen's
        # [sum_patches, 1176] pixels, varying S - reuse it through views instead of pinning a new buffer per shape)
        self._arena: List[torch.Tensor] = [torch.empty(0, dtype=torch.uint8) for _ in range(self.depth)]
        self._copied = [None] * self.depth   # event: the H2D copies out of staging set i have finished
        self.h2d_bytes = 0

    @property
    def pinned_bytes(self) -> int:
        return sum(a.numel() for a in self._arena)

    def _stage(self, k: int, batch):
        s = k % self.depth
        if self._copied[s] is not None:
            self._copied[s].synchronize()
        need = [0]

        def measure(t: torch.Tensor):
            if not t.is_cuda:
                need[0] += (t.numel() * t.element_size() + 255) // 256 * 256
            return t

        _map_tensors(batch, measure)
        if need[0] > self._arena[s].numel():
            self._arena[s] = torch.empty(need[0] + need[0] // 4, dtype=torch.uint8).pin_memory()
        arena, off = self._arena[s], [0]

        def to_dev(t: torch.Tensor):
            if t.is_cuda:
                return t
            nbytes = t.numel() * t.element_size()
            stage = arena[off[0]: off[0] + nbytes].view(t.dtype).view(t.shape)
            off[0] += (nbytes + 255) // 256 * 256
            stage.copy_(t)
            self.h2d_bytes += nbytes
            return stage.to(self.device, non_blocking=True)

        with torch.cuda.stream(self.stream):
            out = _map_tensors(batch, to_dev)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._copied[s] = ev
        return out

    def __iter__(self) -> Iterator:
        it = iter(self.batches)
        k = 0
        try:
            nxt = self._stage(k, next(it))
        except StopIteration:
            return
        while True:
            cur = nxt
            k += 1
            try:
                nxt = self._stage(k, next(it))   # queued behind cur's copies on the side stream
                more = True
            except StopIteration:
                more = False
            consumer = torch.cuda.current_stream(self.device)
            # only the copies of `cur` must have landed; waiting on its event keeps the next batch's copy in flight
            consumer.wait_event(self._copied[(k - 1) % self.depth])
            _map_tensors(cur, lambda t: (t.record_stream(consumer), t)[1] if t.is_cuda else t)
            yield cur
            if not more:
                return
