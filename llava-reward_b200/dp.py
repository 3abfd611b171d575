"""Data-parallel pair scoring: one process per GPU, a full bf16 replica each, pairs sharded round-robin
(pair i -> rank i mod world, so chosen+rejected of a pair stay on one GPU and `preference_compute` is local),
ONE collective at the end: all_gather of [n_local, 2*vhd + 1] fp32 rows (rewards_c | rewards_r | prob).
The reference has no multi-GPU scoring (DistributedSampler(num_replicas=1), eval/batch_inference_rm_phi.py:50-57);
this is the SURVEY.md 8(e) design. There is no data-path collective: NCCL carries ~12-20 B per pair.
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    """Round-robin ownership: item i belongs to rank i % world."""
    return list(range(rank, n_items, world))


def padded_local_count(n_items: int, world: int) -> int:
    return (n_items + world - 1) // world


def gather_rows(local: torch.Tensor, n_items: int, rank: int, world: int) -> torch.Tensor:
    """local [n_local, C] (rows of this rank's shard, in shard order) -> [n_items, C] in original order on every rank."""
    if world == 1:
        return local
    n_pad = padded_local_count(n_items, world)
    buf = torch.zeros(n_pad, local.shape[1], dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    out = torch.empty(world * n_pad, local.shape[1], dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, buf)
    # out[r * n_pad + k] is item r + k * world
    out = out.view(world, n_pad, -1).transpose(0, 1).reshape(world * n_pad, -1)
    return out[:n_items]


def score_pairs_dp(score_batch: Callable[[Sequence[int]], Tuple[torch.Tensor, torch.Tensor, torch.Tensor]],
                   n_pairs: int, micro_batch: int, rank: int, world: int):
    """Score `n_pairs` pairs data-parallel.

    score_batch(pair_indices) -> (rewards_chosen [b, vhd], rewards_rejected [b, vhd], prob [b]) for the given
    GLOBAL pair indices (the caller builds the inputs and runs `custom_forward` twice + `preference_compute`).
    Returns (rewards_chosen [n, vhd], rewards_rejected [n, vhd], prob [n]) in the original pair order on every rank.
    """
    mine = shard_indices(n_pairs, rank, world)
    rows = []
    for i in range(0, len(mine), micro_batch):
        idx = mine[i:i + micro_batch]
        rc, rr, p = score_batch(idx)
        rows.append(torch.cat([rc.float(), rr.float(), p.float().reshape(-1, 1)], dim=1))
    vhd = rows[0].shape[1] // 2 if rows else 1
    if rows:
        local = torch.cat(rows, dim=0)
    else:  # a rank may own nothing when n_pairs < world; it still joins the collective
        local = torch.zeros(0, 2 * vhd + 1)
    if world > 1:
        # ranks with an empty shard need the column count and device of the others
        meta = torch.tensor([local.shape[1], 1 if rows else 0], device=local.device if rows else None)
        if not rows:
            dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
            meta = meta.to(dev)
        dist.all_reduce(meta, op=dist.ReduceOp.MAX)
        if not rows:
            local = torch.zeros(0, int(meta[0]), device=meta.device)
        vhd = (int(meta[0]) - 1) // 2
    full = gather_rows(local, n_pairs, rank, world)
    return full[:, :vhd], full[:, vhd:2 * vhd], full[:, 2 * vhd]
