"""World-size-2 (and 3) data-parallel scoring logic on CPU with the gloo backend: sharding, ragged shards,
empty shards and the single end-of-job all_gather reproduce the single-process result order."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from llava_reward_b200.dp import gather_rows, score_pairs_dp, shard_indices


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_score(idx):
    i = torch.tensor(list(idx), dtype=torch.float32)
    rc = torch.stack([i * 0.5, -i], dim=1)
    rr = torch.stack([i + 1.0, i * 0.25], dim=1)
    prob = torch.sigmoid((rc[:, 0] * rr[:, 1] - rc[:, 1] * rr[:, 0]) / 10.0)
    return rc, rr, prob


def _worker(rank, world, port, n_pairs, micro, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out = score_pairs_dp(_fake_score, n_pairs, micro, rank, world)
        # plain lists: a tensor in an mp.Queue is shared through an fd served by the producer, which may have exited
        q.put((rank, [t.tolist() for t in out]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_pairs,micro", [(2, 9, 2), (2, 8, 3), (3, 2, 4), (2, 1, 1)])
def test_dp_matches_single_process(world, n_pairs, micro):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_pairs, micro, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = _fake_score(range(n_pairs))
    for rank, out in results:
        for a, b in zip(out, ref):
            assert torch.allclose(torch.tensor(a), b), (rank, a, b)


def test_shard_indices_cover_everything_once():
    for world in (1, 2, 3, 8):
        for n in (0, 1, 7, 64):
            seen = sorted(i for r in range(world) for i in shard_indices(n, r, world))
            assert seen == list(range(n))


def test_gather_rows_single_process_is_identity():
    x = torch.arange(12.0).view(4, 3)
    assert gather_rows(x, 4, 0, 1) is x
