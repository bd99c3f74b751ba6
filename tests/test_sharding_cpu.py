"""N>1 host logic on CPU: gloo, world_size 2 — shard bounds, positional gather, and the
weight broadcast helper (tensor equality after broadcast)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sutro_b200.sharding import balanced_shards, broadcast_weights, infer_sharded, shard_bounds


def test_shard_bounds_cover_every_row_once():
    for n in (0, 1, 7, 8, 20000, 100003):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_balanced_shards_partition_rows_and_even_out_cost():
    import random
    rng = random.Random(0)
    costs = [int(rng.lognormvariate(4.5, 0.6)) for _ in range(20001)]
    for world in (1, 2, 3, 8):
        shards = balanced_shards(costs, world)
        assert sorted(i for s in shards for i in s) == list(range(len(costs)))
        assert all(s == sorted(s) for s in shards)
        sizes = [len(s) for s in shards]
        loads = [sum(costs[i] for i in s) for s in shards]
        assert max(sizes) - min(sizes) <= 1
        assert max(loads) - min(loads) <= max(costs)          # within one row of each other
    # contiguous blocks of length-sorted data are badly unbalanced; the snake deal is not
    ordered = sorted(costs)
    blocks = [sum(ordered[slice(*shard_bounds(len(ordered), 4, r))]) for r in range(4)]
    dealt = [sum(ordered[i] for i in s) for s in balanced_shards(ordered, 4)]
    assert max(blocks) > 2 * min(blocks) and max(dealt) - min(dealt) <= max(ordered)
    assert balanced_shards([], 2) == [[], []] and balanced_shards([7], 3) == [[0], [], []]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker_balanced(rank, world, port, n_rows, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rows = [("x" * ((i * 7) % 13)) + f"|{i}" for i in range(n_rows)] + [None]
        out = infer_sharded(rows, lambda shard: [f"{r}@{rank}" for r in shard], dst=0,
                            balance="bytes")
        if rank == 0:
            q.put(out)
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


def test_infer_sharded_balanced_scatter_back_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    n_rows = 23
    procs = [ctx.Process(target=_worker_balanced, args=(r, 2, port, n_rows, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    rows = [("x" * ((i * 7) % 13)) + f"|{i}" for i in range(n_rows)] + [None]
    shards = balanced_shards([0 if r is None else len(r) for r in rows], 2)
    owner = {i: r for r, s in enumerate(shards) for i in s}
    assert out == [f"{rows[i]}@{owner[i]}" for i in range(len(rows))]
    assert {owner[i] for i in range(len(rows))} == {0, 1}


def _worker(rank, world, port, n_rows, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w = [torch.full((4, 3), float(rank + 1)), torch.arange(5.0) * (rank + 1)]
        broadcast_weights(w, src=0)
        assert torch.equal(w[0], torch.full((4, 3), 1.0)) and torch.equal(w[1], torch.arange(5.0))
        rows = [f"row-{i}" for i in range(n_rows)]
        seen = []

        def run_shard(shard):
            seen.extend(shard)
            return [f"{r}@{rank}" for r in shard]       # variable-length strings
        out = infer_sharded(rows, run_shard, dst=0)
        lo, hi = shard_bounds(n_rows, world, rank)
        assert seen == rows[lo:hi]
        if rank == 0:
            q.put(out)
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_rows", [11, 2])
def test_infer_sharded_preserves_row_order_world2(n_rows):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_rows, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    lo, hi = shard_bounds(n_rows, 2, 0)
    assert out == [f"row-{i}@{0 if i < hi else 1}" for i in range(n_rows)]
