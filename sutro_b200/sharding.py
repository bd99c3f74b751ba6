"""Row sharding across GPUs (one process per GPU, launched by torchrun).

The path is embarrassingly parallel — rows are independent and results positional
(sutro/sdk.py:406-412) — so there is no data-path collective: every rank runs the full
hot path on a contiguous block of rows with its own model replica, and the outputs are
gathered on the host (variable-length strings).  The only GPU collective is the initial
weight broadcast from rank 0 over NCCL/NVLink.
"""
from __future__ import annotations

from typing import Any, Callable, List, Optional, Sequence, Tuple


def shard_bounds(n_rows: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous balanced blocks: the first n_rows % world ranks get one extra row."""
    base, extra = divmod(n_rows, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def balanced_shards(costs: Sequence[int], world: int) -> List[List[int]]:
    """Cost-balanced assignment of rows to `world` shards (balance tokens
    rather than rows).  Rows are dealt longest-first in snake order (0..w-1, w-1..0, …), which
    keeps both the row counts (±1) and the summed cost of the shards close; every shard's
    indices come back in ascending order so results can be scattered back positionally.
    `costs` is any per-row proxy for work — the UTF-8 length of the row is what callers use."""
    shards: List[List[int]] = [[] for _ in range(world)]
    order = sorted(range(len(costs)), key=lambda i: (-int(costs[i]), i))
    for k, i in enumerate(order):
        lap, pos = divmod(k, world)
        shards[pos if lap % 2 == 0 else world - 1 - pos].append(i)
    for s in shards:
        s.sort()
    return shards


def broadcast_weights(tensors, src: int = 0) -> float:
    """NCCL broadcast of every weight tensor from `src`; returns milliseconds."""
    import time

    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return 0.0
    if tensors and tensors[0].is_cuda:
        torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for t in tensors:
        dist.broadcast(t, src=src)
    if tensors and tensors[0].is_cuda:
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3


def infer_sharded(rows: Sequence[Any], run_shard: Callable[[Sequence[Any]], List[Any]],
                  dst: int = 0, balance: str = "rows") -> Optional[List[Any]]:
    """Run `run_shard` on this rank's share of `rows` and gather the per-row outputs, in
    the original order, on rank `dst` (other ranks get None).  `balance="rows"` gives
    contiguous blocks, `"bytes"` the cost-balanced assignment of `balanced_shards` (every
    rank computes the same assignment from the same rows, so nothing is exchanged up front).
    Works with any initialised process group (NCCL in production, gloo in the CPU tests)."""
    import torch.distributed as dist
    if balance not in ("rows", "bytes"):
        raise ValueError("balance must be 'rows' or 'bytes'")
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return list(run_shard(rows))
    world, rank = dist.get_world_size(), dist.get_rank()
    if balance == "rows":
        shards = [list(range(*shard_bounds(len(rows), world, r))) for r in range(world)]
    else:
        shards = balanced_shards([len(str(r).encode("utf-8")) if r is not None else 0
                                  for r in rows], world)
    idx = shards[rank]
    mine = list(run_shard([rows[i] for i in idx]))
    if len(mine) != len(idx):
        raise RuntimeError(f"rank {rank}: shard produced {len(mine)} outputs for {len(idx)} rows")
    parts: Optional[List[Any]] = [None] * world if rank == dst else None
    dist.gather_object(mine, parts, dst=dst)
    if rank != dst:
        return None
    out: List[Any] = [None] * len(rows)
    for ids, p in zip(shards, parts):
        for i, v in zip(ids, p):
            out[i] = v
    return out
