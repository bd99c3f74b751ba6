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
                  dst: int = 0) -> Optional[List[Any]]:
    """Run `run_shard` on this rank's block of `rows` and gather the per-row outputs, in
    the original order, on rank `dst` (other ranks get None).  Works with any initialised
    process group (NCCL in production, gloo in the CPU tests)."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return list(run_shard(rows))
    world, rank = dist.get_world_size(), dist.get_rank()
    lo, hi = shard_bounds(len(rows), world, rank)
    mine = list(run_shard(rows[lo:hi]))
    if len(mine) != hi - lo:
        raise RuntimeError(f"rank {rank}: shard produced {len(mine)} outputs for {hi - lo} rows")
    parts: Optional[List[Any]] = [None] * world if rank == dst else None
    dist.gather_object(mine, parts, dst=dst)
    if rank != dst:
        return None
    out: List[Any] = []
    for p in parts:
        out.extend(p)
    return out
