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


# --------------------------------------------------------------------------- sharded frame path
class _CpuEngine:
    """Stand-in for LocalEngine on CPU tensors: the orchestration in infer_frame_sharded
    (broadcast, assignment, row selection, gather, ordered merge) is what is under test; the
    'model' upper-cases a row and appends the rank that processed it."""

    class _Spec:
        embedding_model = False
        d_model = 4

    def __init__(self, rank, embedding=False):
        import torch as _t
        self.device = _t.device("cpu")
        self.rank = rank
        self.spec = self._Spec()
        self.spec.embedding_model = embedding
        self.seen_row_ids = None

    def rows_select(self, d_bytes, d_off, part_rows, part_bytes, d_idx, capacity):
        import numpy as np
        import torch as _t
        b, off, idx = d_bytes.numpy(), d_off.numpy(), d_idx.numpy()
        out, ooff = [], [0]
        for j in idx:
            part, local = divmod(int(j), part_rows)
            o = off[part * (part_rows + 1) + local: part * (part_rows + 1) + local + 2]
            seg = b[part * part_bytes + o[0]: part * part_bytes + o[1]]
            out.append(seg)
            ooff.append(ooff[-1] + len(seg))
        flat = np.concatenate(out) if out else np.zeros(0, np.uint8)
        buf = np.zeros(max(capacity, 1), np.uint8)
        buf[:len(flat)] = flat
        return _t.tensor(ooff, dtype=_t.int64), _t.from_numpy(buf)

    def run_blob_dev(self, d_text, d_off, n_rows, n_bytes, row_ids=None, suffix="", **kw):
        import numpy as np
        import torch as _t
        self.seen_row_ids = None if row_ids is None else list(map(int, row_ids))
        raw, off = d_text.numpy().tobytes(), d_off.numpy()
        rows = [raw[off[i]:off[i + 1]].decode() for i in range(n_rows)]
        if self.spec.embedding_model:
            emb = np.array([[len(r), self.rank, i, 1.0] for i, r in enumerate(rows)], np.float32)
            return dict(d_emb=_t.from_numpy(emb), stats={"rows_done": n_rows, "input_tokens": 1})
        outs = [(r.upper() + f"@{self.rank}{suffix}").encode() for r in rows]
        boff = np.zeros(n_rows + 1, np.int64)
        boff[1:] = np.cumsum([len(o) for o in outs])
        b = np.frombuffer(b"".join(outs), dtype=np.uint8).copy() if outs else np.zeros(1, np.uint8)
        return dict(d_bytes=_t.from_numpy(b), d_boff=_t.from_numpy(boff),
                    stats={"rows_done": n_rows, "input_tokens": int(off[-1]),
                           "output_tokens": int(boff[-1])})


_FRAME_ROWS = ["", "a", None, "héllo wörld", "x" * 40] + [f"row {i} " + "y" * ((i * 5) % 17)
                                                           for i in range(31)]


def _worker_frame(rank, world, port, balance, embedding, q, frame_rows=None):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from sutro_b200.sharding import infer_frame_sharded
        eng = _CpuEngine(rank, embedding)
        frame_rows = _FRAME_ROWS if frame_rows is None else frame_rows
        out = infer_frame_sharded(eng, frame_rows if rank == 0 else None, src=0, balance=balance,
                                  suffix="!")
        if rank == 0:
            q.put((out.get("outputs"), None if out.get("embeddings") is None
                   else out["embeddings"].tolist(), out["stats"]["rows_done"],
                   out["stats"]["n_gpus"]))
        else:
            assert out is None
        if len(frame_rows) >= world:              # rows keep their job-wide ids on every rank
            assert eng.seen_row_ids is not None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("balance,embedding", [("bytes", False), ("rows", False), ("bytes", True)])
def test_infer_frame_sharded_world2_is_positional(balance, embedding):
    """The product's multi-GPU path on gloo: broadcast of the resident column, per-rank row
    selection, per-rank run, padded gather, ordered merge — outputs[i] belongs to rows[i]."""
    import numpy as np
    from sutro_b200.sharding import shard_bounds as sb_, snake_assignment
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_frame, args=(r, 2, port, balance, embedding, q))
             for r in range(2)]
    for p in procs:
        p.start()
    outputs, emb, rows_done, n_gpus = q.get(timeout=180)
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    rows = ["" if r is None else r for r in _FRAME_ROWS]
    lens = [len(r.encode()) for r in rows]
    if balance == "bytes":
        shards = snake_assignment(lens, 2)
    else:
        shards = [np.arange(*sb_(len(rows), 2, r)) for r in range(2)]
    owner = {int(i): r for r, s in enumerate(shards) for i in s}
    assert rows_done == len(rows) and n_gpus == 2
    if embedding:
        assert [e[0] for e in emb] == [float(len(r)) for r in rows]      # positional
        assert [int(e[1]) for e in emb] == [owner[i] for i in range(len(rows))]
    else:
        assert outputs == [f"{rows[i].upper()}@{owner[i]}!" for i in range(len(rows))]
    assert {owner[i] for i in range(len(rows))} == {0, 1}


@pytest.mark.parametrize("world,frame_rows", [(3, None), (3, ["only one", "and another"])])
def test_infer_frame_sharded_odd_world_and_more_ranks_than_rows(world, frame_rows):
    """Three ranks (uneven shards), and a frame with fewer rows than ranks (a rank with an empty
    shard takes part in the broadcast and the gather and contributes nothing)."""
    from sutro_b200.sharding import snake_assignment
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_frame, args=(r, world, port, "bytes", False, q, frame_rows))
             for r in range(world)]
    for p in procs:
        p.start()
    outputs, _, rows_done, n_gpus = q.get(timeout=240)
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    rows = ["" if r is None else r for r in (_FRAME_ROWS if frame_rows is None else frame_rows)]
    shards = snake_assignment([len(r.encode()) for r in rows], world)
    owner = {int(i): r for r, s in enumerate(shards) for i in s}
    assert rows_done == len(rows) and n_gpus == world
    assert outputs == [f"{rows[i].upper()}@{owner[i]}!" for i in range(len(rows))]
    assert len(set(owner.values())) == min(world, len(rows))


def test_infer_frame_sharded_single_process_matches():
    from sutro_b200.sharding import infer_frame_sharded
    out = infer_frame_sharded(_CpuEngine(0), _FRAME_ROWS)
    rows = ["" if r is None else r for r in _FRAME_ROWS]
    assert out["outputs"] == [r.upper() + "@0" for r in rows]
    assert out["stats"]["n_gpus"] == 1 and out["t_total_s"] >= out["t_results_resident_s"]
    assert infer_frame_sharded(_CpuEngine(0), [])["outputs"] == []


def test_snake_assignment_equals_balanced_shards():
    import random
    from sutro_b200.sharding import snake_assignment
    rng = random.Random(3)
    for n in (0, 1, 5, 1000, 20001):
        costs = [int(rng.lognormvariate(4.5, 0.6)) % 300 for _ in range(n)]   # many ties
        for world in (1, 2, 3, 8):
            a = balanced_shards(costs, world)
            b = [s.tolist() for s in snake_assignment(costs, world)]
            assert a == b
