"""CPU, world_size 2 over gloo: the N>1 host logic -- shard bounds are disjoint and cover the corpus, the
max-over-ranks reduction bench.py uses, and per-rank .map() plumbing run in separate processes."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_items):
    sys.path.insert(0, os.path.join(ROOT, "modal-examples_b200"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import modal
        from b200rt.sharding import batches_of, shard_bounds

        b, e = shard_bounds(n_items, world)[rank]
        app = modal.App(f"gloo-{rank}")

        @app.function()
        def count(r):
            return r[1] - r[0]

        # each rank pumps its own shard through .map() in inputs of 32 (remainder dropped per shard, like the reference)
        done = sum(count.map(batches_of(e - b, 32)))
        t = torch.tensor([done, e - b, b, e], dtype=torch.int64)
        gathered = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        total = sum(int(g[1]) for g in gathered)
        assert total == n_items
        for r in range(1, world):
            assert int(gathered[r][2]) == int(gathered[r - 1][3])  # contiguous, disjoint
        assert int(gathered[0][2]) == 0 and int(gathered[-1][3]) == n_items
        assert done == ((e - b) // 32) * 32
        # timing reduction: max over ranks
        ms = torch.tensor([10.0 + rank], dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        assert float(ms) == 10.0 + world - 1
        dist.barrier()
        # bench.py's N > 1 control flow: rank 0 alone drives the pool while the other ranks wait on a host-side (gloo)
        # barrier; what rank 0 measured is then the job's figure on every rank
        host_pg = dist.new_group(backend="gloo")
        v = torch.tensor([123.0 if rank == 0 else 0.0], dtype=torch.float64)
        dist.barrier(group=host_pg)
        if rank == 0:
            v += 1.0  # "drive the pool"
        dist.barrier(group=host_pg)
        dist.broadcast(v, src=0)
        assert float(v) == 124.0
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_sharding_over_gloo():
    mp.spawn(_worker, args=(2, _free_port(), 100_003), nprocs=2, join=True)


def test_shard_and_wave_split_properties():
    sys.path.insert(0, os.path.join(ROOT, "modal-examples_b200"))
    from b200rt.sharding import batches_of, shard_bounds, wave_split

    for n in (0, 1, 7, 8, 1000, 1_000_000):
        for w in (1, 2, 4, 8):
            sb = shard_bounds(n, w)
            assert sb[0][0] == 0 and sb[-1][1] == n and all(a[1] == b[0] for a, b in zip(sb, sb[1:]))
            assert max(e - b for b, e in sb) - min(e - b for b, e in sb) <= 1
            for first in range(w):
                ws = wave_split(n, w, first)
                assert len(ws) == w and sum(e - b for b, e in ws) == n and all(e >= b for b, e in ws)
                live = sorted((b, e) for b, e in ws if e > b)
                assert all(a[1] == b[0] for a, b in zip(live, live[1:])) and (not live or (live[0][0] == 0 and live[-1][1] == n))
                if n >= 8 * w:
                    assert len(live) == w, "a wave with >= 8 items per replica uses the whole pool"
                elif live:
                    assert min(e - b for b, e in live[:-1] or live) >= 1 and len(live) == min(w, max(1, n // 8))
    from b200rt.sharding import bucket_of, length_runs

    assert [bucket_of(x) for x in (1, 64, 65, 128, 449, 512)] == [64, 64, 128, 128, 512, 512]
    lens = [512, 3, 70, 64, 500, 65, 1]
    runs = length_runs(lens)
    assert [b for b, _ in runs] == [64, 128, 512] and [idx for _, idx in runs] == [[1, 3, 6], [2, 5], [0, 4]]
    assert sorted(i for _, idx in runs for i in idx) == list(range(len(lens)))
    assert all(bucket_of(lens[i]) == b for b, idx in runs for i in idx)
    assert list(batches_of(100, 32)) == [(0, 32), (32, 64), (64, 96)]
    assert list(batches_of(100, 32, drop_remainder=False))[-1] == (96, 100)
