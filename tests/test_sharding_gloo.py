"""CPU, world_size 2 over gloo: the N>1 path of bench.py / the sharding helpers.  Frames shard with no data-path
collective; the only communication is the timing barrier + MAX reduction and (here) a gather used to check coverage."""
import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_units, out):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sh = importlib.import_module("amos-slam_b200.sharding")
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = sh.shard_range(n_units, rank, world)
    mine = torch.zeros(n_units, dtype=torch.int32); mine[b:e] = 1
    dist.all_reduce(mine)                                       # test-only: every unit owned exactly once
    t = torch.tensor([float(10 + rank)], dtype=torch.float64)   # max-over-ranks timing reduction as in bench.py
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.barrier()
    if rank == 0:
        out.put((mine.tolist(), float(t.item())))
    dist.destroy_process_group()


def test_two_rank_sharding_covers_all_units():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 513, q)) for r in range(2)]
    [p.start() for p in procs]
    owned, tmax = q.get(timeout=120)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert owned == [1] * 513 and tmax == 11.0


def test_shard_helpers():
    sys.path.insert(0, ROOT)
    sh = importlib.import_module("amos-slam_b200.sharding")
    for n in (0, 1, 7, 512, 513):
        for w in (1, 2, 4, 8):
            parts = [sh.shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            assert max(e - b for b, e in parts) - min(e - b for b, e in parts) <= 1
    assert [sh.gpu_for_sequence(s, 8) for s in range(10)] == [0, 1, 2, 3, 4, 5, 6, 7, 0, 1]
    assert sh.sequences_of_gpu(10, 1, 8) == [1, 9]
    with pytest.raises(ValueError):
        sh.shard_range(4, 2, 2)
    # the pool's rule (C ABI, no device needed): gpu = seq mod G, stream = (seq div G) mod S -- every (gpu, stream) worker gets its share
    assert [sh.worker_for_sequence(s, 4, 2) for s in range(9)] == [(0, 0), (1, 0), (2, 0), (3, 0), (0, 1), (1, 1), (2, 1), (3, 1), (0, 0)]
    with pytest.raises(ValueError):
        sh.worker_for_sequence(-1, 4, 2)
