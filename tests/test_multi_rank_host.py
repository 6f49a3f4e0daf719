"""world_size-2 gloo test (CPU) of the only multi-GPU logic the path has: sharding of independent runs and the
final gather to rank 0.  There is no per-iteration collective to test."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from xframe_b200.distributed import shard_run_ids, gather_results, sort_by_error


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_runs, out_q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    ids = shard_run_ids(n_runs, rank, world)
    # each "result" encodes its global run id so the gather order can be checked
    local = {
        'density': torch.stack([torch.full((2, 3), complex(i, -i), dtype=torch.complex128) for i in ids]) if ids else torch.zeros((0, 2, 3), dtype=torch.complex128),
        'error': torch.tensor([10.0 - i for i in ids], dtype=torch.float64),
        'mask': torch.stack([torch.tensor([i % 2 == 0, True]) for i in ids]) if ids else torch.zeros((0, 2), dtype=torch.bool),
    }
    full = gather_results(local, n_runs)
    if rank == 0:
        out_q.put({k: v.numpy() for k, v in full.items()})
    else:
        assert full is None
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_gather_world2():
    n_runs, world = 5, 2
    assert shard_run_ids(n_runs, 0, world) == [0, 2, 4] and shard_run_ids(n_runs, 1, world) == [1, 3]
    assert sorted(sum((shard_run_ids(128, r, 8) for r in range(8)), [])) == list(range(128))
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_runs, q)) for r in range(world)]
    for p in procs:
        p.start()
    full = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert full['density'].shape == (5, 2, 3)
    assert np.array_equal(full['density'][:, 0, 0], np.array([complex(i, -i) for i in range(5)]))
    assert np.array_equal(full['error'], 10.0 - np.arange(5))
    assert np.array_equal(full['mask'][:, 0], np.arange(5) % 2 == 0)
    assert list(sort_by_error(full['error'])) == [4, 3, 2, 1, 0]          # reconstruct.py:175-177
