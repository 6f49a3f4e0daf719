"""Sharding of independent reconstructions over the GPUs of one box.

One process per GPU (torchrun); run i goes to rank i mod world (the reference's analogue:
gpu = (client_pid // n_control_workers) % n_gpus, Multiprocessing.py:1275-1277).  There is NO per-iteration
collective: the only exchange is the final gather of the finished results to rank 0, which sorts them by final
error as post_processing does (reconstruct.py:170-177).  Backend 'nccl' on GPUs (NVLink/NVSwitch), 'gloo' in the
CPU tests.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_run_ids(n_runs, rank, world):
    return [i for i in range(n_runs) if i % world == rank]


def gather_results(local, n_runs, device=None):
    """local: dict name -> tensor [n_local, ...] (same trailing shapes and dtypes on every rank, run order as
    shard_run_ids).  Returns on rank 0 a dict name -> tensor [n_runs, ...] in global run order; None elsewhere."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return {k: v.clone() for k, v in local.items()}
    world, rank = dist.get_world_size(), dist.get_rank()
    counts = [len(shard_run_ids(n_runs, r, world)) for r in range(world)]
    cap = max(counts)
    out = {}
    for name in sorted(local):
        t = local[name]
        if device is not None:
            t = t.to(device)
        was_bool = t.dtype == torch.bool
        if was_bool:
            t = t.to(torch.uint8)
        is_cplx = t.is_complex()
        if is_cplx:
            t = torch.view_as_real(t)
        pad = torch.zeros((cap,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[:t.shape[0]] = t
        bufs = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
        dist.gather(pad, bufs, dst=0)                      # ncclGather over NVLink / NVSwitch on the GPU box
        if rank == 0:
            full = torch.empty((n_runs,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            for r in range(world):
                ids = shard_run_ids(n_runs, r, world)
                if ids:
                    full[torch.as_tensor(ids, device=t.device)] = bufs[r][:len(ids)]
            if is_cplx:
                full = torch.view_as_complex(full)
            if was_bool:
                full = full.bool()
            out[name] = full
    return out if rank == 0 else None


def sort_by_error(final_errors):
    """Ranking used when results are stored: ascending final error (reconstruct.py:175-177)."""
    return np.argsort(np.asarray(final_errors))
