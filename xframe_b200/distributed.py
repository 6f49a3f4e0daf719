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


def gather_results(local, n_runs, device=None, chunk_bytes=1 << 30):
    """local: dict name -> tensor [n_local, ...] (same trailing shapes and dtypes on every rank, run order as
    shard_run_ids).  Returns on rank 0 a dict name -> HOST tensor [n_runs, ...] in global run order; None elsewhere.

    The runs travel rank by rank in pieces of at most chunk_bytes (NCCL send / recv over NVLink / NVSwitch on the GPU box, gloo in
    the CPU test) through one staging buffer on rank 0 and land in host memory, where post_processing works on them
    (reconstruct.py:160-183): at L=127 / N_r=256 the 512 densities of a full job are 64 GiB -- more than is free next to the
    working set of rank 0's own shard."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return {k: v.detach().cpu().clone() for k, v in local.items()}
    world, rank = dist.get_world_size(), dist.get_rank()
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
        t = t.contiguous()
        tail = tuple(t.shape[1:])
        per_run = max(1, int(np.prod(tail)) * t.element_size())
        step = max(1, int(chunk_bytes // per_run))
        full = torch.empty((n_runs,) + tail, dtype=t.dtype) if rank == 0 else None
        for r in range(world):
            ids = shard_run_ids(n_runs, r, world)
            for k0 in range(0, len(ids), step):
                part = ids[k0:k0 + step]
                if r == 0:
                    if rank == 0:
                        for j, i in enumerate(part):           # straight into the run's (contiguous) row of the host result
                            full[i].copy_(t[k0 + j])
                elif rank == r:
                    dist.send(t[k0:k0 + len(part)].contiguous(), dst=0)
                elif rank == 0:
                    buf = torch.empty((len(part),) + tail, dtype=t.dtype, device=t.device)
                    dist.recv(buf, src=r)
                    for j, i in enumerate(part):
                        full[i].copy_(buf[j])
        if rank == 0:
            if is_cplx:
                full = torch.view_as_complex(full)
            if was_bool:
                full = full.bool()
            out[name] = full
    return out if rank == 0 else None


def sort_by_error(final_errors):
    """Ranking used when results are stored: ascending final error (reconstruct.py:175-177)."""
    return np.argsort(np.asarray(final_errors))
