"""Host <-> device copy ceiling of the e2e leg, all ranks at once (torchrun --nproc-per-node N tools/copy_probe.py [runs_total]).

Every rank moves its share of `runs_total` densities (16 MiB each at L=63 / N_r=128) from pinned host memory to its GPU and back,
H2D and D2H concurrently on two streams -- exactly the traffic of one `xfb_mtip_step_host` without any kernel.  Rank 0 prints
one JSON line: aggregate GB/s per direction (max over ranks of the device-timed span) and the step rate that ceiling allows.
"""
import json
import os
import sys

import torch
import torch.distributed as dist


def main():
    total = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    nb = len([i for i in range(total) if i % world == rank])
    per_run = 128 * 64 * 128                      # complex128 grid points of one density
    h_in = torch.empty((nb, per_run), dtype=torch.complex128).pin_memory()
    h_out = torch.empty_like(h_in).pin_memory()
    d_in = torch.empty((nb, per_run), dtype=torch.complex128, device='cuda')
    d_out = torch.zeros_like(d_in)
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    res = {}
    for mode in ('h2d', 'd2h', 'both'):
        reps = 5
        for it in range(reps + 1):
            if it == 1:
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                s_in.wait_event(e0)
                s_out.wait_event(e0)
            if mode in ('h2d', 'both'):
                with torch.cuda.stream(s_in):
                    d_in.copy_(h_in, non_blocking=True)
            if mode in ('d2h', 'both'):
                with torch.cuda.stream(s_out):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s_in)
        torch.cuda.current_stream().wait_stream(s_out)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[mode] = float(t.item())
    if rank == 0:
        gb = total * per_run * 16 / 1e9
        print(json.dumps({'probe': 'pinned host <-> device copies of one e2e step, all ranks concurrently', 'n_gpus': world, 'runs_total': total,
                          'GB_per_direction_per_step': gb, 'ms_h2d_only': res['h2d'], 'ms_d2h_only': res['d2h'], 'ms_both_directions': res['both'],
                          'GBps_h2d_only': gb / res['h2d'] * 1e3, 'GBps_d2h_only': gb / res['d2h'] * 1e3,
                          'GBps_per_direction_full_duplex': gb / res['both'] * 1e3,
                          'e2e_ceiling_iterations_per_s': total / (res['both'] * 1e-3)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
