"""Multi-GPU result path on hardware: `torchrun --nproc-per-node N tools/distributed_check.py` runs the golden small case
as 5 reconstructions sharded over the N ranks (xframe_b200.worker.run_distributed: final NCCL gather to rank 0, ranked
record like post_processing, reconstruct.py:160-183) and compares every gathered array with a single-process run on rank 0."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from helpers import load_golden, golden_settings, golden_data  # noqa: E402


def main():
    import copy
    from xframe_b200.worker import ProjectWorker, run_distributed, assemble_reconstruction_record
    g = load_golden('ref_small_ftstab')
    sd = copy.deepcopy(golden_settings(g))
    sd['GPU'] = {'use': True, 'n_gpu_workers': 1}
    n = 5
    rho0 = [g['rho0'] * (1.0 + 0.05 * i) for i in range(n)]
    rank = int(os.environ.get('RANK', 0))
    rec = run_distributed(sd, golden_data(g), n_reconstructions=n, initial_densities=rho0)
    if rank != 0:
        assert rec is None
        return
    w = ProjectWorker(sd, golden_data(g), n_reconstructions=n, initial_densities=rho0)
    res, _ = w.run()
    want = assemble_reconstruction_record(res, {}, 1.23984, 2.0)
    assert list(rec['reconstruction_results']) == list(want['reconstruction_results']), 'ranking differs'
    for k in want['reconstruction_results']:
        a, b = rec['reconstruction_results'][k], want['reconstruction_results'][k]
        for name in ('real_density', 'last_real_density', 'reciprocal_density', 'support_mask', 'last_deg2_invariant', 'initial_density'):
            assert np.array_equal(a[name], b[name]), (k, name)
        assert np.array_equal(a['error_dict']['main'], b['error_dict']['main']) and a['final_error'] == b['final_error']
        for u, v in zip(a['fxs_unknowns'], b['fxs_unknowns']):
            assert np.array_equal(u, v)
    errs = [rec['reconstruction_results'][k]['error_dict']['main'][-1] for k in rec['reconstruction_results']]
    assert errs == sorted(errs)
    print(f'distributed_check OK: world {os.environ.get("WORLD_SIZE", 1)}, {n} runs, ranking {list(rec["reconstruction_results"])}')


if __name__ == '__main__':
    main()
