"""Drop-in for the fxs `reconstruct` ProjectWorker (projects/fxs/reconstruct.py:89-209).

Same constructor / run() contract and result schema (reconstruct.py:1003-1021), but instead of forking one
process per reconstruction (reconstruct.py:141-157) it runs all requested reconstructions as device-resident
batches on the local GPU, sharded over ranks when launched under torchrun.

    settings : dict in the schema of settings/reconstruct/default_0.01.yaml (see xframe_b200/settings.py)
    data     : the invariants record `db.load('invariants')` returns
"""
import time

import numpy as np
import torch

from . import setup_host as S
from ._lib import XfbError
from .distributed import shard_run_ids
from .plan import Plan
from .reconstruct import run_schedule


def number_of_gpus():
    """Multiprocessing.get_number_of_gpus (Multiprocessing.py:892-898)."""
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


class ProjectWorker:
    def __init__(self, settings, data, n_reconstructions=None, rank=0, world=1, device=None, seeds=None, initial_densities=None):
        self.opt = settings
        self.data = data
        self.dims = int(settings['dimensions'])
        if self.dims not in (2, 3):
            raise XfbError(f"dimensions={settings['dimensions']} is not supported")
        mods = settings.get('output_density_modifiers', {})
        self.shift_to_center = bool(mods.get('shift_to_center', False))
        # 2-D: fix_orientation (with SO_freedom) = shift_to_center followed by the orientation fix (reconstruct.py:745-751)
        self.fix_orientation = bool(self.dims == 2 and mods.get('fix_orientation', False)
                                    and settings['projections']['reciprocal'].get('SO_freedom', {}).get('use', False))
        if self.fix_orientation:
            self.shift_to_center = True
        if not settings['GPU']['use']:
            raise XfbError("GPU.use is False: xframe_b200 has no CPU path (the reference falls back to CPU, reconstruct.py:96-102)")
        if number_of_gpus() == 0:
            raise XfbError("no CUDA device: xframe_b200 has no CPU fallback")
        mp = settings['multi_process']
        n = n_reconstructions
        if n is None:
            npr = mp.get('n_parallel_reconstructions', False)
            n = int(npr) if (mp.get('use', True) and not isinstance(npr, bool)) else 1
        self.n_runs = int(n)
        self.rank, self.world = rank, world
        self.run_ids = shard_run_ids(self.n_runs, rank, world)
        g = settings['grid']
        max_q = g['max_q']
        if not isinstance(max_q, float):                                       # reconstruct.py:258-261
            max_q = float(np.max(data['data_radial_points']))
        fto = settings['fourier_transform']
        batch = settings['GPU'].get('batch', 0) or len(self.run_ids)
        self.batch = max(1, min(int(batch), max(1, len(self.run_ids))))
        self.plan = Plan(int(g['max_order']), int(g['n_radial_points']), max_q, n_theta=g.get('n_theta', 0), n_phi=g.get('n_phi', 0),
                         reciprocity_coefficient=fto.get('reciprocity_coefficient', np.pi), ft_type=fto['type'],
                         max_batch=self.batch, device=device, dimensions=self.dims)
        setup = S.ProjectionSetup if self.dims == 3 else S.ProjectionSetup2D
        self.proj = setup(self.plan.qs, data, self.plan.l_max, settings['projections']['reciprocal'])
        self.proj.apply_to(self.plan)
        popt = settings['projections']['real']['projections']
        self.initial_support = S.initial_support(self.plan, popt['support']['initial_support'])
        err = settings['main_loop']['error']['methods']
        recip = list(err['reciprocal'].get('calculate', []))
        main_metrics = err.get('main', {}).get('metrics', {'real': ['l2_projection_diff'], 'reciprocal': []})
        if list(err['real']['calculate']) != ['l2_projection_diff'] or list(main_metrics.get('real', [])) != ['l2_projection_diff'] \
                or list(main_metrics.get('reciprocal', [])):
            raise XfbError("xframe_b200 drives the loop with the default main error (real: [l2_projection_diff])")
        if recip not in ([], ['deg2_invariant_l2_diff']) or (recip and self.dims != 3):
            raise XfbError(f"reciprocal error metrics {recip}: xframe_b200 computes deg2_invariant_l2_diff (3-D) only")
        self.deg2_metric = bool(recip)
        inside = err['real'].get('l2_projection_diff', {}).get('inside_initial_support', False)
        hio = settings['projections']['real']['HIO']
        self.plan.set_real(popt['apply'], self.initial_support, popt.get('value_threshold', {}).get('threshold', (False, False)),
                           popt.get('limit_imag', {}).get('threshold', 0.0), hio.get('considered_projections', ['all']), inside,
                           average_center_shells=int(popt.get('average_center', {}).get('max_radial_id', 1)))
        if self.deg2_metric:      # reference invariants V_l V_l^H of the final projection matrices (fxs_Projections.py:631-637)
            pm = self.proj.projection_matrices
            ref = np.zeros((self.plan.l_max + 1, self.plan.n_r, self.plan.n_r))
            for l, v in enumerate(pm):
                ref[l] = np.asarray(v).real @ np.asarray(v).real.T
            self.n_used_orders = len(pm)
            self.plan.set_deg2_reference(ref, self.proj.radial_mask, self.proj.number_of_particles)
        base = settings['GPU'].get('seed', None)
        self.seeds = seeds if seeds is not None else [None if base is None else base + i for i in range(self.n_runs)]
        self.initial_densities = initial_densities
        self.results = {'stats': {}}

    def _guess(self, run_id):
        if self.initial_densities is not None:
            return np.asarray(self.initial_densities[run_id], dtype=complex)
        rng = np.random.default_rng(self.seeds[run_id])    # None -> OS entropy, like reconstruct.py:1119
        return S.density_guess(self.plan, self.opt['density_guess'], self.opt['particle_radius'], self.proj.integrated_intensity, rng)

    def _shift_to_center(self, rho_hat, rho):
        """Output modifier `shift_to_center` (reconstruct.py:732-738) for a batch on the device:
        (rho_hat e^{+i q.c}, IFT(FT(rho) e^{+i q.c})), c = centre of mass of Re rho (misk.py:295-312,
        fxs_Projections.py:1419-1444).  One-off post-processing: transforms through the plan, elementwise work in torch."""
        plan = self.plan
        if not hasattr(self, '_cart'):
            ang = (plan.thetas, plan.phis) if self.dims == 3 else (plan.phis,)

            def cart(radial):
                g = np.meshgrid(radial, *ang, indexing='ij')
                if self.dims == 3:
                    r, t, p = g
                    return np.stack((np.cos(p) * r * np.sin(t), np.sin(p) * r * np.sin(t), r * np.cos(t)), axis=-1)
                r, p = g
                return np.stack((r * np.cos(p), r * np.sin(p)), axis=-1)
            w = plan.int_weight[:, :, None] * np.ones(plan.grid_shape) if self.dims == 3 else plan.int_weight
            self._cart = (torch.from_numpy(cart(plan.rs)).to(plan.device), torch.from_numpy(cart(plan.qs)).to(plan.device),
                          torch.from_numpy(np.ascontiguousarray(w)).to(plan.device))
        cart_r, cart_q, w = self._cart
        dims = tuple(range(1, rho.dim()))
        total = (w * rho.real).sum(dim=dims)
        total = torch.where(total == 0, torch.ones_like(total), total)
        center = (w[None, ..., None] * cart_r[None] * rho.real[..., None]).sum(dim=dims) / total[:, None]     # [nb, dims]
        phases = torch.exp(1j * (cart_q[None] * center.reshape((-1,) + (1,) * (rho.dim() - 1) + (self.dims,))).sum(-1))
        return rho_hat * phases, plan.ift((plan.ft(rho.contiguous()) * phases).contiguous()), center

    def run(self):
        t0 = time.time()
        plan, out = self.plan, []
        ang = (plan.thetas, plan.phis) if self.dims == 3 else (plan.phis,)
        rs = np.stack(np.meshgrid(plan.rs, *ang, indexing='ij'), axis=-1)
        qs = np.stack(np.meshgrid(plan.qs, *ang, indexing='ij'), axis=-1)
        masked_pm = self.proj.masked_projection_matrices()
        for b0 in range(0, len(self.run_ids), self.batch):
            ids = self.run_ids[b0:b0 + self.batch]
            rho0 = torch.from_numpy(np.stack([self._guess(i) for i in ids])).to(plan.device)
            if self.deg2_metric:
                from .reconstruct import iteration_count
                plan.mtip_enable_deg2_metric(True, iteration_count(self.opt)[0])
            res = run_schedule(plan, self.opt, rho0)
            deg2_hist = plan.mtip_deg2_errors(res['errors'].shape[1]).cpu().numpy()[..., :self.n_used_orders] if self.deg2_metric else None
            self.results['stats'].setdefault('nonfinite_iterations', {}).update({int(r): int(n) for r, n in zip(ids, res['nonfinite_iterations'])})
            unknowns = [plan.unknowns(k) for k in range(len(ids))]      # of the last mtip_start (reconstruct.py:523,1013)
            if self.shift_to_center:                                    # output modifier on the best and the last pair (:988-989)
                if self.fix_orientation and not hasattr(self, '_so_rot'):
                    self._so_rot = self.proj.remaining_so_rotation(plan.n_phi)
                for a, b in (('best_reciprocal', 'best_real'), ('last_reciprocal', 'last_real')):
                    rh, rr, _ = self._shift_to_center(torch.from_numpy(res[a]).to(plan.device), torch.from_numpy(res[b]).to(plan.device))
                    if self.fix_orientation:                            # fix_orientation sketch (:740-745, fxs_Projections.py:1081-1094)
                        rot_fn, orders_m = self._so_rot
                        rot = torch.tensor([rot_fn(u) for u in unknowns], dtype=torch.float64, device=plan.device)
                        ramp = torch.exp(1j * rot[:, None, None] * torch.from_numpy(orders_m.astype(np.float64)).to(plan.device)[None, None, :])
                        rh = plan.sht_inverse((plan.sht_forward(rh.contiguous()) * ramp).contiguous())
                        rr = plan.sht_inverse((plan.sht_forward(rr.contiguous()) * ramp).contiguous())
                    res[a], res[b] = rh.cpu().numpy(), rr.cpu().numpy()
            # last_deg2_invariant: B_l = I_l I_l^H of the last density (reconstruct.py:757-765,993)
            last = torch.from_numpy(res['last_real']).to(plan.device)
            fd = plan.ft(last)
            I = plan.sht_forward((fd * fd.conj()).real.to(torch.complex128).contiguous())
            if self.dims == 3:                             # B_l = I_l I_l^H on the device (grouped DMMA GEMM, csrc/procrustes.cuh)
                deg2 = plan.deg2_invariants(I).to(torch.complex128).cpu().numpy()
            else:                                          # B_m = I_m I_m^* (fxs_invariant_tools.py:906-914): outer products
                Im = I[..., :plan.l_max + 1]
                deg2 = torch.einsum('bqm,bpm->bmqp', Im, Im.conj()).cpu().numpy()
            for k, rid in enumerate(ids):
                n_it = res['errors'].shape[1]
                out.append({
                    'run_id': rid,
                    'real_density': res['best_real'][k], 'last_real_density': res['last_real'][k],
                    'reciprocal_density': res['best_reciprocal'][k], 'last_reciprocal_density': res['last_reciprocal'][k],
                    'final_error': float(res['best_error'][k]), 'initial_density': res['initial_density'][k],
                    'initial_support': self.initial_support.copy(),
                    'error_dict': {'main': res['errors'][k].copy(), 'real': {'l2_projection_diff': res['errors'][k].copy()},
                                   'reciprocal': {'deg2_invariant_l2_diff': deg2_hist[k].copy()} if self.deg2_metric else {}},
                    'support_mask': res['best_support'][k], 'last_support_mask': res['last_support'][k],
                    'loop_iterations': res['loop_iterations'], 'fxs_unknowns': unknowns[k],
                    'n_particles': np.array([[self.proj.number_of_particles]] * n_it), 'n_particles_gradients': np.array([]),
                    'n_particles_fraction': np.array([]),
                    'grid_pair': {'real_grid': rs, 'reciprocal_grid': qs}, 'projection_matrices': masked_pm,
                    'last_deg2_invariant': deg2[k],
                })
        self.results['MTIP'] = out
        self.results['stats']['run_time'] = time.time() - t0
        return out, {'worker': self}


# ---------------------------------------------------------------------------------------------------------------------
# post-processing and the multi-GPU entry point
# ---------------------------------------------------------------------------------------------------------------------
def assemble_reconstruction_record(results, stats, xray_wavelength, reciprocity_coefficient):
    """The dict the reference hands to db.save('reconstructions', ...) (post_processing, reconstruct.py:160-183; layout read
    back by _database_.py:223-390 and asserted in tests/test_fxs_integration.py:388-421): results ranked by their LAST main
    error, grid pair and projection matrices stored once."""
    results = [dict(r) for r in results]
    errors, grid_pair, projection_matrices = [], None, None
    for r in results:
        grid_pair = r.pop('grid_pair')
        projection_matrices = r.pop('projection_matrices')
        r.pop('run_id', None)
        errors.append(r['error_dict']['main'][-1])
    order = np.argsort(errors)
    return {'configuration': {'internal_grid': grid_pair, 'xray_wavelength': xray_wavelength, 'reciprocity_coefficient': reciprocity_coefficient},
            'reconstruction_results': {str(i): results[i] for i in order},
            'projection_matrices': projection_matrices, 'stats': stats}


_GATHERED = ('real_density', 'last_real_density', 'reciprocal_density', 'last_reciprocal_density', 'initial_density',
             'support_mask', 'last_support_mask', 'last_deg2_invariant')


def run_distributed(settings, data, n_reconstructions=None, seeds=None, initial_densities=None):
    """`torchrun --nproc-per-node N` entry: run i -> rank i mod N (reconstruct.py:141-157 forks one process per run instead),
    no per-iteration collective, one final gather of every run's arrays to rank 0 over NCCL (NVLink / NVSwitch) where the
    reference's post_processing builds the ranked record.  Returns the record on rank 0, None elsewhere."""
    import os
    import torch.distributed as dist
    from .distributed import gather_results
    rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    own_group = False
    if world > 1 and not dist.is_initialized():
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        own_group = True
    w = ProjectWorker(settings, data, n_reconstructions=n_reconstructions, rank=rank, world=world, device=local, seeds=seeds,
                      initial_densities=initial_densities)
    res, _ = w.run()
    dev = w.plan.device
    n_unk = sum(int(np.prod(u.shape)) for u in res[0]['fxs_unknowns']) if res else 0
    local_t = {k: torch.from_numpy(np.stack([np.asarray(r[k]) for r in res])).to(dev) for k in _GATHERED}
    local_t['main_error'] = torch.from_numpy(np.stack([r['error_dict']['main'] for r in res])).to(dev)
    local_t['final_error'] = torch.tensor([r['final_error'] for r in res], dtype=torch.float64, device=dev)
    if n_unk:
        local_t['fxs_unknowns'] = torch.from_numpy(np.stack([np.concatenate([np.asarray(u).ravel() for u in r['fxs_unknowns']]) for r in res])).to(dev)
    full = gather_results(local_t, w.n_runs, device=dev)
    record = None
    if rank == 0:
        shapes = [u.shape for u in res[0]['fxs_unknowns']]
        out = []
        for i in range(w.n_runs):
            r = dict(res[0])                      # run-independent entries (grid pair, projection matrices, initial support, ...)
            r.update({k: full[k][i].cpu().numpy() for k in _GATHERED})
            err = full['main_error'][i].cpu().numpy()
            r['error_dict'] = {'main': err, 'real': {'l2_projection_diff': err.copy()}, 'reciprocal': {}}
            r['final_error'] = float(full['final_error'][i])
            if n_unk:
                flat, unk, o = full['fxs_unknowns'][i].cpu().numpy(), [], 0
                for sh in shapes:
                    n = int(np.prod(sh))
                    unk.append(flat[o:o + n].reshape(sh))
                    o += n
                r['fxs_unknowns'] = tuple(unk)
            r['run_id'] = i
            out.append(r)
        fto = settings['fourier_transform']
        record = assemble_reconstruction_record(out, dict(w.results['stats']), data.get('xray_wavelength', None),
                                                fto.get('reciprocity_coefficient', np.pi))
    if own_group:
        dist.barrier()
        dist.destroy_process_group()
    w.plan.close()
    return record

