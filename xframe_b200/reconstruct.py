"""Host driver of the batched, device-resident MTIP phasing loop.

Mirrors ``MTIP.assemble_phasing_loop`` of the reference (projects/fxs/reconstruct.py:768-1036): the nested
sub-loop / method schedule, the beta ramp, the shrink-wrap sigma / threshold ramps and the enforce-initial-support
rule are evaluated here as host scalars; every array operation runs in the CUDA library (xframe_b200/csrc) on a
batch of independent reconstructions.  Nothing here touches numpy arrays of grid size between init and output.
"""
import numpy as np

from .plan import HIO, ER
from .ramps import ExponentialRamp, LinearRamp
from ._lib import XfbError

_METHODS = {'HIO': HIO, 'ER': ER, 'HIO_non_FXS': HIO, 'ER_non_FXS': ER}


def _sw_ramps(opt, default_sigma):
    """generate_update_shrink_wrap (reconstruct.py:1212-1258)."""
    sw_opt = opt['projections']['real']['shrink_wrap']
    names = opt['main_loop']['sub_loops']['order']
    sig, thr = [], []
    for lid in range(len(names)):
        s = sw_opt['sigmas'][lid] if len(sw_opt['sigmas']) - 1 >= lid else False
        if not isinstance(s, (list, tuple)):
            s = [s]
        sig.append(LinearRamp(*s, default_start=default_sigma, default_stop=default_sigma))
        t = sw_opt['thresholds'][lid] if len(sw_opt['thresholds']) - 1 >= lid else 0.1
        if not isinstance(t, (list, tuple)):
            t = [t]
        thr.append(LinearRamp(*t))
    return sig, thr


class _ShrinkWrapScalars:
    """sigma / threshold setters of ShrinkWrapParts (fxs_Projections.py:215-243)."""

    def __init__(self, default_sigma):
        self.default_sigma = default_sigma
        self.sigma = default_sigma
        self.threshold = 0.06

    def set_threshold(self, v):
        self.threshold = 0 if v < 0 else (1 if v >= 1 else v)

    def set_sigma(self, v):
        ok = (np.issubdtype(np.array(v).dtype, np.number) and not isinstance(v, bool)) and v > 0
        self.sigma = v if ok else self.default_sigma


def iteration_count(opt):
    """(n HIO/ER iterations, n SW steps) of the whole schedule."""
    n_it = n_sw = 0
    loops = opt['main_loop']['sub_loops']
    for name in loops['order']:
        lopt = loops[name]
        for key in lopt['order']:
            mo = lopt['methods'][key]
            rep = mo['iterations'] if isinstance(mo, dict) else mo
            if key in ('SW', 'SW_center'):
                n_sw += lopt['iterations'] * (rep if key == 'SW_center' else 1)
            else:
                n_it += lopt['iterations'] * rep
    return n_it, n_sw


def run_schedule(plan, opt, rho0, default_sigma=None, collect=True):
    """Run the whole schedule of ``opt['main_loop']`` on the batch ``rho0`` [nb, N_r, n_theta, n_phi] (CUDA complex128).

    Returns numpy arrays per run when ``collect`` (the result-dict payload of reconstruct.py:1003-1021), else nothing
    (bench mode: everything stays on the device).
    """
    loops = opt['main_loop']['sub_loops']
    hio_opt = opt['projections']['real']['HIO']
    sup_opt = opt['projections']['real']['projections']['support']['enforce_initial_support']
    if default_sigma is None:
        default_sigma = np.pi / plan.qs.max()                       # fxs_Projections.py:189-193
    sig_ramps, thr_ramps = _sw_ramps(opt, default_sigma)
    sw = _ShrinkWrapScalars(default_sigma)

    def update_sw(it, lid):
        if not sig_ramps[lid].undefined:
            sw.set_sigma(sig_ramps[lid](it))
        if not thr_ramps[lid].undefined:
            sw.set_threshold(thr_ramps[lid](it))

    plan.mtip_init(rho0)
    initial = plan.mtip_grid('last_real').cpu().numpy() if collect else None
    has_non_fxs = any(k.endswith('_non_FXS') for name in loops['order'] for k in loops[name]['order'])
    iterations = []
    for lid, name in enumerate(loops['order']):
        lopt = loops[name]
        beta = hio_opt['beta'][lid] if len(hio_opt['beta']) - 1 >= lid else [0.5, 0.5, -1 / 700, 1600]
        beta_ramp = ExponentialRamp(*beta)
        limit = sup_opt['if_error_bigger_than'] if sup_opt['apply'] else np.inf
        n_first = lopt.get('best_density_not_in_first_n_iterations', np.inf)
        if 'SW' in lopt['order']:
            update_sw(0, lid)
        if has_non_fxs:
            plan.mtip_snapshot_intensity()        # `hist` is bound at the start of the sub-loop (reconstruct.py:853)
        fixed = False                             # latest_intensity (:862,899-904)
        step = sw_step = 0
        it = 0
        for it in range(1, lopt['iterations'] + 1):
            plan.mtip_set_outer_iteration(it)
            for key in lopt['order']:
                mo = lopt['methods'][key]
                repeats = mo['iterations'] if isinstance(mo, dict) else mo
                if key == 'SW':
                    plan.mtip_shrinkwrap(sw.sigma, sw.threshold, limit)
                    sw_step += 1
                    update_sw(sw_step, lid)
                elif key == 'SW_center':          # :886-897 (the reference's exchanged pair included, see oracle/mtip.py)
                    for _ in range(repeats):
                        plan.mtip_shrinkwrap_center(sw.sigma, sw.threshold, limit)
                        sw_step += 1
                        update_sw(sw_step, lid)
                elif key in _METHODS:
                    ft_stab = mo.get('ft_stab', False) if isinstance(mo, dict) else False
                    if not isinstance(ft_stab, bool):
                        raise XfbError(f"ft_stab: '{ft_stab}' is not supported by xframe_b200 (True / False)")
                    non_fxs = key.endswith('_non_FXS')
                    if non_fxs and not fixed:
                        plan.mtip_fix_intensity()
                    fixed = non_fxs
                    plan.mtip_set_non_fxs(non_fxs)
                    betas = [beta_ramp.eval(step + i) for i in range(repeats)]
                    if has_non_fxs and repeats > 0:
                        # `hist` is re-bound at the start of every iteration (:911): after the block it names the pair its LAST
                        # iteration started from
                        if repeats > 1:
                            plan.mtip_iterate(_METHODS[key], ft_stab, betas[:-1])
                        plan.mtip_snapshot_intensity()
                        plan.mtip_iterate(_METHODS[key], ft_stab, betas[-1:])
                    elif repeats > 0:
                        plan.mtip_iterate(_METHODS[key], ft_stab, betas)
                    step += repeats
                else:
                    raise XfbError(f"method '{key}' is not supported by xframe_b200 (HIO, ER, HIO_non_FXS, ER_non_FXS, SW, SW_center)")
        plan.mtip_set_non_fxs(False)
        if np.isfinite(n_first):                  # :945-949
            plan.mtip_select_best(int(np.floor(n_first)))
        iterations.append(it)
    if not collect:
        return None
    hist, best = plan.mtip_errors()
    return {
        'errors': hist.cpu().numpy(), 'best_error': best.cpu().numpy(), 'initial_density': initial,
        'last_real': plan.mtip_grid('last_real').cpu().numpy(), 'last_reciprocal': plan.mtip_grid('last_reciprocal').cpu().numpy(),
        'best_real': plan.mtip_grid('best_real').cpu().numpy(), 'best_reciprocal': plan.mtip_grid('best_reciprocal').cpu().numpy(),
        'last_support': plan.mtip_grid('last_support').cpu().numpy(), 'best_support': plan.mtip_grid('best_support').cpu().numpy(),
        'loop_iterations': int(np.sum(iterations) + 1), 'nonfinite_iterations': plan.mtip_nonfinite(),
    }
