"""Golden vectors for the radial-transform modes (midpoint, trapz, gauss) from the UNMODIFIED reference.

Run in the build container only:    python tests/golden/make_golden_hankel_modes.py

For every mode and for dimensions 3 and 2 the reference's own functions give
  * the radial grid pair                      ft_grid_pairs.py:274-299 (radial_grid_func_zernike [= trapz, :541], _midpoint, radial_grid_gauss)
  * the raw weights w[order, p, k]            hankel_transforms.py calc_{spherical,polar}_{trapz,mid,gauss}_weights
  * the assembled complex weights             hankel_transforms.assemble_weights (:540-553)
  * zht / izht of seeded coefficients         hankel_transforms.generate_ht, CPU flavour (:602-658)
stored as tests/golden/ref_hankel_modes.npz.  The worker functions are called directly (generate_weightDict only
farms them out to processes, hankel_transforms.py:386-391).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference  # noqa: E402

L_MAX, N_R, RC, Q_MAX = 6, 14, 2.0, 0.3


def main():
    import_reference()
    from xframe.projects.fxs.projectLibrary import hankel_transforms as H
    from xframe.projects.fxs.projectLibrary import ft_grid_pairs as G
    grids = {'midpoint': G.radial_grid_func_midpoint, 'trapz': G.radial_grid_func_zernike, 'gauss': G.radial_grid_gauss}
    workers = {3: {'midpoint': H.calc_spherical_mid_weights, 'trapz': H.calc_spherical_trapz_weights, 'gauss': H.calc_spherical_gauss_weights},
               2: {'midpoint': H.calc_polar_mid_weights, 'trapz': H.calc_polar_trapz_weights, 'gauss': H.calc_polar_gauss_weights}}
    rng = np.random.default_rng(11)
    out = {'l_max': L_MAX, 'n_r': N_R, 'rc': RC, 'q_max': Q_MAX}
    orders = np.arange(L_MAX + 1)
    for mode in ('midpoint', 'trapz', 'gauss'):
        g = grids[mode](Q_MAX, N_R, RC)
        rs, qs = (np.asarray(v() if callable(v) else v) for v in (g['real'], g['reciprocal']))      # the trapz/Zernike grid is a pair of closures (:274-280)
        out[f'{mode}_rs'], out[f'{mode}_qs'] = rs, qs
        r_max = rs.max()                                   # what reconstruct.py:329 passes on
        for dim in (3, 2):
            w = np.asarray(workers[dim][mode](orders, N_R, RC))
            out[f'{mode}_{dim}_weights'] = w
            aw = H.assemble_weights(w, orders, r_max, RC, dimensions=dim, mode=mode)
            out[f'{mode}_{dim}_forward'], out[f'{mode}_{dim}_inverse'] = np.asarray(aw['forward']), np.asarray(aw['inverse'])
            zht, izht = H.generate_ht(w, orders, r_max, reciprocity_coefficient=RC, dimensions=dim, use_gpu=False, mode=mode)
            if dim == 3:
                m_orders = np.concatenate((np.arange(L_MAX + 1), -np.arange(L_MAX, 0, -1)))
                cm = [rng.normal(size=(N_R, L_MAX - abs(m) + 1)) + 1j * rng.normal(size=(N_R, L_MAX - abs(m) + 1)) for m in m_orders]
                f, b = zht(cm), izht(cm)
                for i in range(len(m_orders)):
                    out[f'{mode}_3_in_{i}'], out[f'{mode}_3_zht_{i}'], out[f'{mode}_3_izht_{i}'] = cm[i], np.array(f[i]), np.array(b[i])
            else:
                c = rng.normal(size=(N_R, 2 * L_MAX + 1)) + 1j * rng.normal(size=(N_R, 2 * L_MAX + 1))
                out[f'{mode}_2_in'] = c
                out[f'{mode}_2_zht'] = np.array(zht(c.copy()))       # the 2-D CPU flavour returns a shared buffer (:616-624)
                out[f'{mode}_2_izht'] = np.array(izht(c.copy()))
    path = os.path.join(HERE, 'ref_hankel_modes.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
