"""Shared test helpers: golden-case loading and settings reconstruction."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))


def load_golden(tag):
    return dict(np.load(os.path.join(GOLDEN, tag + '.npz'), allow_pickle=False))


def golden_settings(g):
    from make_golden import settings_dict
    sd = settings_dict(int(g['n_r']), int(g['l_max']), int(g['n_theta']), int(g['n_phi']), float(g['max_q']), bool(g['ft_stab']))
    sd['output_density_modifiers'] = {'shift_to_center': bool(g['shift_to_center']) if 'shift_to_center' in g else False}
    return sd


def golden_data(g):
    l_max = int(g['l_max'])
    return {'dimensions': 3, 'xray_wavelength': 1.23984, 'average_intensity': g['avg_intensity'],
            'data_radial_points': g['data_q'], 'data_angular_points': g['phis'], 'max_order': l_max,
            'data_projection_matrices': [g[f'pm_{l}'] for l in range(l_max + 1)], 'number_of_particles': 1}


def rel_l2(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))
