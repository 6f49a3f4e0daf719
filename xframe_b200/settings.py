"""Settings of `fxs reconstruct`: the reference's YAML schema is part of the drop-in boundary.

`default_settings()` is the resolved content of settings/reconstruct/default_0.01.yaml (3-D values of the `_if`
switches); `load_settings(path)` merges a user file such as settings/reconstruct/tutorial.yaml over it the way the
reference's SettingsParser does for plain keys (database/database.py:643-684): recursive dict merge, `command:`
strings evaluated with numpy in scope (database.py:500-506).  Unknown keys are kept and ignored.
"""
import copy

import numpy as np


def default_settings():
    return {
        'dimensions': 3, 'structure_name': 'default_structure', 'particle_radius': 150,
        'grid': {'max_q': False, 'max_order': 63, 'n_phi': 0, 'n_theta': 0, 'n_radial_points': 128},
        'fourier_transform': {'type': 'midpoint', 'reciprocity_coefficient': 2.0, 'allow_weight_calculation': True,
                              'allow_weight_saving': True},
        'density_guess': {'type': 'bump', 'bump': {'slope': 0.3}, 'low_resolution_autocorrelation': {'threshold_to_max': 0.01},
                          'radius': None, 'amplitude_function': 'random', 'random': {'SNR': 2}},
        'projections': {
            'real': {
                'projections': {
                    'apply': ['support', 'value_threshold', 'assert_real'],
                    'value_threshold': {'threshold': [0, False]},
                    'limit_imag': {'threshold': 2},
                    'support': {'initial_support': {'type': 'max_radius', 'max_radius': None, 'auto_correlation': {'threshold': 0.01}},
                                'enforce_initial_support': {'apply': True, 'if_error_bigger_than': 6e-3}},
                },
                'shrink_wrap': {'sigmas': [[False, [False, False], False], [False, [False, False], False]],
                                'thresholds': [[0.08, [0, 0], 0], [0.08, [0, 0], 0]]},
                'HIO': {'beta': [[0.5, 0.4, -1 / 700, 1600], [0.01, 0.002, -1 / 200, 200]], 'considered_projections': ['all']},
            },
            'reciprocal': {
                'number_of_particles': {'initial': 1.0, 'estimate': False},
                'regrid': {'interpolation': 'cubic'},
                'used_order_ids': np.arange(64),
                'odd_orders_to_0': True, 'use_averaged_intensity': True,
                'q_mask': {'type': 'none'},
                'SO_freedom': {'use': None, 'radial_high_pass': 0.2},     # None: resolved from `dimensions` (False in 3-D, True in 2-D)
            },
        },
        'output_density_modifiers': {'shift_to_center': False, 'fix_orientation': None},   # fix_orientation: True in 2-D only
        'main_loop': {
            'error': {'methods': {
                'real': {'calculate': ['l2_projection_diff'], 'l2_projection_diff': {'inside_initial_support': True}},
                'reciprocal': {'calculate': []},
                'main': {'metrics': {'real': ['l2_projection_diff'], 'reciprocal': []}, 'type': 'mean'}},
                'limits': {'use': False}, 'gain_limits': {'use': False}},
            'sub_loops': {
                'order': ['main', 'refinement'],
                'main': {'methods': {'HIO': {'iterations': 60, 'ft_stab': True}, 'ER': {'iterations': 40, 'ft_stab': True}, 'SW': 1},
                         'order': ['HIO', 'SW', 'ER'], 'iterations': 5, 'best_density_not_in_first_n_iterations': np.inf},
                'refinement': {'methods': {'ER': {'iterations': 100, 'ft_stab': True}, 'SW': 1},
                               'order': ['SW', 'ER'], 'iterations': 2, 'best_density_not_in_first_n_iterations': np.inf},
            }},
        'GPU': {'use': True, 'n_gpu_workers': 1, 'batch': 0, 'seed': None},
        'multi_process': {'use': True, 'n_parallel_reconstructions': False},
        'profiling': {'enable': False, 'reconstruction_process_id': 1, 'gpu_worker_id': -1},
    }


def tutorial_overrides():
    """settings/reconstruct/tutorial.yaml:1-72 (max_order set to the 63 the docs and BASELINE.json quote; the
    file itself says 64 while constraining orders 0..63)."""
    return {
        'structure_name': 'tutorial', 'dimensions': 3, 'particle_radius': 250,
        'grid': {'n_radial_points': 128, 'max_order': 63},
        'density_guess': {'type': 'bump', 'bump': {'slope': 0.3}, 'amplitude_function': 'random', 'random': {'SNR': 2}},
        'projections': {
            'real': {
                'shrink_wrap': {'sigmas': [[20, [False, 5], -2], False], 'thresholds': [0.09, 0.09]},
                'HIO': {'beta': [[0.5, 0.4, -1 / 250, 500], [0.01, 0.002, -1 / 200, 200]]},
                'projections': {'apply': ['support', 'value_threshold', 'limit_imag'],
                                'support': {'initial_support': {'type': 'max_radius'},
                                            'enforce_initial_support': {'apply': True, 'if_error_bigger_than': 6e-3}},
                                'value_threshold': {'threshold': [0, False]}, 'limit_imag': {'threshold': 2}}},
            'reciprocal': {'number_of_particles': {'initial': 1}, 'use_averaged_intensity': True, 'q_mask': {'type': 'none'},
                           'used_order_ids': np.arange(64)}},
        'multi_process': {'n_parallel_reconstructions': True},
        'GPU': {'use': True},
        'main_loop': {'sub_loops': {
            'order': ['main', 'refinement'],
            'main': {'methods': {'HIO': {'iterations': 60}, 'ER': {'iterations': 40}, 'SW': {'iterations': 1}},
                     'iterations': 5, 'order': ['HIO', 'SW', 'ER']},
            'refinement': {'methods': {'ER': {'iterations': 100}, 'SW': {'iterations': 1}}, 'iterations': 1, 'order': ['SW', 'ER']}}},
    }


def _resolve_commands(node):
    if isinstance(node, dict):
        if set(node.keys()) == {'command'}:
            return eval(node['command'], {'np': np})    # same contract as the reference: settings are trusted code
        return {k: _resolve_commands(v) for k, v in node.items()}
    if isinstance(node, list):
        return [_resolve_commands(v) for v in node]
    return node


def merge(base, over):
    out = copy.deepcopy(base)
    for k, v in over.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict):
            out[k] = merge(out[k], v)
        else:
            out[k] = copy.deepcopy(v)
    return out


def finalize(opt):
    """`_copy` links of the defaults file (radius <- particle_radius, max_radius <- particle_radius)."""
    opt = copy.deepcopy(opt)
    if opt['density_guess'].get('radius') is None:
        opt['density_guess']['radius'] = opt['particle_radius']
    sup = opt['projections']['real']['projections']['support']['initial_support']
    if sup.get('max_radius') is None:
        sup['max_radius'] = opt['particle_radius']
    # `_if` / `_only_if` switches on /dimensions (default_0.01.yaml:185-200)
    so = opt['projections']['reciprocal'].setdefault('SO_freedom', {})
    if so.get('use') is None:
        so['use'] = opt['dimensions'] == 2
    so.setdefault('radial_high_pass', 0.2)
    mods = opt.setdefault('output_density_modifiers', {})
    if mods.get('fix_orientation') is None:
        mods['fix_orientation'] = opt['dimensions'] == 2
    if opt['dimensions'] != 2:
        mods['fix_orientation'] = False
    return opt


def tutorial_settings(**over):
    return finalize(merge(merge(default_settings(), tutorial_overrides()), over))


def load_settings(path):
    import yaml
    with open(path) as f:
        user = _resolve_commands(yaml.safe_load(f) or {})
    return finalize(merge(default_settings(), user))
