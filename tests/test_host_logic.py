"""CPU tests of the host side of xframe_b200 (no GPU): tables, ramps, settings, setup, schedule driver, C-ABI exports."""
import os
import re

import numpy as np
import pytest

from helpers import load_golden, golden_settings, golden_data, rel_l2, ROOT
from oracle import mtip as O
from oracle.sht import normalized_legendre, gauss_grid
from xframe_b200 import tables, ramps, setup_host as S, settings as ST
from xframe_b200.reconstruct import run_schedule, iteration_count
from jacobi_model import jacobi_project, qr_polar


def test_library_exports_every_declared_symbol():
    from xframe_b200 import _lib
    lib = _lib.load()                       # raises if the .so is missing or a symbol of EXPORTS is absent
    hdr = open(os.path.join(ROOT, 'include', 'xfb200.h')).read()
    declared = set(re.findall(r'\b(xfb_[a-z0-9_]+)\s*\(', hdr))
    assert declared, 'no declarations parsed'
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name)


def test_missing_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from xframe_b200.plan import Plan
    from xframe_b200._lib import XfbError
    with pytest.raises(XfbError):
        Plan(7, 16, 0.04)


def test_legendre_tables_match_oracle():
    L, nt, nph = 15, 16, 32
    packed, NP = tables.pack_legendre(L, nt, nph)
    K2 = nt // 2
    tab = (L + 1) * K2 * NP
    FE, FO = packed[:tab].reshape(L + 1, K2, NP), packed[tab:2 * tab].reshape(L + 1, K2, NP)
    IE, IO = packed[2 * tab:3 * tab].reshape(L + 1, NP, K2), packed[3 * tab:].reshape(L + 1, NP, K2)
    x, w = gauss_grid(nt)
    P = normalized_legendre(L, x)
    for m in range(L + 1):
        pe, po = P[m][0::2, :K2], P[m][1::2, :K2]
        assert np.allclose(IE[m, :pe.shape[0]], pe, rtol=1e-13, atol=1e-15)
        assert np.allclose(IO[m, :po.shape[0]], po, rtol=1e-13, atol=1e-15)
        assert np.allclose(FE[m, :, :pe.shape[0]].T, pe * w[:K2] * 2 * np.pi / nph, rtol=1e-13, atol=1e-16)
        assert np.allclose(FO[m, :, :po.shape[0]].T, po * w[:K2] * 2 * np.pi / nph, rtol=1e-13, atol=1e-16)
        assert not IE[m, pe.shape[0]:].any() and not FO[m, :, po.shape[0]:].any()
        # mirror symmetry used by the kernels: P_l^m(-x) = (-1)^(l+m) P_l^m(x)
        full = P[m]
        sign = (-1.0) ** (np.arange(full.shape[0]))[:, None]
        assert np.allclose(full[:, ::-1], sign * full, atol=1e-13)


@pytest.mark.parametrize('mode', ['midpoint', 'trapz', 'gauss'])
def test_hankel_and_grids_match_oracle(mode):
    assert np.array_equal(tables.hankel_weights(9, 12, 2.0, mode), O.hankel_weights(9, 12, 2.0, mode))
    a, b = tables.radial_grids(mode, 0.3, 12, 2.0), O.radial_grids(mode, 0.3, 12, 2.0)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    w = O.assemble_weights(O.hankel_weights(3, 12, 2.0, mode), a[0].max(), 2.0, mode)
    fs, iscale = tables.hankel_scales(a[0].max(), 12, 2.0, mode)
    assert np.allclose(w['forward'][..., 1], (-1j) * fs * np.moveaxis(O.hankel_weights(3, 12, 2.0, mode), 0, 2)[..., 1])
    assert np.allclose(w['inverse'][..., 2], (1j) ** 2 * iscale * np.moveaxis(O.hankel_weights(3, 12, 2.0, mode), 0, 2)[..., 2])


def test_integration_weights_match_spherical_integrator():
    g = load_golden('ref_small_ftstab')
    m = O.MTIP(golden_settings(g), golden_data(g))
    wt = tables.integration_weights(m.rs, len(m.sh.theta))
    f = np.random.default_rng(0).random(m.real_grid.shape[:-1])
    assert abs(np.sum(wt[:, :, None] * f) - m.integrator.integrate(f)) < 1e-12 * abs(m.integrator.integrate(f))


def test_ramps_match_oracle():
    for args in [(0.5, 0.4, -1 / 250, 500), (0.01, 0.002, -1 / 200, 200), (0.3, 0.6, 0.01, 100)]:
        a, b = ramps.ExponentialRamp(*args), O.ExponentialRamp(*args)
        assert all(abs(a(x) - b(x)) < 1e-15 for x in range(0, 700, 7))
    cases = [((20, [False, 5], -2), dict(default_start=9.7, default_stop=9.7)), ((False,), dict(default_start=9.7, default_stop=9.7)),
             ((0.09,), {}), (([0.08, [0, 0], 0]), {}), ((0.1, [0.05, 4]), {}), ((False,), {})]
    for args, kw in cases:
        a, b = ramps.LinearRamp(*args, **kw), O.LinearRamp(*args, **kw)
        assert a.undefined == b.undefined
        for x in range(8):
            va, vb = a(x), b(x)
            assert (np.isnan(va) and np.isnan(vb)) or abs(va - vb) < 1e-15


def test_projection_setup_and_guess_match_oracle():
    g = load_golden('ref_small_ftstab')
    sd = golden_settings(g)
    m = O.MTIP(sd, golden_data(g))
    ps = S.ProjectionSetup(m.qs, golden_data(g), m.l_max, sd['projections']['reciprocal'])
    for a, b in zip(ps.projection_matrices, m.rp.projection_matrices):
        assert np.array_equal(a, b)
    assert np.array_equal(ps.radial_mask, m.rp.radial_mask)
    assert ps.integrated_intensity == m.rp.integrated_intensity

    class P:
        pass
    p = P()
    p.rs, p.qs, p.thetas, p.phis = m.rs, m.qs, m.sh.theta, m.sh.phi
    p.grid_shape = m.real_grid.shape[:-1]
    p.int_weight = tables.integration_weights(m.rs, len(m.sh.theta))
    assert np.array_equal(S.initial_support(p, sd['projections']['real']['projections']['support']['initial_support']), m.real_pr.initial_support)
    assert rel_l2(S.density_guess(p, sd['density_guess'], 250.0, ps.integrated_intensity, np.random.default_rng(7)), g['rho0']) < 1e-14
    assert np.array_equal(S.six_sphere_density(p), O.six_sphere_density(m.real_grid))


def test_settings_loader_reads_reference_schema(tmp_path):
    y = tmp_path / 'tutorial.yaml'
    y.write_text("""
structure_name: 'tutorial'
particle_radius: 250
grid:
  n_radial_points: 128
  max_order: 63
projections:
  real:
    shrink_wrap:
      sigmas: [[20,[False,5],-2],False]
      thresholds: [0.09,0.09]
    HIO:
      beta:
        command: '[[0.5,0.4,-1/250,500],[0.01,0.002,-1/200,200]]'
    projections:
      apply: [support,value_threshold,limit_imag]
  reciprocal:
    used_order_ids:
      command: 'np.arange(64)'
main_loop:
  sub_loops:
    refinement:
      iterations: 1
""")
    o = ST.load_settings(str(y))
    assert o['projections']['real']['HIO']['beta'][0] == [0.5, 0.4, -1 / 250, 500]
    assert list(o['projections']['reciprocal']['used_order_ids']) == list(range(64))
    assert o['density_guess']['radius'] == 250
    assert o['projections']['real']['projections']['support']['initial_support']['max_radius'] == 250
    assert o['main_loop']['sub_loops']['main']['methods']['HIO'] == {'iterations': 60, 'ft_stab': True}
    assert iteration_count(o) == (600, 6)                       # 5x(60 HIO + 40 ER) + 100 ER ; 6 SW
    assert iteration_count(ST.tutorial_settings()) == (600, 6)


class FakePlan:
    """Records the device calls of run_schedule (host logic only)."""

    def __init__(self):
        self.calls = []
        self.qs = np.linspace(0.001, 0.32, 16)

    def mtip_init(self, rho0):
        self.calls.append(('init',))

    def mtip_iterate(self, method, ft_stab, betas):
        self.calls.append(('it', method, ft_stab, tuple(betas)))

    def mtip_shrinkwrap(self, sigma, thr, limit):
        self.calls.append(('sw', sigma, thr, limit))

    def mtip_grid(self, which):
        raise AssertionError('collect=False must not read grids')

    # sketch / option tail: recorded under their own kinds
    def mtip_set_outer_iteration(self, it):
        self.outer = it

    def mtip_set_non_fxs(self, on):
        self.non_fxs = bool(on)

    def mtip_snapshot_intensity(self):
        self.calls.append(('snap',))

    def mtip_fix_intensity(self):
        self.calls.append(('fix',))

    def mtip_select_best(self, n_first):
        self.calls.append(('best', n_first))

    def mtip_shrinkwrap_center(self, sigma, thr, limit):
        self.calls.append(('swc', sigma, thr, limit))


def test_run_schedule_follows_reference_schedule():
    sd = ST.tutorial_settings()
    fp = FakePlan()
    run_schedule(fp, sd, None, collect=False)
    kinds = [c[0] for c in fp.calls]
    assert kinds == ['init'] + ['it', 'sw', 'it'] * 5 + ['sw', 'it']
    beta0, beta1 = O.ExponentialRamp(0.5, 0.4, -1 / 250, 500), O.ExponentialRamp(0.01, 0.002, -1 / 200, 200)
    its = [c for c in fp.calls if c[0] == 'it']
    step = 0
    for k, c in enumerate(its[:10]):                       # main loop: step counts HIO+ER within the sub-loop
        n = 60 if k % 2 == 0 else 40
        assert c[1] == (0 if k % 2 == 0 else 1) and c[2] is True and len(c[3]) == n
        assert np.allclose(c[3], [beta0.eval(step + i) for i in range(n)], rtol=1e-15)
        step += n
    assert np.allclose(its[10][3], [beta1.eval(i) for i in range(100)], rtol=1e-15)
    sws = [c for c in fp.calls if c[0] == 'sw']
    ds = np.pi / fp.qs.max()
    assert [round(s[1], 9) for s in sws[:5]] == [round(max(20 - 2 * i, ds), 9) for i in range(5)]   # sigma ramp 20,-2/step, floor default
    assert abs(sws[5][1] - ds) < 1e-12 and all(s[2] == 0.09 and s[3] == 6e-3 for s in sws)


def test_run_schedule_sketch_tail():
    """SW_center, non-FXS blocks and a finite best_density_not_in_first_n_iterations (reconstruct.py:886-904,945-949): the
    intensity snapshot is taken at the start of a sub-loop and before the LAST iteration of every block (the pair the
    reference's `hist` variable names afterwards), the fixed intensity is set when a non-FXS block follows a FXS one."""
    import copy
    sd = copy.deepcopy(ST.tutorial_settings())
    sd['main_loop']['sub_loops'] = {
        'order': ['main'],
        'main': {'iterations': 2, 'order': ['HIO', 'SW_center', 'HIO_non_FXS', 'ER_non_FXS', 'ER'], 'best_density_not_in_first_n_iterations': 1,
                 'methods': {'HIO': {'iterations': 3, 'ft_stab': True}, 'SW_center': {'iterations': 2}, 'HIO_non_FXS': {'iterations': 2},
                             'ER_non_FXS': {'iterations': 1}, 'ER': {'iterations': 2}}}}
    fp = FakePlan()
    run_schedule(fp, sd, None, collect=False)
    kinds = [c[0] for c in fp.calls]
    one = ['it', 'snap', 'it', 'swc', 'swc', 'fix', 'it', 'snap', 'it', 'snap', 'it', 'it', 'snap', 'it']
    assert kinds == ['init', 'snap'] + one + one + ['best']
    its = [c for c in fp.calls if c[0] == 'it']
    assert [len(c[3]) for c in its[:7]] == [2, 1, 1, 1, 1, 1, 1] and fp.calls[-1] == ('best', 1) and fp.non_fxs is False
    assert iteration_count(sd) == (16, 4)


@pytest.mark.parametrize('tag', ['ref_medium_ops'])
def test_jacobi_algorithm_reproduces_reference_projection(tag):
    """The device algorithm (real basis + one-sided Jacobi + cut-off), run in numpy, against the reference's SVD result."""
    g = load_golden(tag)
    sd = golden_settings(g)
    m = O.MTIP(sd, golden_data(g))
    L = m.l_max
    splits = np.arange(1, L + 1) ** 2
    I = np.split(g['I_direct'], splits, axis=1)
    Ip = np.split(g['Iproj_direct'], splits, axis=1)
    for l in range(2, L + 1, 2):
        T, sweeps = jacobi_project(m.rp.projection_matrices[l], m.qs, I[l])
        assert sweeps < 30
        assert rel_l2(T, Ip[l]) < 1e-6, (l, rel_l2(T, Ip[l]))
        if l <= 4:
            assert rel_l2(T, Ip[l]) < 1e-10
        # QR-preconditioned variant (the path the 512-thread kernel takes): same answer, fewer sweeps
        Tq, sweeps_q = qr_polar(m.rp.projection_matrices[l], m.qs, I[l])
        assert sweeps_q <= sweeps and rel_l2(Tq, Ip[l]) < 1e-6, (l, sweeps_q, sweeps, rel_l2(Tq, Ip[l]))
        assert rel_l2(Tq, T) < 1e-7


def test_selective_reorthogonalisation_matches_always_twice():
    """The device QR repeats a projection only where it cancelled more than half of the squared norm.  On graded /
    rank-deficient columns it reconstructs A to rounding and loses no more orthogonality than the always-twice variant
    (right-looking Gram-Schmidt loses cond * eps either way; the Jacobi stage works on R, not on Q^T Q)."""
    from jacobi_model import mgs2
    rng = np.random.default_rng(3)
    for m, r, decades in ((40, 25, 12), (96, 70, 30), (64, 64, 8), (30, 20, 2)):
        U = np.linalg.qr(rng.normal(size=(m, r)))[0]
        Vt = np.linalg.qr(rng.normal(size=(r, r)))[0]
        A = (U * np.logspace(0, -decades, r)) @ Vt
        Q0, R0 = mgs2(A, selective=False)
        Q1, R1 = mgs2(A, selective=True)
        assert np.abs(Q1 @ R1 - A).max() < 1e-14 * np.abs(A).max()
        k = int((np.abs(np.diag(R0)) > 1e-13 * R0[0, 0]).sum())           # numerically resolved directions
        e0 = np.abs(Q0[:, :k].T @ Q0[:, :k] - np.eye(k)).max()
        e1 = np.abs(Q1[:, :k].T @ Q1[:, :k] - np.eye(k)).max()
        assert e1 < 10 * e0 + 1e-14, (decades, e0, e1)


REF_DEFAULTS = '/root/reference/xframe/projects/fxs/settings/reconstruct/default_0.01.yaml'


@pytest.mark.skipif(not os.path.exists(REF_DEFAULTS), reason='reference settings file not available (build container only)')
def test_default_settings_transcribe_the_reference_yaml():
    """settings.default_settings() against the reference's own defaults file, parsed here with its `_value` / `command` / `_copy`
    conventions (database.py:500-506,643-684).  Every leaf that default_settings() carries must equal the YAML's value, except the
    documented GPU additions and the `_if` switches on /dimensions that finalize() resolves."""
    import yaml
    with open(REF_DEFAULTS) as f:
        raw = yaml.safe_load(f)
    COPY = object()

    def resolve(node):
        if isinstance(node, dict):
            if '_if' in node:
                return COPY                                       # switch on another setting (e.g. /dimensions): resolved by finalize()
            if '_value' in node:
                v = node['_value']
                if isinstance(v, dict) and set(v) == {'command'}:
                    try:
                        return eval(v['command'], {'np': np})
                    except NameError:                             # refers to the running framework (e.g. Multiprocessing.free_cpus)
                        return COPY
                if isinstance(v, dict) and '_copy' in v:
                    return COPY
                if isinstance(v, dict) and ('_if' in v or any(k.startswith('_') for k in v)):
                    return COPY                                   # switch on another setting: resolved by finalize()
                return resolve(v)
            return {k: resolve(v) for k, v in node.items() if not k.startswith('_')}
        if isinstance(node, list):
            return [resolve(v) for v in node]
        if isinstance(node, str):                                 # PyYAML (YAML 1.1) leaves '6e-3' a string; ruamel (YAML 1.2) reads a float
            try:
                return float(node)
            except ValueError:
                return node
        return node
    ref = resolve(raw)
    ours = ST.default_settings()
    skipped, checked = [], [0]

    def same(a, b):
        if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
            return np.array_equal(np.asarray(a), np.asarray(b))
        if isinstance(a, (list, tuple)) and isinstance(b, (list, tuple)):
            return len(a) == len(b) and all(same(x, y) for x, y in zip(a, b))
        if isinstance(a, bool) or isinstance(b, bool):
            return a is b
        if isinstance(a, (int, float)) and isinstance(b, (int, float)):
            return a == b or abs(float(a) - float(b)) <= 1e-15 * max(1.0, abs(float(b)))
        return a == b and isinstance(a, bool) == isinstance(b, bool)

    def walk(o, r, path):
        for k, v in o.items():
            pth = path + '/' + k
            if pth in ('/GPU/batch', '/GPU/seed'):               # the only additions
                continue
            if pth == '/projections/reciprocal/SO_freedom/radial_high_pass':     # default lives in the code: .get('radial_high_pass', 0.2)
                continue
            assert k in r, f'{pth} is not a key of the reference defaults'
            if isinstance(v, dict) and isinstance(r[k], dict):
                walk(v, r[k], pth)
            elif r[k] is COPY or v is None:                      # `_copy` links / dimension switches: finalize() fills them
                skipped.append(pth)
            else:
                assert same(v, r[k]), f'{pth}: {v!r} != reference {r[k]!r}'
                checked[0] += 1
    walk(ours, ref, '')
    assert checked[0] >= 40, (checked[0], skipped)

