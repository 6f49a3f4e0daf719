"""Run the UNMODIFIED reference (European-XFEL/xFrame) in-process, for tests and for `bench.py --impl reference`.

The reference package is taken from `baseline/_ref/` (the git-ignored pip install made by `__graft_entry__.build()`; it travels
to the GPU box) or, in the build container only, straight from `/root/reference`.  What the harness provides is what the
reference's own start-up would (startup_routines.py:57,63; Multiprocessing.py:32-40; control/Control.py:62-87):

  * the setuptools_scm artefact `xframe/_version.py` is stubbed when the install lacks it (pyproject.toml:64-67),
  * HOME points at a scratch directory (xframe/logger.py:6-9 writes ~/.xframe/log.txt),
  * a spherical-harmonic plugin class is injected at the reference's own slot `xframe.library.mathLibrary.shtns`
    (shtns itself is absent from this image): `oracle.sht.sh` (numpy) or `xframe_b200.harmonic_transforms.sh` (CUDA),
  * optionally the GPU-access layer is routed to `xframe_b200.gpu_access` (CudaPlugin, comm_module),
  * `settings.project` / `database.project` are populated by hand (no ruamel / h5py here).
"""
import importlib
import os
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))      # repo root (this file lives in baseline/)
_CANDIDATES = (os.path.join(ROOT, 'baseline', '_ref'), '/root/reference')


def reference_root():
    for c in _CANDIDATES:
        if os.path.isdir(os.path.join(c, 'xframe', 'projects', 'fxs')):
            return c
    return None


def import_reference(sh_class=None, cuda_gpu_layer=False):
    """Import xframe from reference_root(); returns the package.  sh_class: plugin injected at the shtns slot."""
    root = reference_root()
    if root is None:
        raise ImportError('no reference package (baseline/_ref or /root/reference)')
    os.environ['HOME'] = tempfile.mkdtemp(prefix='xf_home_')
    if not os.path.exists(os.path.join(root, 'xframe', '_version.py')):
        m = types.ModuleType('xframe._version')
        m.__version__ = m.version = '0.0.0'
        sys.modules['xframe._version'] = m
    if root not in sys.path:
        sys.path.insert(0, root)
    import warnings
    warnings.filterwarnings('ignore')
    cwd = os.getcwd()
    import xframe
    import xframe.library.mathLibrary as mLib
    if sh_class is not None:
        mLib.shtns = sh_class
    if cuda_gpu_layer:
        import xframe.Multiprocessing as MP
        from xframe_b200 import gpu_access
        MP.openCL_plugin = gpu_access.CudaPlugin                 # Multiprocessing.py:32-40
        MP.comm_module.add_gpu_process = gpu_access.comm_module.add_gpu_process      # communicators.py:79-82
        MP.get_number_of_gpus = gpu_access.get_number_of_gpus    # Multiprocessing.py:892-898
    os.chdir(cwd)
    return xframe


class FakeDB:
    """Minimal stand-in for projects/fxs/_database_.py used by MTIP.preinit (reconstruct.py:241-285)."""

    def __init__(self, data):
        self.data = data
        self.saved = {}

    def load(self, name, **kw):
        if name == 'invariants':
            return self.data
        raise FileNotFoundError(name)

    def save(self, name, data, *a, **k):
        self.saved[name] = data

    def update_settings(self, *a, **k):
        pass


def fresh_data(inv):
    """Invariants record as the reference's loader hands it over.  The reference regrids
    average_intensity in place (fxs_Projections.py:668), so every MTIP gets its own copy."""
    from xframe.library.gridLibrary import SampledFunction, NestedArray
    data = dict(inv)
    src = inv['data_projection_matrices']
    if isinstance(src, np.ndarray) and src.dtype != object:        # 2-D: one array [M+1, n_q]
        pm = np.array(src, dtype=complex)
    else:                                                          # 3-D: one [n_q, n_l] matrix per order
        pm = np.empty(len(src), dtype=object)
        for i, p in enumerate(src):
            pm[i] = np.array(p, dtype=complex)
    data['data_projection_matrices'] = pm
    data['average_intensity'] = SampledFunction(NestedArray(np.asarray(inv['data_radial_points'])[:, None].copy(), 1),
                                                np.asarray(inv['average_intensity']).copy(), coord_sys='cartesian')
    return data


def preinit(settings_dict, inv):
    """Master-process part (ProjectWorker.__init__, reconstruct.py:89-110): populate the reference's globals and run
    MTIP.preinit() (grids, Fourier-transform weights -- its weight workers may only be requested from the master process,
    Multiprocessing.py:813).  Returns the reconstruct module."""
    from xframe.library.pythonLibrary import DictNamespace
    from xframe import settings
    import xframe.database as database
    settings.project = DictNamespace.dict_to_dictnamespace(settings_dict)
    settings.general.cache_aware = False
    settings.general.n_control_workers = 0
    database.project = FakeDB(fresh_data(inv))
    cwd = os.getcwd()
    name = 'xframe.projects.fxs.reconstruct'
    rec = importlib.reload(sys.modules[name]) if name in sys.modules else importlib.import_module(name)
    os.chdir(cwd)                                   # the module does os.chdir(plugin_dir) at import (reconstruct.py:7-9)
    rec.set_globals()
    rec.MTIP.preinit()
    return rec


def new_mtip(rec, inv, rho0=None):
    """Per-reconstruction part (setup_phasing_loop, reconstruct.py:113-115,141-148; runs in the child processes of the
    reference): a fresh MTIP object with its phasing loop.  rho0: injected initial density (the reference draws it from
    os.urandom, reconstruct.py:1119)."""
    from xframe.library.pythonLibrary import RecipeFactory
    if rho0 is not None:
        rec.MTIP.generate_density_guess_method = lambda self, spec, grid: (lambda: np.array(rho0, dtype=complex))
    rec.MTIP.mtip_data = fresh_data(inv)
    m = rec.MTIP(RecipeFactory({}))
    m.generate_phasing_loop()
    return m


def make_mtip(settings_dict, inv, rho0=None):
    """preinit + new_mtip in one process.  Returns (reconstruct module, MTIP instance)."""
    rec = preinit(settings_dict, inv)
    return rec, new_mtip(rec, inv, rho0)
