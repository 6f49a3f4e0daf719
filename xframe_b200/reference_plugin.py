"""`ProjectWorker` with the reference's own constructor: the module a maintainer drops in for projects/fxs/reconstruct.py.

    import xframe_b200.reference_plugin as reconstruct          # instead of xframe.projects.fxs.reconstruct
    worker = reconstruct.ProjectWorker()                          # no arguments: reads xframe.settings.project, xframe.database.project
    result, _ = worker.run()                                      # Controller.run -> job.run() (control/Control.py:72-75)

What it keeps from the reference worker (reconstruct.py:89-209):
  * constructor without arguments; settings from `xframe.settings.project` (DictNamespace -> plain dict), invariants through
    `database.project.load('invariants', path_modifiers={'structure_name', 'dimensions'})` (MTIP.load_mtip_data, :278-285),
  * `GPU.use` / `Multiprocessing.get_number_of_gpus()` check (:96-105) -- but where the reference falls back to the CPU with a warning,
    this worker raises: xframe_b200 has no CPU path,
  * `run()` returns `(result, locals())` with `result` an object array of per-run dicts in the schema of :1003-1021,
  * `post_processing()` ranks by the last main error and calls `db.save('reconstructions', {...})` with the reference's record (:160-183).
What changes: the reconstructions are not forked as processes (:141-157) but run as one device-resident batch
(`multi_process.n_parallel_reconstructions` keeps its meaning: the number of reconstructions).
"""
import logging
import time
import traceback

import numpy as np

from ._lib import XfbError
from .worker import ProjectWorker as _CudaWorker, assemble_reconstruction_record, number_of_gpus

log = logging.getLogger('root')


def _plain(node):
    """DictNamespace / dict tree of the reference's settings -> plain dicts (lists and arrays kept)."""
    if hasattr(node, 'dict') and callable(node.dict):
        node = node.dict()
    if isinstance(node, dict):
        return {k: _plain(v) for k, v in node.items()}
    if isinstance(node, (list, tuple)):
        return type(node)(_plain(v) for v in node)
    return node


def _read_number_of_processes(n_parallel):
    """Multiprocessing._read_number_of_processes (Multiprocessing.py:63-65,136-138): True / False -> the default count."""
    import os
    if isinstance(n_parallel, bool):
        return max(1, (os.cpu_count() or 2) // 2 - 1)
    return int(n_parallel)


class ProjectWorker:
    def __init__(self):
        from xframe import settings, database
        self.opt = settings.project
        self.db = database.project
        sd = _plain(settings.project)
        if not sd['GPU']['use']:
            raise XfbError('GPU.use is False: xframe_b200 has no CPU path')
        if number_of_gpus() == 0:
            raise XfbError('no CUDA device: the reference would fall back to the CPU here (reconstruct.py:96-102); xframe_b200 does not')
        data = self.db.load('invariants', path_modifiers={'structure_name': sd['structure_name'], 'dimensions': sd['dimensions']})
        data = dict(data)
        avg = data['average_intensity']
        if hasattr(avg, 'data'):                         # SampledFunction of the reference's loader (fxs_Projections.py:473-476,668)
            data['average_intensity'] = np.asarray(avg.data)
        self.data = data
        mp = sd['multi_process']
        n = _read_number_of_processes(mp.get('n_parallel_reconstructions', False)) if mp.get('use', True) else 1
        self.worker = _CudaWorker(sd, data, n_reconstructions=n)
        self.settings_dict = sd
        self.results = {'stats': {}}

    def post_processing(self):
        """reconstruct.py:160-183, including its catch-and-log of save errors."""
        try:
            fto = self.settings_dict['fourier_transform']
            record = assemble_reconstruction_record(list(self.results['MTIP']), self.results['stats'], self.data.get('xray_wavelength'),
                                                    fto.get('reciprocity_coefficient', np.pi))
            self.db.save('reconstructions', record)
        except Exception as e:      # noqa: BLE001
            log.info(f'Error during postprocessing / saving with message:\n {e}')
            log.debug(traceback.format_exc())

    def run(self):
        start = time.time()
        res, _ = self.worker.run()
        result = np.empty(len(res), dtype=object)
        for i, r in enumerate(res):
            result[i] = r
        self.results['MTIP'] = result
        self.results['stats'].update(self.worker.results['stats'])
        self.results['stats']['run_time'] = time.time() - start
        self.post_processing()
        return result, locals()
