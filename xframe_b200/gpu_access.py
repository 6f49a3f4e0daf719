"""GPU-access layer of xFrame for the fxs path, on CUDA.

The reference reaches the GPU through an OpenCL plugin that lives in worker processes and is driven by RPC with
shared-memory copies (xframe/Multiprocessing.py:885-1290, externalLibraries/openCL_plugin.py:42-384,
control/communicators.py:79-82).  Here the CUDA library is called in-process; this module keeps the reference's
surface so `hankel_transforms.generate_spherical_ht_gpu`-style callers work unchanged:

    get_number_of_gpus()                                      Multiprocessing.py:890-898
    CudaPlugin.create_context / get_number_of_gpus / ClProcess / ClFunction /
               create_process_buffers_on_all_gpus              openCL_plugin.py:42-61,302-384
    comm_module.add_gpu_process(proc) -> callable(*arrays)     communicators.py:79-82, Multiprocessing.py:1247-1262

The one deliberate restriction: a ClProcess names a PRECOMPILED kernel.  Arbitrary OpenCL source in
kernel_dict['kernel'] cannot be honoured; the function name selects the device routine
('apply_weights' = the Hankel contraction of hankel_transforms.py:660-766, 'apply_matrix' = the matrix-times-vectors demo of
the reference's framework test, tests/test_framework_integration.py:230-309) and anything else raises XfbError.
"""
import ctypes as C

import numpy as np

from ._lib import XfbError, load, check


def get_number_of_gpus():
    """Number of CUDA devices (xfb_device_count).  Raises XfbError when the library is missing."""
    n = C.c_int(0)
    check(load().xfb_device_count(C.byref(n)))
    return int(n.value)


def _split_phase(w, inverse):
    """Complex [p, k, l] weights = real W[l, p, k] * (-+i)^l ?  Returns W or None."""
    nl = w.shape[2]
    ph = ((1.j if inverse else -1.j) ** np.arange(nl))[None, None, :]
    r = w / ph
    scale = np.abs(w).max()
    if scale == 0 or np.abs(r.imag).max() <= 1e-13 * scale:
        return np.ascontiguousarray(np.moveaxis(r.real, 2, 0))
    return None


class ClFunction:
    """openCL_plugin.py:64-122: bookkeeping of one kernel function (roles, dtypes, shapes, constant inputs)."""

    def __init__(self, func_data, kernel=None, context_data=False):
        d = func_data
        self.dict, self.kernel = d, kernel
        self.name = d['name']
        self.dtypes, self.shapes, self.arg_roles, self.const_inputs = d['dtypes'], d['shapes'], d['arg_roles'], d['const_inputs']
        self.global_range, self.local_range = d.get('global_range'), d.get('local_range')
        self.n_args = len(self.arg_roles)
        ids = lambda role: [i for i in range(self.n_args) if self.arg_roles[i] == role]   # noqa: E731
        self.input_ids, self.output_ids = ids('input'), ids('output')
        self.input_dtypes = [self.dtypes[i] for i in self.input_ids]
        self.input_shapes = [self.shapes[i] for i in self.input_ids]
        self.output_dtypes = [self.dtypes[i] for i in self.output_ids]
        self.output_shapes = [self.shapes[i] for i in self.output_ids]
        self.n_inputs, self.n_outputs = len(self.input_ids), len(self.output_ids)
        self.call = None


class ClProcess:
    """openCL_plugin.py:302-358.  `run` is assembled by `assemble_run()` (the reference does this in the GPU worker)."""

    _KNOWN = ('apply_weights', 'apply_matrix')

    def __init__(self, process_data, context_data=False, device=None):
        self.dict = process_data
        self.name = process_data['name']
        self.kernel = process_data.get('kernel')
        self.functions = [ClFunction(fd, self.kernel, context_data) for fd in process_data['functions']]
        self.n_functions = len(self.functions)
        f0, fl = self.functions[0], self.functions[-1]
        self.input_dtypes, self.input_shapes, self.n_inputs = f0.input_dtypes, f0.input_shapes, f0.n_inputs
        self.output_dtypes, self.output_shapes, self.n_outputs = fl.output_dtypes, fl.output_shapes, f0.n_outputs
        self.device = device
        self.run = None
        for f in self.functions:
            if f.name not in self._KNOWN:
                raise XfbError(f"ClProcess '{self.name}': kernel function '{f.name}' is not a precompiled xframe_b200 routine "
                               f"(known: {self._KNOWN}); arbitrary OpenCL source cannot be run on the CUDA backend")
        if self.n_functions != 1:
            raise XfbError("ClProcess: chained functions are not supported by the CUDA backend")

    def _assemble_apply_matrix(self, f):
        """apply_matrix(out[nq, nvec], matrix[nq, nq], vect[nq, nvec], nq, nvec): the GPU demo of the reference's framework test
        (tests/test_framework_integration.py:230-309) on the precompiled CUDA kernel xfb_apply_matrix."""
        import torch
        lib = load()
        mat = np.ascontiguousarray(np.asarray(f.const_inputs[1], dtype=np.float64))
        nq, nvec = (int(v) for v in f.shapes[0])
        if mat.shape != (nq, nq) or tuple(f.shapes[2]) != (nq, nvec):
            raise XfbError(f"apply_matrix: unexpected shapes matrix{mat.shape} out({nq},{nvec})")
        dev = torch.device('cuda', torch.cuda.current_device() if self.device is None else self.device)
        mat_d = torch.from_numpy(mat).to(dev)

        def run(vect):
            was_torch = isinstance(vect, torch.Tensor)
            x = vect if was_torch else torch.from_numpy(np.ascontiguousarray(vect, dtype=np.float64))
            x = x.to(device=dev, dtype=torch.float64).contiguous()
            if tuple(x.shape) != (nq, nvec):
                raise ValueError(f"apply_matrix input shape {tuple(x.shape)} != {(nq, nvec)}")
            out = torch.empty((nq, nvec), dtype=torch.float64, device=dev)
            with torch.cuda.device(dev):
                check(lib.xfb_apply_matrix(C.c_void_p(mat_d.data_ptr()), C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), nq, nvec,
                                           C.c_void_p(torch.cuda.current_stream().cuda_stream)))
            return out if was_torch else out.cpu().numpy()
        self.run = run
        return run

    def assemble_run(self):
        f = self.functions[0]
        if f.name == 'apply_matrix':
            return self._assemble_apply_matrix(f)
        # apply_weights(out[nq, nlm], w[n_sum, nq, nl], rho[nq, nlm], nq, nlm, nl)   hankel_transforms.py:672-740
        w = np.asarray(f.const_inputs[1])
        nq, nlm = (int(v) for v in f.shapes[0])
        nl = int(f.const_inputs[5])
        l_max = nl - 1
        if w.ndim != 3 or w.shape[1] != nq or w.shape[2] != nl or nlm != (l_max + 1) ** 2 or w.shape[0] not in (nq, nq - 1):
            raise XfbError(f"apply_weights: unexpected shapes w{w.shape} out({nq},{nlm}) nl={nl}")
        inverse, W = False, _split_phase(w, False)
        if W is None:
            inverse, W = True, _split_phase(w, True)
        if W is None:
            raise XfbError("apply_weights: weights are not real Hankel weights times (-i)^l or (+i)^l per order")
        from .plan import Plan
        plan = Plan(l_max, nq, 1.0, max_batch=1, device=self.device, hankel_weights=W, hankel_scales=(1.0, 1.0))
        self._plan = plan

        def run(rho):
            import torch
            was_torch = isinstance(rho, torch.Tensor)
            x = rho if was_torch else torch.from_numpy(np.ascontiguousarray(rho, dtype=np.complex128))
            x = x.to(device=plan.device, dtype=torch.complex128).contiguous()
            if tuple(x.shape) != (nq, nlm):
                raise ValueError(f"apply_weights input shape {tuple(x.shape)} != {(nq, nlm)}")
            out = plan.hankel(x[None], inverse=inverse)[0]
            return out if was_torch else out.cpu().numpy()      # the reference also hands back a fresh copy (Multiprocessing.py:1072)

        self.run = run
        return run


class CudaPlugin:
    """Same members as OpenClPlugin (openCL_plugin.py:42-61,360-384; Multiprocessing_interfaces.py:44-59)."""
    cl_state = False
    contexts_created = False
    ClFunction = ClFunction
    ClProcess = ClProcess

    @classmethod
    def create_context(cls, allow_master_process=False):
        """CUDA contexts are per process and created lazily by the runtime; record the visible devices."""
        n = get_number_of_gpus()
        if n == 0:
            raise XfbError("no CUDA device visible")
        cls.cl_state = [{'device': i} for i in range(n)]
        cls.contexts_created = True

    @staticmethod
    def get_number_of_gpus():
        return get_number_of_gpus()

    @classmethod
    def create_process_buffers_on_all_gpus(cls, process):
        """openCL_plugin.py:360-384: one runnable copy of the process per GPU."""
        if not cls.contexts_created:
            cls.create_context()
        procs = []
        for st in cls.cl_state:
            p = ClProcess(process.dict, device=st['device'])
            p.assemble_run()
            procs.append(p)
        return procs


class _CommModule:
    """The slice of Multiprocessing.comm_module the fxs path uses."""

    def __init__(self):
        self.gpu_processes = {}

    def add_gpu_process(self, gpu_process):
        """Multiprocessing.py:1247-1262: register the process, return a blocking callable(*arrays) -> array | tuple."""
        run = gpu_process.run or gpu_process.assemble_run()
        self.gpu_processes[gpu_process.name] = gpu_process
        return run

    def get_number_of_gpus(self):
        return get_number_of_gpus()


comm_module = _CommModule()
openCL_plugin = CudaPlugin          # name under which Multiprocessing exposes the plugin (Multiprocessing.py:32-40)
