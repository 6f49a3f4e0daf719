"""ctypes binding of include/xfb200.h.  Fails loudly when the CUDA library is missing:
there is no CPU fallback on the product path."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('XFB200_LIB', os.path.join(_HERE, 'lib', 'libxfb200.so'))   # override: kernel experiments only


class XfbError(RuntimeError):
    pass


class PlanDesc(C.Structure):
    _fields_ = [('l_max', C.c_int32), ('n_r', C.c_int32), ('n_theta', C.c_int32), ('n_phi', C.c_int32),
                ('max_batch', C.c_int32), ('hankel_skip', C.c_int32),
                ('cos_theta', C.POINTER(C.c_double)), ('gauss_w', C.POINTER(C.c_double)),
                ('legendre', C.POINTER(C.c_double)), ('legendre_len', C.c_int64),
                ('hankel_w', C.POINTER(C.c_double)), ('hankel_n_sum', C.c_int32),
                ('hankel_fwd_scale', C.c_double), ('hankel_inv_scale', C.c_double),
                ('int_weight', C.POINTER(C.c_double)), ('r_points', C.POINTER(C.c_double)),
                ('q_points', C.POINTER(C.c_double)), ('dimensions', C.c_int32)]


class ProjectionDesc(C.Structure):
    _fields_ = [('n_orders', C.c_int32), ('n_cols', C.POINTER(C.c_int32)),
                ('v', C.POINTER(C.POINTER(C.c_double))), ('radial_mask', C.POINTER(C.c_uint8)),
                ('sqrt_n_particles', C.c_double), ('sv_cutoff', C.c_double), ('max_sweeps', C.c_int32)]


class RealDesc(C.Structure):
    _fields_ = [('n_ops', C.c_int32), ('ops', C.c_int32 * 4), ('hio_considered', C.c_int32 * 4),
                ('use_lo', C.c_int32), ('use_hi', C.c_int32), ('lo', C.c_double), ('hi', C.c_double),
                ('imag_limit', C.c_double), ('error_inside_initial_support', C.c_int32), ('average_center_shells', C.c_int32)]


EXPORTS = {
    'xfb_last_error': (C.c_char_p, []),
    'xfb_device_count': (C.c_int, [C.POINTER(C.c_int)]),
    'xfb_set_device': (C.c_int, [C.c_int]),
    'xfb_plan_create': (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(PlanDesc)]),
    'xfb_plan_destroy': (C.c_int, [C.c_void_p]),
    'xfb_plan_set_projection': (C.c_int, [C.c_void_p, C.POINTER(ProjectionDesc)]),
    'xfb_plan_set_projection_2d': (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_double, C.c_int32]),
    'xfb_get_unknowns_2d': (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    'xfb_plan_set_real': (C.c_int, [C.c_void_p, C.POINTER(RealDesc), C.c_void_p]),
    'xfb_plan_workspace_bytes': (C.c_int64, [C.c_void_p]),
    'xfb_debug_jacobi_phase_cycles': (C.c_int, [C.POINTER(C.c_double)]),
    'xfb_debug_jacobi_sweeps': (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    'xfb_plan_set_fused_ft_stab': (C.c_int, [C.c_void_p, C.c_int32]),
    'xfb_plan_set_dual_stream': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    'xfb_plan_set_sht_chunk': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32]),
    'xfb_sht_forward': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    'xfb_sht_inverse': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    'xfb_hankel_apply': (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    'xfb_ft': (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    'xfb_project_invariants': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    'xfb_get_unknowns': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    'xfb_deg2_invariants': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    'xfb_plan_set_deg2_reference': (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    'xfb_deg2_invariant_diff': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    'xfb_mtip_enable_deg2_metric': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32]),
    'xfb_mtip_get_deg2_errors': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    'xfb_modify_intensity': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    'xfb_real_update': (C.c_int, [C.c_void_p, C.c_int32, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    'xfb_shrinkwrap': (C.c_int, [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_int32, C.c_void_p]),
    'xfb_mtip_init': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    'xfb_mtip_iterate': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_double), C.c_void_p]),
    'xfb_mtip_shrinkwrap': (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_void_p]),
    'xfb_mtip_step_host': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'xfb_mtip_get_nonfinite': (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.c_void_p]),
    'xfb_plan_set_host_chunk': (C.c_int, [C.c_void_p, C.c_int32]),
    'xfb_mtip_set_outer_iteration': (C.c_int, [C.c_void_p, C.c_int32]),
    'xfb_mtip_select_best': (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    'xfb_mtip_snapshot_intensity': (C.c_int, [C.c_void_p, C.c_void_p]),
    'xfb_mtip_fix_intensity': (C.c_int, [C.c_void_p, C.c_void_p]),
    'xfb_mtip_set_non_fxs': (C.c_int, [C.c_void_p, C.c_int32]),
    'xfb_mtip_shrinkwrap_center': (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_void_p]),
    'xfb_mtip_get_grid': (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    'xfb_mtip_get_errors': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.POINTER(C.c_int32), C.c_void_p]),
    'xfb_apply_matrix': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    'xfb_plan_launch_count': (C.c_int64, [C.c_void_p]),
    'xfb_plan_graph_replays': (C.c_int64, [C.c_void_p]),
    'xfb_profile_enable': (C.c_int, [C.c_void_p, C.c_int32]),
    'xfb_profile_read': (C.c_int, [C.c_void_p, C.c_int32, C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_int64),
                                   C.POINTER(C.c_int32)]),
}

_lib = None


def load():
    """Load libxfb200.so (built in-tree by __graft_entry__.build() / csrc/Makefile)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise XfbError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(xframe_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)      # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise XfbError(load().xfb_last_error().decode())
