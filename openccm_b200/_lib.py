"""
ctypes binding of liboccm_b200.so (C ABI declared in include/occm_b200.h).

The library is the product: if it cannot be loaded this module raises — there is no CPU or PyTorch fallback
for any solver entry point.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('OCCM_B200_LIB', os.path.join(HERE, 'csrc', 'liboccm_b200.so'))   # override: kernel experiments
ABI_VERSION = 2


class OccmError(RuntimeError):
    pass


class SystemDesc(C.Structure):
    _fields_ = [('model', C.c_int32), ('n_species', C.c_int32), ('n_models', C.c_int32), ('points_per_model', C.c_int32),
                ('row_ptr', C.c_void_p), ('row_src', C.c_void_p), ('row_w', C.c_void_p), ('a', C.c_void_p),
                ('pulse_ptr', C.c_void_p), ('pulse_fn', C.c_void_p), ('pulse_w', C.c_void_p),
                ('tables', C.c_void_p), ('tables_bytes', C.c_int32), ('fields', C.c_void_p), ('n_fields', C.c_int32),
                ('ell_width', C.c_int32), ('ell_src', C.c_void_p), ('ell_w', C.c_void_p), ('point_flags', C.c_void_p), ('ell_packed', C.c_void_p),
                ('flagged_points', C.c_void_p), ('n_flagged', C.c_int32)]


class Members(C.Structure):
    _fields_ = [('n_members', C.c_int32), ('k_scale', C.c_void_p), ('q_scale', C.c_void_p), ('bc_scale', C.c_void_p)]


class Rk45Problem(C.Structure):
    _fields_ = [('t0', C.c_double), ('t_bound', C.c_double), ('first_step', C.c_double), ('rtol', C.c_double),
                ('atol', C.c_double), ('y0', C.c_void_p), ('t_eval', C.c_void_p), ('n_t_eval', C.c_int32),
                ('save_idx', C.c_void_p), ('n_save', C.c_int32), ('out', C.c_void_p), ('out_t', C.c_void_p),
                ('max_steps', C.c_int64)]


class BdfProblem(C.Structure):
    _fields_ = [('t0', C.c_double), ('t_bound', C.c_double), ('first_step', C.c_double), ('rtol', C.c_double),
                ('atol', C.c_double), ('y0', C.c_void_p), ('t_eval', C.c_void_p), ('n_t_eval', C.c_int32),
                ('save_idx', C.c_void_p), ('n_save', C.c_int32), ('out', C.c_void_p), ('out_t', C.c_void_p),
                ('max_steps', C.c_int64), ('tdiag', C.c_void_p), ('gmres_restart', C.c_int32), ('lin_tol_factor', C.c_double),
                ('precond_mode', C.c_int32)]


# name -> (restype, argtypes); every symbol include/occm_b200.h declares
SIGNATURES = {
    'occm_abi_version': (C.c_int, []),
    'occm_last_error': (C.c_char_p, []),
    'occm_launch_count': (C.c_int64, []),
    'occm_system_create': (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(SystemDesc)]),
    'occm_system_destroy': (C.c_int, [C.c_void_p]),
    'occm_system_ndof': (C.c_int64, [C.c_void_p]),
    'occm_system_load_specialized': (C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t]),
    'occm_system_is_specialized': (C.c_int, [C.c_void_p]),
    'occm_rhs': (C.c_int, [C.c_void_p, C.POINTER(Members), C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]),
    'occm_rk45_workspace_bytes': (C.c_size_t, [C.c_void_p, C.c_int32]),
    'occm_rk45_init': (C.c_int, [C.POINTER(C.c_void_p), C.c_void_p, C.POINTER(Members), C.POINTER(Rk45Problem),
                                 C.c_void_p, C.c_size_t, C.c_void_p]),
    'occm_rk45_run': (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_int32), C.c_void_p]),
    'occm_rk45_resume': (C.c_int, [C.c_void_p, C.c_void_p]),
    'occm_rk45_stats': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p]),
    'occm_rk45_get_state': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    'occm_rk45_free': (C.c_int, [C.c_void_p]),
    'occm_rk45_time_stage': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_float), C.c_void_p]),
    'occm_bdf_workspace_bytes': (C.c_size_t, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]),
    'occm_bdf_precond_stats': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'occm_bdf_init': (C.c_int, [C.POINTER(C.c_void_p), C.c_void_p, C.POINTER(Members), C.POINTER(BdfProblem), C.c_void_p, C.c_size_t,
                                C.c_void_p]),
    'occm_bdf_run': (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_int32), C.c_void_p]),
    'occm_bdf_resume': (C.c_int, [C.c_void_p, C.c_void_p]),
    'occm_bdf_stats': (C.c_int, [C.c_void_p] * 11),
    'occm_bdf_get_state': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    'occm_bdf_free': (C.c_int, [C.c_void_p]),
    'occm_bdf_time_kernel': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_float), C.c_void_p]),
    'occm_dofs_to_elements': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p]),
    'occm_rtd_pulse': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]),
    'occm_rtd': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                           C.c_void_p, C.c_int32, C.c_int32, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]),
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load the shared library (once). Raises OccmError when it is missing or its ABI does not match."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OccmError(f'{LIB_PATH} not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                        f'(nvcc, sm_100a). There is no CPU fallback.')
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)           # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.occm_abi_version() != ABI_VERSION:
        raise OccmError(f'ABI mismatch: library {lib.occm_abi_version()} vs binding {ABI_VERSION}')
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise OccmError(f'liboccm_b200 error {rc}: {load().occm_last_error().decode()}')
