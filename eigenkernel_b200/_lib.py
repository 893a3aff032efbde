"""ctypes binding of libekb200.so (the C-ABI of include/ekb200.h).

The product path FAILS LOUDLY when the CUDA library is missing or no GPU is visible: there is no CPU
fallback (BASELINE.json north_star).  `load()` only dlopens; creating a context needs a GPU.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, byref, c_char, c_char_p, c_double, c_int, c_int32, c_int64, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# EKB200_LIB: an alternative build of the same C-ABI (experiments with build-time variants; tests and bench use the default)
LIB_PATH = os.environ.get("EKB200_LIB") or os.path.join(_HERE, "libekb200.so")

_dp = POINTER(c_double)


class Ekb200Error(RuntimeError):
    def __init__(self, fn: str, info: int, text: str = ""):
        self.fn, self.info = fn, info
        super().__init__(f"{fn}: info = {info} ({text})")


_lib = None

# name -> (argtypes)  ; every function returns int info unless listed in _RESTYPE
_SIGS = {
    "ekb200_version": [],
    "ekb200_device_count": [],
    "ekb200_create": [POINTER(c_void_p), c_int],
    "ekb200_destroy": [c_void_p],
    "ekb200_strerror": [c_int],
    "ekb200_last_error": [c_void_p],
    "ekb200_set_option": [c_void_p, c_char_p, c_int64],
    "ekb200_num_events": [c_void_p],
    "ekb200_get_event": [c_void_p, c_int, POINTER(c_char_p), POINTER(c_double), POINTER(c_int)],
    "ekb200_clear_events": [c_void_p],
    "ekb200_dev_alloc": [c_void_p, c_int64, POINTER(c_void_p)],
    "ekb200_dev_free": [c_void_p, c_void_p],
    "ekb200_h2d": [c_void_p, c_void_p, c_void_p, c_int64],
    "ekb200_d2h": [c_void_p, c_void_p, c_void_p, c_int64],
    "ekb200_h2d_matrix": [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64],
    "ekb200_d2h_matrix": [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64],
    "ekb200_sync": [c_void_p],
    "ekb200_coo_to_dense": [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_int64],
    "ekb200_fill_synthetic": [c_void_p, c_int64, c_uint64, c_double, c_int, c_double, c_void_p, c_int64],
    "ekb200_dgemm": [c_void_p, c_char, c_char, c_int64, c_int64, c_int64, c_double, c_void_p, c_int64, c_void_p,
                     c_int64, c_double, c_void_p, c_int64],
    "ekb200_potrf": [c_void_p, c_int64, c_void_p, c_int64],
    "ekb200_sygst": [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64],
    "ekb200_trtrs_lt": [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_int64],
    "ekb200_sy2sb": [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p],
    "ekb200_sy2sb_num_panels": [c_void_p, c_int64],
    "ekb200_get_band": [c_void_p],
    "ekb200_stedc": [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, POINTER(c_double)],
    "ekb200_stebz_stein": [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int64],
    "ekb200_sb2st": [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p],
    "ekb200_sb2st_max_tasks": [c_void_p, c_int64],
    "ekb200_apply_q2": [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64],
    "ekb200_apply_q1": [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64],
    "ekb200_syevd_dev": [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64],
    "ekb200_sygvd_dev": [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p,
                         c_int64],
    "ekb200_syevd": [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64],
    "ekb200_sygvd": [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64],
    "ekb200_sygvd_coo": [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_void_p,
                         c_void_p, c_void_p, c_int64],
    "ekb200_last_merge_flops": [c_void_p],
    "ekb200_host_alloc": [c_void_p, c_int64, POINTER(c_void_p)],
    "ekb200_host_free": [c_void_p, c_void_p],
    "ekb200_num_launches": [c_void_p],
    "ekb200_timer_start": [c_void_p],
    "ekb200_timer_stop": [c_void_p, POINTER(c_double)],
    "ekb200_gemm_profile": [c_void_p, POINTER(c_double), POINTER(c_double), POINTER(c_int64)],
    "ekb200_measure_fp64_peak": [c_void_p, POINTER(c_double), POINTER(c_double)],
    "ekb200_kernel_profile": [c_void_p, POINTER(c_double), POINTER(c_double), POINTER(c_int64)],
    "ekb200_profile_rows": [c_void_p],
    "ekb200_profile_row": [c_void_p, c_int, POINTER(c_char_p), POINTER(c_int), POINTER(c_double), POINTER(c_double),
                           POINTER(c_int64)],
    "ekb200_eval_residual_norm": [c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_int64, POINTER(c_double), POINTER(c_double),
                                  POINTER(c_double)],
    "ekb200_eval_orthogonality": [c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                                  c_int64, POINTER(c_double)],
    "ekb200_get_ipratios": [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_void_p],
    "ekb200_eval_residual_norm_dev": [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p,
                                      c_void_p, c_int64, POINTER(c_double), POINTER(c_double), POINTER(c_double)],
    "ekb200_eval_orthogonality_dev": [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_int64,
                                      POINTER(c_double)],
    "ekb200_eval_b_orthonormality_dev": [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_int64,
                                         POINTER(c_double), POINTER(c_double)],
    "ekb200_get_ipratios_dev": [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p],
    "ekb200_comm_unique_id": [c_void_p],
    "ekb200_comm_init": [c_void_p, c_int, c_int, c_void_p],
    "ekb200_comm_info": [c_void_p, POINTER(c_int), POINTER(c_int)],
    "ekb200_comm_slab": [c_void_p, c_int64, POINTER(c_int64), POINTER(c_int64)],
    "ekb200_comm_local_cols": [c_void_p, c_int64, POINTER(c_int64)],
    "ekb200_comm_allgather_slabs": [c_void_p, c_int64, c_int64, c_void_p, c_int64],
    "ekb200_comm_bcast": [c_void_p, c_void_p, c_int64, c_int],
    "ekb200_num_collectives": [c_void_p],
}
_RESTYPE = {"ekb200_strerror": c_char_p, "ekb200_last_error": c_char_p, "ekb200_last_merge_flops": c_double,
            "ekb200_num_launches": c_int64, "ekb200_num_collectives": c_int64}


def exported_symbols():
    """Names every build of libekb200.so must export (= what include/ekb200.h declares)."""
    return sorted(_SIGS)


def load():
    """dlopen libekb200.so; raises ImportError (loudly) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C eigenkernel_b200/csrc). There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in _SIGS.items():
        f = getattr(lib, name)
        f.argtypes = argtypes
        f.restype = _RESTYPE.get(name, c_int)
    _lib = lib
    return lib
