"""ctypes binding of include/flexlmm_cuda.h — the same C ABI the R .Call shim binds (flexlmm_b200/r/r_shim.c).

There is no CPU fallback: if the shared library is missing it is built with nvcc; if no B200 is visible
flx_create fails with FLX_ERR_NO_DEVICE and FlxError is raised.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

FLX_MAX_K = 24
FLX_MAX_S = 8
ERR_NAMES = {0: "FLX_OK", -1: "FLX_ERR_BAD_ARG", -2: "FLX_ERR_MISSING_GENOTYPE", -3: "FLX_ERR_CUDA", -4: "FLX_ERR_NCCL",
             -5: "FLX_ERR_OOM", -6: "FLX_ERR_PGEN", -7: "FLX_ERR_STATE", -8: "FLX_ERR_NO_DEVICE"}

EXPORTS = ["flx_last_error", "flx_version", "flx_device_count", "flx_create", "flx_destroy", "flx_set_whitening", "flx_whitening_begin",
           "flx_whitening_cols", "flx_whitening_end", "flx_decorrelate",
           "flx_set_null", "flx_set_design", "flx_set_design_probes", "flx_set_use_dosage", "flx_set_perm", "flx_set_perm_implicit", "flx_fit_null_model", "flx_open_pgen", "flx_pgen_info", "flx_set_packed", "flx_make_resident",
           "flx_run", "flx_prepare", "flx_get_stats", "flx_perm_indices", "flx_decode_tile", "flx_whiten", "flx_comm_unique_id",
           "flx_comm_attach", "flx_minp_threshold", "flx_set_whitening_bcast", "flx_set_perm_bcast", "flx_multi_create", "flx_multi_destroy",
           "flx_multi_device_count", "flx_multi_set_whitening", "flx_multi_set_null", "flx_multi_set_design", "flx_multi_set_design_probes", "flx_multi_set_use_dosage", "flx_multi_set_perm",
           "flx_multi_set_perm_implicit", "flx_multi_open_pgen", "flx_multi_set_packed", "flx_multi_make_resident", "flx_multi_prepare",
           "flx_multi_run", "flx_multi_get_stats"]


class FlxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


class FlxConfig(C.Structure):
    _fields_ = [("device", C.c_int32), ("batch_tiles", C.c_int32), ("verbose", C.c_int32), ("flags", C.c_int32)]


class FlxSnpOut(C.Structure):
    _fields_ = [("lrt_chisq", C.c_void_p), ("pval", C.c_void_p), ("lrt_df", C.c_void_p), ("rank", C.c_void_p),
                ("pivot", C.c_void_p), ("coef", C.c_void_p)]


class FlxStats(C.Structure):
    _fields_ = [("ms_total", C.c_double), ("ms_decode", C.c_double), ("ms_whiten", C.c_double), ("ms_fit", C.c_double),
                ("ms_perm", C.c_double), ("ms_h2d", C.c_double), ("ms_d2h", C.c_double), ("launches", C.c_int64),
                ("launches_whiten", C.c_int64), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64), ("snps", C.c_int64),
                ("n_pad", C.c_int64), ("cols_per_tile", C.c_int64), ("int8_path", C.c_int64), ("refit_snps", C.c_int64),
                ("int8_passes", C.c_int64), ("fallback_reason", C.c_int64), ("ms_wall", C.c_double), ("ms_batches", C.c_double)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


_LIB = None


def lib_path():
    return _build.LIB


def load(build_if_missing=True):
    """dlopen libflexlmm_cuda.so (building it in-tree if needed) and declare every prototype of the header."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if build_if_missing:
        _build.build()
    if not os.path.exists(_build.LIB):
        raise FlxError(-8, f"{_build.LIB} is missing and could not be built; there is no CPU fallback")
    L = C.CDLL(_build.LIB, mode=C.RTLD_GLOBAL)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    L.flx_last_error.restype = C.c_char_p
    L.flx_version.restype = C.c_char_p
    L.flx_device_count.restype = C.c_int
    L.flx_create.argtypes = [C.POINTER(FlxConfig), C.POINTER(vp)]
    L.flx_destroy.argtypes = [vp]
    L.flx_destroy.restype = None
    L.flx_set_whitening.argtypes = [vp, vp, i64, i64, C.c_int]
    L.flx_whitening_begin.argtypes = [vp, i64]
    L.flx_whitening_cols.argtypes = [vp, i64, i64, vp, i64]
    L.flx_whitening_end.argtypes = [vp, C.c_int]
    L.flx_decorrelate.argtypes = [vp, vp, i64, dbl, dbl, vp, vp, i32, dbl, dbl, vp, vp, vp, C.POINTER(i32)]
    L.flx_set_null.argtypes = [vp, vp, i32, vp, dbl, i32]
    L.flx_set_design.argtypes = [vp, vp, vp, vp, i32]
    L.flx_set_design_probes.argtypes = [vp, vp, vp]
    L.flx_set_use_dosage.argtypes = [vp, C.c_int]
    L.flx_set_perm.argtypes = [vp, vp, vp, vp, i64, vp, i32]
    L.flx_set_perm_implicit.argtypes = [vp, vp, i32]
    L.flx_fit_null_model.argtypes = [vp, i32, vp, i64, C.POINTER(dbl), C.POINTER(i32), vp, vp, vp, vp, C.POINTER(i64)]
    L.flx_open_pgen.argtypes = [vp, C.c_char_p, vp, i64]
    L.flx_pgen_info.argtypes = [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i32)]
    L.flx_set_packed.argtypes = [vp, vp, i64, i64, i64, vp, i64]
    L.flx_make_resident.argtypes = [vp]
    L.flx_run.argtypes = [vp, vp, i64, C.POINTER(FlxSnpOut), vp, C.POINTER(i64)]
    L.flx_prepare.argtypes = [vp]
    L.flx_get_stats.argtypes = [vp, C.POINTER(FlxStats)]
    L.flx_perm_indices.argtypes = [i32, i64, vp]
    L.flx_decode_tile.argtypes = [vp, vp, i64, vp, vp]
    L.flx_whiten.argtypes = [vp, vp, i64, vp]
    L.flx_comm_unique_id.argtypes = [vp]
    L.flx_comm_attach.argtypes = [vp, vp, i32, i32]
    L.flx_set_whitening_bcast.argtypes = [vp, vp, i64, i64, C.c_int, i32]
    L.flx_set_perm_bcast.argtypes = [vp, vp, vp, vp, i64, vp, i32, i32]
    L.flx_multi_create.argtypes = [C.POINTER(FlxConfig), vp, i32, C.POINTER(vp)]
    L.flx_multi_destroy.argtypes = [vp]
    L.flx_multi_destroy.restype = None
    L.flx_multi_device_count.argtypes = [vp]
    L.flx_multi_set_whitening.argtypes = [vp, vp, i64, i64, C.c_int]
    L.flx_multi_set_null.argtypes = [vp, vp, i32, vp, dbl, i32]
    L.flx_multi_set_design.argtypes = [vp, vp, vp, vp, i32]
    L.flx_multi_set_design_probes.argtypes = [vp, vp, vp]
    L.flx_multi_set_use_dosage.argtypes = [vp, C.c_int]
    L.flx_multi_set_perm.argtypes = [vp, vp, vp, vp, i64, vp, i32]
    L.flx_multi_set_perm_implicit.argtypes = [vp, vp, i32]
    L.flx_multi_open_pgen.argtypes = [vp, C.c_char_p, vp, i64]
    L.flx_multi_set_packed.argtypes = [vp, vp, i64, i64, i64, vp, i64]
    L.flx_multi_make_resident.argtypes = [vp]
    L.flx_multi_prepare.argtypes = [vp]
    L.flx_multi_run.argtypes = [vp, vp, i64, C.POINTER(FlxSnpOut), vp, C.POINTER(i64)]
    L.flx_multi_get_stats.argtypes = [vp, i32, C.POINTER(FlxStats)]
    L.flx_minp_threshold.argtypes = [vp, i64, dbl]
    L.flx_minp_threshold.restype = dbl
    _LIB = L
    return L


def check(rc):
    if rc != 0:
        raise FlxError(rc, load().flx_last_error().decode())


def f64(a, fortran=False):
    a = np.asarray(a, dtype=np.float64)
    return np.asfortranarray(a) if fortran else np.ascontiguousarray(a)


def ptr(a):
    return None if a is None else a.ctypes.data
