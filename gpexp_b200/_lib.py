"""ctypes binding of libgpexp_b200.so (the C ABI declared in include/gpexp_b200.h).

There is no CPU fallback: if the shared library has not been built, importing this module
raises ImportError, and creating a handle without a CUDA device raises GpxError.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GPX_LIB") or os.path.join(HERE, "lib", "libgpexp_b200.so")  # GPX_LIB: A/B builds only

GPX_MAX_DIM = 16
GPX_KROWS = 16
GPX_PIVOT_HDR = 3 + GPX_MAX_DIM
SE, MATERN32, MEHLER = 0, 1, 2
SIDE_A, SIDE_B = 0, 1
ROW_KERNEL, ROW_MATRIX = 0, 1


class GpxError(RuntimeError):
    """A C-ABI call returned non-zero (message from gpx_last_error)."""


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "or `make -C gpexp_b200/csrc`.  gpexp_b200 has no CPU fallback.")

lib = C.CDLL(LIB_PATH)

_p = C.c_void_p
_i64 = C.c_int64
_int = C.c_int
_dbl = C.c_double

# name -> argtypes (restype int unless listed in _RESTYPES); mirrors include/gpexp_b200.h one to one
SIGNATURES = {
    "gpx_version": [],
    "gpx_last_error": [],
    "gpx_create": [_int, C.POINTER(_p)],
    "gpx_destroy": [_p],
    "gpx_set_kernel": [_p, _int, _int, C.POINTER(_dbl), _int],
    "gpx_kernel_pairwise": [_p, _p, _i64, _i64, _p, _i64, _i64, _p, _p],
    "gpx_prior_diag": [_p, _p, _i64, _i64, _p, _p],
    "gpx_gram": [_p, _p, _i64, _i64, _p, _i64, _i64, _p, _i64, _int, _p, _dbl, _p],
    "gpx_potrf": [_p, _p, _i64, _i64, _p, _p],
    "gpx_chol_append": [_p, _p, _i64, _i64, _p, _dbl, _p, _p],
    "gpx_prep_side": [_p, _int, _p, _i64, _i64, _p, _p, _i64, _p],
    "gpx_trsm_gram": [_p, _p, _i64, _i64, _p, _p, _i64, _p, _p, _p, _i64, _i64, _p, _i64, _p, _p],
    "gpx_trsm": [_p, _p, _i64, _i64, _p, _i64, _i64, _p],
    "gpx_trsm_back": [_p, _p, _i64, _i64, _p, _i64, _i64, _p],
    "gpx_trtri_t": [_p, _p, _i64, _i64, _p, _i64, _p],
    "gpx_dgemm_tn_sub": [_p, _p, _i64, _p, _i64, _p, _i64, _i64, _i64, _i64, _int, _p],
    "gpx_dgemm_tn_sub_padded": [_p, _p, _i64, _p, _i64, _p, _i64, _i64, _i64, _i64, _int, _p],
    "gpx_gather_pivot": [_p, _p, _i64, _i64, _p, _p, _i64, _p, _p, _i64, _dbl, _p, _p],
    "gpx_select_pivot": [_p, _p, _int, _i64, _i64, _int, _p, _p],
    "gpx_append_row": [_p, _int, _p, _p, _p, _i64, _i64, _p, _i64, _i64, _p, _p],
    "gpx_argreduce": [_p, _p, _p, _p, _i64, _int, _p, _p, _p],
    "gpx_sum": [_p, _p, _i64, _p, _p],
    "gpx_score_ivar_workspace": [_p, _i64, _i64],
    "gpx_score_ivar": [_p, _p, _i64, _p, _p, _p, _i64, _p, _i64, _p, _p, _p, _i64, _i64, _dbl, _dbl, _p, _p, _p, _p,
                       _p, _p],
    "gpx_cov_segments": [_i64],
    "gpx_cov_update": [_p, _p, _i64, _i64, _i64, _p, _p, _p, _i64, _p],
    "gpx_cov_from_factors": [_p, _p, _i64, _p, _p, _i64, _p, _i64, _p, _p, _i64, _i64, _p, _i64, _p],
    "gpx_score_ivar_partials": [_p, _p, _int, _i64, _p, _i64, _p, _i64, _dbl, _dbl, _p, _p, _p, _p, _p],
    "gpx_score_mi": [_p, _p, _p, _dbl, _p, _i64, _p, _p, _p, _p],
    "gpx_mi_prec_column_workspace": [_i64, _i64],
    "gpx_mi_prec_column": [_p, _p, _i64, _i64, _i64, _i64, _p, _p, _p, _p, _p],
    "gpx_gather_column": [_p, _p, _i64, _i64, _p, _p, _p, _p],
    "gpx_local_index": [_p, _p, _i64, _i64, _p, _p],
    "gpx_colsumsq": [_p, _p, _i64, _i64, _i64, _p, _p, _p],
    "gpx_transpose": [_p, _p, _i64, _i64, _i64, _p, _i64, _p],
    "gpx_set_mask": [_p, _p, _p, C.c_uint8, _p],
    "gpx_store_pivot": [_p, _p, _i64, _p, _i64, _p, _p, _p],
    "gpx_se_dgram": [_p, _p, _i64, _i64, _p, _i64, _i64, _p, _i64, _p],
    "gpx_se_var_grad": [_p, _p, _i64, _i64, _p, _i64, _i64, _p, _p, _p, _p],
    "gpx_rowsum": [_p, _p, _i64, _i64, _i64, _dbl, _p, _p],
    "gpx_logdet_chol": [_p, _p, _i64, _i64, _p, _p],
    "gpx_bench_dmma": [_p, _i64, _p, _p],
    "gpx_bench_dfma": [_p, _i64, _p, _p],
}
_RESTYPES = {"gpx_last_error": C.c_char_p, "gpx_score_ivar_workspace": _i64, "gpx_mi_prec_column_workspace": _i64}

for _name, _args in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here == the .so does not export a declared symbol
    _fn.argtypes = _args
    _fn.restype = _RESTYPES.get(_name, _int)


def last_error() -> str:
    msg = lib.gpx_last_error()
    return msg.decode() if msg else ""


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise GpxError(f"{what or 'gpexp_b200'} failed with code {rc}: {last_error()}")
