"""ctypes binding of libalpha_eigen_b200.so (the C ABI declared in include/alpha_eigen_b200.h).

There is no CPU fallback: if the shared library is missing the import of any solver fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# AE_LIB selects another build of the same sources (kernel A/B experiments, tools/); the product is the in-tree default
LIB_PATH = os.environ.get("AE_LIB") or os.path.join(_HERE, "libalpha_eigen_b200.so")

AE_OK, AE_NOT_CONVERGED, AE_BAD_ARG, AE_CUDA_ERROR = 0, 1, 2, 3
AE_TILE, AE_LANE_CELLS = 256, 8
AE_WORK_NO_FUSED, AE_WORK_REREAD_PSI = 1, 2

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)


class AeCtrl(C.Structure):
    _fields_ = [("change", C.c_double), ("outer_change", C.c_double), ("k_new", C.c_double), ("k_old", C.c_double),
                ("k_change", C.c_double), ("inner_done", C.c_int32), ("inner_iters", C.c_int32),
                ("power_done", C.c_int32), ("power_iters", C.c_int32), ("ticket", C.c_uint32), ("reserved", C.c_int32)]


class AeSlab(C.Structure):
    _fields_ = [("Z", C.c_int64), ("Zp", C.c_int64), ("A", C.c_int32), ("n_neg", C.c_int32), ("G", C.c_int32),
                ("M", C.c_int32), ("mu_ihx", C.c_void_p), ("wgt", C.c_void_p), ("mat", C.c_void_p),
                ("scat", C.c_void_p), ("chi", C.c_void_p), ("nusf", C.c_void_p), ("hs", C.c_void_p),
                ("cdt", C.c_void_p), ("s_ext", C.c_void_p), ("tile_info", C.c_void_p), ("inflow", C.c_void_p), ("g0", C.c_int32), ("gl", C.c_int32)]


class AeWork(C.Structure):
    _fields_ = [("phi", C.c_void_p), ("phi_alt", C.c_void_p), ("phi_outer", C.c_void_p), ("base", C.c_void_p),
                ("q0", C.c_void_p), ("q1", C.c_void_p), ("partial", C.c_void_p), ("reduced", C.c_void_p),
                ("phi_fixed", C.c_void_p), ("norm_scratch", C.c_void_p), ("ctrl", C.c_void_p),
                ("ctrl_host", C.c_void_p), ("sweep_sync", C.c_void_p), ("comm", C.c_void_p), ("P", C.c_int32), ("batch", C.c_int32),
                ("flags", C.c_int32), ("reserved", C.c_int32)]


class AeMember(C.Structure):
    _fields_ = [("slab", C.POINTER(AeSlab)), ("work", C.POINTER(AeWork)), ("psi", C.c_void_p), ("psi_snap", C.c_void_p),
                ("phi_snap", C.c_void_p), ("hist", C.c_void_p)]


# every symbol include/alpha_eigen_b200.h declares: (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "ae_abi_version": (C.c_int, []),
    "ae_last_error": (C.c_char_p, []),
    "ae_padded_cells": (C.c_int64, [C.c_int64]),
    "ae_padded_cells_sharded": (C.c_int64, [C.c_int64, C.c_int32]),
    "ae_cell_offset": (C.c_int64, [C.c_int64]),
    "ae_partial_count": (C.c_int32, [C.c_int32, C.c_int32]),
    "ae_norm_blocks": (C.c_int64, [C.c_int64]),
    "ae_sweep_set": (C.c_int, [C.POINTER(AeSlab), _P, _P, _P, _P, _P, _P]),
    "ae_sweep_set_phi": (C.c_int, [C.POINTER(AeSlab), _P, _P, _P, c_ip, _P]),
    "ae_tile_info": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.c_int32, _P, _P]),
    "ae_sweep_sync_bytes": (C.c_int64, []),
    "ae_sweep_plan": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_int32, c_ip, c_ip, c_ip, c_ip,
                                  C.c_int64]),
    "ae_reduce_partials": (C.c_int, [C.POINTER(AeSlab), _P, C.c_int32, _P, _P]),
    "ae_flux_update": (C.c_int, [C.POINTER(AeSlab), C.POINTER(AeWork), _P, C.c_int32, _P, C.c_double, C.c_int32, _P]),
    "ae_sweep_and_update": (C.c_int, [C.POINTER(AeSlab), C.POINTER(AeWork), _P, _P, _P, _P, C.c_double, C.c_int32, _P]),
    "ae_exchange_partials": (C.c_int, [C.POINTER(AeSlab), C.POINTER(AeWork), _P]),
    "ae_pack_cells": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int32, _P, _P]),
    "ae_unpack_cells": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int32, _P, _P]),
    "ae_pack_cells_range": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int64, C.c_int32, _P, _P]),
    "ae_unpack_cells_range": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int64, C.c_int32, _P, _P]),
    "ae_pack_psi": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int32, C.c_int32, _P, C.c_int32, _P, _P]),
    "ae_unpack_psi": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int32, C.c_int32, _P, C.c_int32, _P, _P]),
    "ae_source_iteration": (C.c_int, [C.POINTER(AeSlab), C.POINTER(AeWork), _P, C.c_double, C.c_double, C.c_int32,
                                      c_ip, c_ip, _P]),
    "ae_power_iteration": (C.c_int, [C.POINTER(AeSlab), C.POINTER(AeWork), C.c_double, C.c_int32, C.c_double,
                                     C.c_int32, _P, c_dp, c_dp, c_ip, c_dp, c_ip, c_ip, _P]),
    "ae_time_step": (C.c_int, [C.POINTER(AeSlab), C.POINTER(AeWork), _P, _P, C.c_double, c_ip, c_ip, _P]),
    "ae_time_march": (C.c_int, [C.POINTER(AeSlab), C.POINTER(AeWork), _P, _P, _P, C.c_int32, C.c_double, c_ip, c_ip,
                                _P]),
    "ae_ensemble_time_march": (C.c_int, [C.POINTER(AeMember), C.c_int32, C.c_int32, C.c_double, c_ip, c_ip, _P]),
    "ae_gram": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int32, _P, _P]),
    "ae_dmd_from_gram": (C.c_int, [_P, C.c_int32, C.c_double, C.c_double, c_dp, c_dp, c_dp, c_ip, _P]),
    "ae_tsqr": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int32, _P, _P]),
    "ae_tsqr_merge": (C.c_int, [_P, C.c_int32, C.c_int32, _P, _P]),
    "ae_dmd_from_r": (C.c_int, [_P, C.c_int32, C.c_double, C.c_double, c_dp, c_dp, c_dp, c_ip, _P]),
    "ae_dmd_last_jacobi_sweeps": (C.c_int, []),
    "ae_dmd_from_r_batched": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_double, C.c_double, _P, _P, _P, _P, _P, _P]),
    "ae_comm_unique_id": (C.c_int, [_P]),
    "ae_comm_init": (C.c_int, [C.POINTER(_P), C.c_int32, C.c_int32, _P]),
    "ae_comm_allreduce_sum": (C.c_int, [_P, _P, C.c_int64, _P]),
    "ae_comm_gather_rows": (C.c_int, [_P, _P, C.c_int32, C.c_int64, _P]),
    "ae_comm_gather_groups": (C.c_int, [_P, _P, C.c_int32, C.c_int64, _P]),
    "ae_comm_exchange_mode": (C.c_int32, [_P]),
    "ae_group_range": (None, [C.c_int32, C.c_int32, C.c_int32, c_ip, c_ip]),
    "ae_comm_destroy": (C.c_int, [_P]),
}

_lib = None


class AlphaEigenError(RuntimeError):
    pass


def load():
    """Load the in-tree shared library; raise if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AlphaEigenError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.ae_abi_version() != 2:
            raise AlphaEigenError("ABI version mismatch")
        _lib = lib
    return _lib


def last_error() -> str:
    return load().ae_last_error().decode()


def check(status: int, convergence_message: str = "did not converge"):
    """Map ae_status to the reference's exceptions: non-convergence is an AssertionError there
    (group.py:72,102; time_unit.py:35,51)."""
    if status == AE_OK:
        return
    if status == AE_NOT_CONVERGED:
        raise AssertionError(convergence_message)
    if status == AE_BAD_ARG:
        raise ValueError(last_error())
    raise AlphaEigenError(last_error())
