"""ctypes binding of ``libapple_b200.so`` (the C ABI declared in ``include/apple_b200.h``).

There is no CPU fallback: if the library is missing, or a CUDA call fails, an exception is raised.
"""

from __future__ import annotations

import ctypes
from ctypes import POINTER, c_char_p, c_double, c_int, c_int32, c_int64, c_uint8, c_void_p
from pathlib import Path

import numpy as np
import torch

import os as _os

# APL_LIB: a tuning-experiment build of the same sources (apple_b200/build.py); the product library otherwise
LIB_PATH = Path(_os.environ.get("APL_LIB") or Path(__file__).resolve().parent / "libapple_b200.so")

# constants of include/apple_b200.h
OK = 0
F32, F64 = 0, 1
KIND_SNH, KIND_ARAP, KIND_SNH_MUSCLE, KIND_SNH_ARAP = 0, 1, 2, 3
OP_FUN, OP_GRAD, OP_HESS_DIAG, OP_HESS_PROD, OP_HESS_QUAD = 1, 2, 4, 8, 16
OP_HESS_OFFD, OP_PSD = 32, 64      # opt-in supersets: vertex-block off-diagonals, eigenvalue-clamped Hessian
SCATTER_TILE, SCATTER_ATOMIC, SCATTER_TILE_SIMPLE = 0, 1, 2
PNCG_NSCAL = 96
PCG_NSCAL = 16
S_F, S_F_PREV, S_GP, S_PHP, S_ALPHA, S_BETA, S_GNORM2, S_GNORM2_FIRST = 0, 1, 2, 3, 4, 5, 6, 7
S_ACCEPTED, S_LS_STEPS, S_K, S_N_ACCEPTED, S_DIAG_MEAN, S_GPG, S_DONE, S_FAILS, S_F_NEW, S_J = 8, 9, 10, 11, 12, 13, 15, 16, 17, 18
S_SUMS, S_ALPHA_J, S_ACC_J, S_FT_J = 20, 32, 48, 64
(PHASE_INIT, PHASE_REDUCE, PHASE_FINALIZE, PHASE_DIRECTION, PHASE_PASS_B, PHASE_ALPHA, PHASE_TRIAL, PHASE_LS,
 PHASE_COMMIT) = range(9)

# every symbol include/apple_b200.h declares: name -> (restype, argtypes)
PART_ALL, PART_BOUNDARY, PART_INTERIOR = 0, 1, 2

SIGNATURES = {
    "apl_version": (c_int, []),
    "apl_last_error": (c_char_p, []),
    "apl_device_count": (c_int, []),
    "apl_fem_create": (c_int, [c_int, c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_int, POINTER(c_void_p)]),
    "apl_fem_create_snh_arap": (c_int, [c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_void_p, c_int, POINTER(c_void_p)]),
    "apl_fem_destroy": (None, [c_void_p]),
    "apl_fem_create_from_mesh": (c_int, [c_int, c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_void_p, c_int, c_int, POINTER(c_void_p)]),
    "apl_fem_info": (c_int, [c_void_p, POINTER(c_int64)]),
    "apl_fem_host_tables": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "apl_fem_host_planes": (c_int, [c_void_p, c_void_p, POINTER(c_int64), POINTER(c_int64)]),
    "apl_fem_set_materials": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "apl_fem_eval": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                             c_void_p, c_int, c_int, c_void_p]),
    "apl_fem_eval_part": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "apl_fem_mark_boundary": (c_int, [c_void_p, c_void_p, POINTER(c_int64)]),
    "apl_fem_mixed_derivative_prod": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "apl_ext_force_eval": (c_int, [c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                   c_int, c_void_p]),
    "apl_field_copy": (c_int, [c_int, c_int64, c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "apl_halo_pack": (c_int, [c_int, c_int64, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "apl_halo_unpack": (c_int, [c_int, c_int64, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                c_int, c_void_p, c_void_p]),
    "apl_xchg_create": (c_int, [c_int, c_int, c_int, c_int64, POINTER(c_void_p)]),
    "apl_xchg_destroy": (None, [c_void_p]),
    "apl_xchg_ipc_handle": (c_int, [c_void_p, c_void_p]),
    "apl_xchg_connect": (c_int, [c_void_p, c_void_p]),
    "apl_xchg_set_plan": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "apl_xchg_push": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "apl_xchg_pull": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "apl_pncg_create": (c_int, [c_int, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p, POINTER(c_void_p)]),
    "apl_pncg_destroy": (None, [c_void_p]),
    "apl_pncg_add_fem": (c_int, [c_void_p, c_void_p]),
    "apl_pncg_add_ext_force": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "apl_pncg_set_params": (c_int, [c_void_p, c_double, c_double, c_double, c_double, c_double, c_double, c_double,
                                    c_int, c_int, c_int]),
    "apl_pncg_set_block_jacobi": (c_int, [c_void_p, c_void_p, c_void_p, c_int]),
    "apl_pncg_set_exchange": (c_int, [c_void_p, c_void_p]),
    "apl_pncg_current": (c_int, [c_void_p]),
    "apl_pncg_flip": (c_int, [c_void_p]),
    "apl_pncg_phase": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "apl_pncg_iterate": (c_int, [c_void_p, c_int, c_void_p]),
    "apl_pcg_create": (c_int, [c_int, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_void_p, POINTER(c_void_p)]),
    "apl_pcg_destroy": (None, [c_void_p]),
    "apl_pcg_add_fem": (c_int, [c_void_p, c_void_p]),
    "apl_pcg_set_params": (c_int, [c_void_p, c_double, c_double, c_int64, c_int, c_int, c_int]),
    "apl_pcg_init": (c_int, [c_void_p, c_int, c_void_p]),
    "apl_pcg_iterate": (c_int, [c_void_p, c_int, c_void_p]),
}

_lib = None


class NativeError(RuntimeError):
    """A call into libapple_b200.so returned an error code."""


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise NativeError(
                f"{LIB_PATH} is missing: build it with `python -m apple_b200.build` "
                "(apple_b200 has no CPU or PyTorch fallback)"
            )
        handle = ctypes.CDLL(str(LIB_PATH))
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc != OK:
        msg = lib().apl_last_error()
        raise NativeError(f"apple_b200 native call failed ({rc}): {msg.decode() if msg else '?'}")


def dtype_code(dtype: torch.dtype) -> int:
    if dtype == torch.float32:
        return F32
    if dtype == torch.float64:
        return F64
    raise TypeError(f"apple_b200 supports float32 and float64, got {dtype}")


def np_dtype(dtype: torch.dtype):
    return np.float32 if dtype == torch.float32 else np.float64


def host_ptr(a: np.ndarray | None) -> c_void_p | None:
    return None if a is None else a.ctypes.data_as(c_void_p)


def dev_ptr(t: torch.Tensor | None) -> c_void_p | None:
    """Raw device pointer of a CUDA tensor (contiguity is the caller's responsibility)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise NativeError("apple_b200 operators take CUDA tensors only (there is no CPU path)")
    return c_void_p(t.data_ptr())


def stream_ptr(device: torch.device | int | None = None) -> c_void_p:
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def field_ld(t: torch.Tensor, n_points: int, dtype: torch.dtype, name: str) -> int:
    """Validates a nodal field (n_points, 3|4) and returns its leading dimension."""
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if t.dim() != 2 or t.shape[0] != n_points or t.shape[1] not in (3, 4) or not t.is_contiguous():
        raise ValueError(f"{name}: expected a contiguous ({n_points}, 3|4) tensor, got {tuple(t.shape)}")
    return int(t.shape[1])
