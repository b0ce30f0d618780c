"""ctypes binding of ``libfvs2d_gpu.so`` -- the C-ABI declared in ``include/fvs2d_gpu.h``.

There is no fallback: if the shared library is missing the import of :func:`lib` raises, and every
computing entry point fails without a CUDA device (``fvs2d_gpu_init`` reports it).
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from .config import Fvs2dConfig

_HERE = os.path.dirname(os.path.abspath(__file__))
# FVS2D_GPU_LIB selects an alternative build of the same library (tuning experiments: other tile sizes)
LIB_PATH = os.environ.get("FVS2D_GPU_LIB") or os.path.join(_HERE, "csrc", "libfvs2d_gpu.so")
_LIB = None

# every symbol include/fvs2d_gpu.h declares (tests check that the .so exports them all)
SYMBOLS = [
    "fvs2d_gpu_init", "fvs2d_gpu_comm_unique_id", "fvs2d_gpu_comm_init", "fvs2d_gpu_set_mesh", "fvs2d_gpu_set_lsq",
    "fvs2d_gpu_initialize_solution", "fvs2d_gpu_set_state", "fvs2d_gpu_get_state", "fvs2d_gpu_set_state_local",
    "fvs2d_gpu_get_state_local",
    "fvs2d_gpu_time_integration", "fvs2d_gpu_compute_residual", "fvs2d_gpu_get_aux", "fvs2d_gpu_test_resid",
    "fvs2d_gpu_interpolate_cell2node", "fvs2d_gpu_wall_values",
    "fvs2d_gpu_sizes", "fvs2d_gpu_scalars", "fvs2d_gpu_mesh_array", "fvs2d_host_build", "fvs2d_gpu_last_timing",
    "fvs2d_gpu_set_option", "fvs2d_gpu_last_error", "fvs2d_gpu_finalize",
]


class Fvs2dError(RuntimeError):
    """Non-zero return of a C-ABI call; the message is ``fvs2d_gpu_last_error()`` (the text the
    Fortran host would print before ``stop``)."""


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
    L = ctypes.CDLL(LIB_PATH)
    vp, ci, cd = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
    cfgp = ctypes.POINTER(Fvs2dConfig)
    L.fvs2d_gpu_init.argtypes = [cfgp, ci]
    L.fvs2d_gpu_comm_unique_id.argtypes = [vp]
    L.fvs2d_gpu_comm_init.argtypes = [ci, ci, vp]
    L.fvs2d_gpu_set_mesh.argtypes = [ci, ci, ci, vp, vp, vp, ci, vp, vp, vp]
    L.fvs2d_gpu_set_lsq.argtypes = [vp, vp, vp, vp]
    L.fvs2d_host_build.argtypes = [cfgp, ci, ci, ci, ci, ci, vp, vp, vp, ci, vp, vp, vp]
    L.fvs2d_gpu_initialize_solution.argtypes = []
    L.fvs2d_gpu_set_state.argtypes = [vp]
    L.fvs2d_gpu_get_state.argtypes = [vp]
    L.fvs2d_gpu_set_state_local.argtypes = [vp]
    L.fvs2d_gpu_get_state_local.argtypes = [vp]
    L.fvs2d_gpu_time_integration.argtypes = [cd, ci, vp, vp, vp]
    L.fvs2d_gpu_compute_residual.argtypes = [cd, vp, vp]
    L.fvs2d_gpu_get_aux.argtypes = [vp, vp, vp]
    L.fvs2d_gpu_test_resid.argtypes = [ci, vp, vp]
    L.fvs2d_gpu_interpolate_cell2node.argtypes = [vp, vp]
    L.fvs2d_gpu_wall_values.argtypes = [ci, vp]
    L.fvs2d_gpu_sizes.argtypes = [vp]
    L.fvs2d_gpu_scalars.argtypes = [vp]
    L.fvs2d_gpu_mesh_array.argtypes = [ctypes.c_char_p, vp]
    L.fvs2d_gpu_mesh_array.restype = ctypes.c_long
    L.fvs2d_gpu_last_timing.argtypes = [vp, ctypes.POINTER(ctypes.c_long)]
    L.fvs2d_gpu_set_option.argtypes = [ctypes.c_char_p, ci]
    L.fvs2d_gpu_last_error.restype = ctypes.c_char_p
    L.fvs2d_gpu_finalize.argtypes = []
    _LIB = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        raise Fvs2dError(lib().fvs2d_gpu_last_error().decode(errors="replace"))


def ptr(a):
    """Raw host pointer of a numpy array / torch CPU tensor / None."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(ctypes.c_void_p)
    return ctypes.c_void_p(a.data_ptr())  # torch tensor (pinned host memory in bench.py)


_INT_ARRAYS = {"en1", "en2", "ec1", "ec2", "cedge", "nghbre", "cell_intr", "b_edge", "b_edge_ptr", "grad_ptr", "grad_idx",
               "perm", "f_off", "f_nbr", "f_edge", "g_off", "g_idx", "orig_id", "loc2new", "bf_type", "bf_edge", "peers",
               "send_ptr", "send_idx", "recv_begin", "recv_count", "tile_es", "tile_ne", "tile_hc_ptr", "tile_he_ptr",
               "tile_hc_idx", "tile_he_idx", "f_bf", "tile_hdr", "t_bf", "fz_hdr", "fz_h2_idx", "fz_info", "fz_tile_int", "fz_tile_bnd", "gh_ptr", "gh_idx", "sub_orig", "sub_new_id"}
_U32_ARRAYS = {"f_pack", "t_pack", "fz_pack2", "fz_hf", "fz_uf"}
_U16_ARRAYS = {"fz_gslot"}
_BYTE_ARRAYS = {"is_intr"}


def mesh_array(name: str) -> np.ndarray:
    L = lib()
    n = L.fvs2d_gpu_mesh_array(name.encode(), None)
    if n < 0:
        raise Fvs2dError(L.fvs2d_gpu_last_error().decode())
    dt = np.int32 if name in _INT_ARRAYS else np.uint8 if name in _BYTE_ARRAYS else np.uint32 if name in _U32_ARRAYS else np.uint16 if name in _U16_ARRAYS else np.float64
    out = np.zeros(n, dtype=dt)
    if n:
        L.fvs2d_gpu_mesh_array(name.encode(), ptr(out))
    return out
