"""ctypes wrapper over ``oracle/liboracle.so`` -- the CPU parity oracle (TEST INFRASTRUCTURE).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module.  Parity unpinned: see the header of ``fvs2d_oracle.c``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from fvs2d_b200.config import Fvs2dConfig
from fvs2d_b200.meshio import Mesh

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


def build(force: bool = False) -> None:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "fvs2d_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "all"])


def _lib(fast=False):
    """fast: False -> parity build; True -> -Ofast timing build; "omp" -> all-cores context variant (see Makefile)."""
    name = "liboracle_omp.so" if fast == "omp" else "liboracle_fast.so" if fast else "liboracle.so"
    if name in _LIBS:
        return _LIBS[name]
    path = os.path.join(_HERE, name)
    if not os.path.exists(path):
        build()
    L = ctypes.CDLL(path)
    dp, ip, vp = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int), ctypes.c_void_p
    L.orc_create.restype = vp
    L.orc_create.argtypes = [ctypes.c_int] * 3 + [vp, vp, vp, ctypes.c_int, vp, vp, vp]
    L.orc_last_error.restype = ctypes.c_char_p
    L.orc_last_error.argtypes = [vp]
    L.orc_setup.argtypes = [vp, ctypes.POINTER(Fvs2dConfig)]
    L.orc_initialize_solution.argtypes = [vp]
    L.orc_compute_residual.argtypes = [vp, ctypes.c_double]
    L.orc_time_integration.argtypes = [vp, ctypes.c_double, ctypes.c_int, vp, vp, vp]
    L.orc_test_resid.argtypes = [vp, ctypes.c_int, vp, vp]
    L.orc_vortex_error.argtypes = [vp, ctypes.c_double, vp, vp]
    L.orc_interpolate_cell2node.argtypes = [vp, ctypes.c_int, vp]
    L.orc_wall_values.argtypes = [vp, ctypes.c_int, vp]
    L.orc_sizes.argtypes = [vp, vp]
    L.orc_scalars.argtypes = [vp, vp]
    L.orc_timers.argtypes = [vp, vp]
    L.orc_reset_timers.argtypes = [vp]
    L.orc_array.restype = vp
    L.orc_array.argtypes = [vp, ctypes.c_char_p, ctypes.POINTER(ctypes.c_long), ctypes.POINTER(ctypes.c_int)]
    L.orc_set_state.argtypes = [vp, vp]
    L.orc_destroy.argtypes = [vp]
    L.orc_roe_flux.argtypes = [ctypes.c_double, vp, vp, ctypes.c_double, ctypes.c_double, vp, vp]
    L.orc_vortex_point.argtypes = [ctypes.POINTER(Fvs2dConfig), ctypes.c_double, ctypes.c_double, ctypes.c_double, vp]
    L.orc_mms_point.argtypes = [ctypes.POINTER(Fvs2dConfig), ctypes.c_double, ctypes.c_double, vp, vp, ctypes.c_int]
    L.orc_limiter.restype = ctypes.c_double
    L.orc_limiter.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double]
    _LIBS[name] = L
    return L


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


class OracleError(RuntimeError):
    pass


class Oracle:
    """One mesh + one configuration of the reference algorithm on the CPU."""

    def __init__(self, mesh: Mesh, cfg: Fvs2dConfig | None = None, fast=False):
        self.L = _lib(fast)
        self.mesh = mesh
        xy = np.ascontiguousarray(mesh.node_xy, dtype=np.float64)
        ptr, node = mesh.csr()
        bn, bt, bc = mesh.bc_arrays()
        self._keep = (xy, ptr, node, bn, bt, bc)
        self.h = self.L.orc_create(mesh.nnodes, mesh.ntri, mesh.nquad, _p(xy), _p(ptr), _p(node),
                                   len(bn), _p(bn), _p(bt), _p(bc))
        err = self.L.orc_last_error(self.h).decode()
        if err:
            raise OracleError(err)
        self.cfg = None
        if cfg is not None:
            self.setup(cfg)

    def _check(self, rc: int):
        if rc:
            raise OracleError(self.L.orc_last_error(self.h).decode())

    def setup(self, cfg: Fvs2dConfig):
        self.cfg = cfg
        self._check(self.L.orc_setup(self.h, ctypes.byref(cfg)))
        return self

    # -- reference-named entry points -------------------------------------------------------
    def initialize_solution(self):
        self._check(self.L.orc_initialize_solution(self.h))

    def compute_residual(self, time: float):
        self._check(self.L.orc_compute_residual(self.h, float(time)))
        return self.array("resid").reshape(-1, 4)

    def time_integration(self, t1: float, nsub: int):
        """-> (res_l2[nsub,4], vortex_err[nsub,14] | None, vortex_xy[nsub,2] | None)"""
        res = np.zeros((nsub, 4))
        if self.cfg.lvortex:
            ve, vxy = np.zeros((nsub, 14)), np.zeros((nsub, 2))
            self._check(self.L.orc_time_integration(self.h, float(t1), nsub, _p(res), _p(ve), _p(vxy)))
            return res, ve, vxy
        self._check(self.L.orc_time_integration(self.h, float(t1), nsub, _p(res), None, None))
        return res, None, None

    def test_resid(self, corrected: bool = False):
        l2, li = np.zeros(4), np.zeros(4)
        self._check(self.L.orc_test_resid(self.h, int(corrected), _p(l2), _p(li)))
        return l2, li

    def vortex_error(self, time: float):
        out, xy = np.zeros(14), np.zeros(2)
        self.L.orc_vortex_error(self.h, float(time), _p(out), _p(xy))
        return out, xy

    def interpolate_cell2node(self, ivar: int) -> np.ndarray:
        """Primitive variable ivar (0 rho, 1 u, 2 v, 3 p) of the current state at the nodes (src/io.f90:122-150)."""
        fv = np.zeros(self.mesh.nnodes)
        self._check(self.L.orc_interpolate_cell2node(self.h, int(ivar), _p(fv)))
        return fv

    def wall_values(self, ib: int) -> np.ndarray:
        """[nedges(ib), 4] = x_f, p_w, p_cell, u_n per edge of boundary ib (src/io.f90:340-449)."""
        bptr = self.array("b_edge_ptr")
        out = np.zeros((int(bptr[ib + 1] - bptr[ib]), 4))
        self._check(self.L.orc_wall_values(self.h, int(ib), _p(out)))
        return out

    # -- data access ------------------------------------------------------------------------
    def array(self, name: str) -> np.ndarray:
        n, k = ctypes.c_long(), ctypes.c_int()
        ptr = self.L.orc_array(self.h, name.encode(), ctypes.byref(n), ctypes.byref(k))
        if k.value < 0:
            raise KeyError(name)
        if not ptr or n.value == 0:
            return np.zeros(0, dtype=np.float64 if k.value == 0 else np.int32)
        ct = ctypes.c_double if k.value == 0 else ctypes.c_int
        return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ct)), shape=(n.value,)).copy()

    def set_state(self, cvar: np.ndarray):
        c = np.ascontiguousarray(cvar, dtype=np.float64)
        assert c.size == 4 * self.mesh.ncells
        self.L.orc_set_state(self.h, _p(c))

    @property
    def cvar(self):
        return self.array("cvar").reshape(-1, 4)

    def sizes(self) -> dict:
        out = np.zeros(10, dtype=np.int32)
        self.L.orc_sizes(self.h, _p(out))
        keys = ["nnodes", "ncells", "nedges", "nedges_intr", "nedges_bndr", "ncells_intr", "ncells_bndr", "nslots",
                "lsq_total", "ggnb_total"]
        return dict(zip(keys, (int(v) for v in out)))

    def scalars(self) -> dict:
        out = np.zeros(6)
        self.L.orc_scalars(self.h, _p(out))
        return dict(zip(["heff1", "heff2", "vol_sum", "vol_green", "lsq_verify_err", "lsq_verified"], out.tolist()))

    def timers(self) -> dict:
        out = np.zeros(4)
        self.L.orc_timers(self.h, _p(out))
        return dict(zip(["grad", "limiter", "flux", "rk"], out.tolist()))

    def reset_timers(self):
        self.L.orc_reset_timers(self.h)

    def close(self):
        if getattr(self, "h", None):
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# pointwise helpers ---------------------------------------------------------------------------
def roe_flux(gamma, pL, pR, nx, ny):
    L = _lib()
    pL = np.ascontiguousarray(pL, dtype=np.float64)
    pR = np.ascontiguousarray(pR, dtype=np.float64)
    f, ws = np.zeros(4), np.zeros(1)
    L.orc_roe_flux(float(gamma), _p(pL), _p(pR), float(nx), float(ny), _p(f), _p(ws))
    return f, float(ws[0])


def vortex_point(cfg, t, x, y):
    pv = np.zeros(4)
    _lib().orc_vortex_point(ctypes.byref(cfg), float(t), float(x), float(y), _p(pv))
    return pv


def mms_point(cfg, x, y, corrected=False):
    sol, rhs = np.zeros(4), np.zeros(4)
    _lib().orc_mms_point(ctypes.byref(cfg), float(x), float(y), _p(sol), _p(rhs), int(corrected))
    return sol, rhs


def limiter(kind: int, a: float, b: float, vol: float) -> float:
    return float(_lib().orc_limiter(int(kind), float(a), float(b), float(vol)))
