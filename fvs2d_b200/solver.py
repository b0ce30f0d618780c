"""Host-side mirror of the reference's solver interface over the C-ABI.

Names follow the Fortran modules: ``input_read`` (src/input.f90:64), ``grid_procs_init``
(src/grid_procs.f90:31), ``initialize_solution`` (src/initialize.f90:19), ``time_integration``
(src/runge_kutta.f90:94), ``compute_residual`` (src/residual.f90:23), ``test_resid`` (src/test.f90:481).
Everything numeric happens in ``libfvs2d_gpu.so`` on the GPU; this class only moves host arrays in/out.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import capi
from .config import Fvs2dConfig, RunInput
from .meshio import Mesh

SIZE_KEYS = ["nnodes", "ncells", "nedges", "nedges_intr", "nedges_bndr", "ncells_intr", "ncells_bndr", "ncells_own",
             "ncells_local", "nedges_local"]
SCALAR_KEYS = ["heff1", "heff2", "vol_sum", "vol_green", "lsq_verify_err", "device_bytes"]


class Fvs2dGpu:
    """One process, one GPU, one mesh (the library keeps a single context per process)."""

    def __init__(self, cfg: Fvs2dConfig | RunInput, device: int = -1, comm=None):
        """``comm`` = (rank, nranks, unique_id_bytes) to join an NCCL communicator."""
        self.L = capi.lib()
        self.cfg = cfg.to_config() if isinstance(cfg, RunInput) else cfg
        capi.check(self.L.fvs2d_gpu_init(ctypes.byref(self.cfg), device))
        self.rank, self.nranks = 0, 1
        if comm is not None:
            self.rank, self.nranks, uid = comm
            buf = ctypes.create_string_buffer(bytes(uid), 128)
            capi.check(self.L.fvs2d_gpu_comm_init(self.rank, self.nranks, buf))
        self.ncells = 0

    @staticmethod
    def comm_unique_id() -> bytes:
        buf = ctypes.create_string_buffer(128)
        capi.check(capi.lib().fvs2d_gpu_comm_unique_id(buf))
        return buf.raw

    # -- grid_procs_init + gradient_init -----------------------------------------------------
    def set_mesh(self, mesh: Mesh):
        xy = np.ascontiguousarray(mesh.node_xy, dtype=np.float64)
        cptr, cnode = mesh.csr()
        bn, bt, bc = mesh.bc_arrays()
        capi.check(self.L.fvs2d_gpu_set_mesh(mesh.nnodes, mesh.ntri, mesh.nquad, capi.ptr(xy), capi.ptr(cptr),
                                             capi.ptr(cnode), len(bn), capi.ptr(bn), capi.ptr(bt), capi.ptr(bc)))
        self.ncells = mesh.ncells
        self._bndry_ncells = [len(c) for c in mesh.bndry_cell]
        return self

    def set_lsq(self, ptr, cell, w, coef):
        """Use the caller's least-squares table (the reference's public ``lsq(:)``, src/gradient_lsq.f90:16-27) instead of
        the library's: CSR over the cells in the original numbering, 0-based ids, ``coef`` shaped [entries, 2]."""
        ptr = np.ascontiguousarray(ptr, dtype=np.int32)
        cell = np.ascontiguousarray(cell, dtype=np.int32)
        w = np.ascontiguousarray(w, dtype=np.float64)
        coef = np.ascontiguousarray(coef, dtype=np.float64)
        assert ptr.size == self.ncells + 1 and cell.size == w.size == ptr[-1] and coef.size == 2 * ptr[-1]
        capi.check(self.L.fvs2d_gpu_set_lsq(capi.ptr(ptr), capi.ptr(cell), capi.ptr(w), capi.ptr(coef)))
        return self

    def sizes(self) -> dict:
        out = np.zeros(10, dtype=np.int32)
        capi.check(self.L.fvs2d_gpu_sizes(capi.ptr(out)))
        return dict(zip(SIZE_KEYS, (int(v) for v in out)))

    def scalars(self) -> dict:
        out = np.zeros(6)
        capi.check(self.L.fvs2d_gpu_scalars(capi.ptr(out)))
        return dict(zip(SCALAR_KEYS, out.tolist()))

    # -- initialize_solution / state ---------------------------------------------------------
    def initialize_solution(self):
        capi.check(self.L.fvs2d_gpu_initialize_solution())

    def set_state(self, cvar):
        if isinstance(cvar, np.ndarray):
            cvar = np.ascontiguousarray(cvar, dtype=np.float64)
            assert cvar.size == 4 * self.ncells
        capi.check(self.L.fvs2d_gpu_set_state(capi.ptr(cvar)))

    def get_state(self, out=None):
        if out is None:
            out = np.zeros((self.ncells, 4))
        capi.check(self.L.fvs2d_gpu_get_state(capi.ptr(out)))
        return out

    def set_state_local(self, cvar_own):
        """owned cells only, library cell order (see ``capi.mesh_array("orig_id")``)."""
        capi.check(self.L.fvs2d_gpu_set_state_local(capi.ptr(cvar_own)))

    def get_state_local(self, out):
        capi.check(self.L.fvs2d_gpu_get_state_local(capi.ptr(out)))
        return out

    # -- the hot path ------------------------------------------------------------------------
    def time_integration(self, t1: float, nsub: int, logs: bool = True):
        """-> (res_l2[nsub,4], vortex_err[nsub,14] | None, vortex_xy[nsub,2] | None); logs=False skips
        every device->host copy (state and logs stay resident)."""
        if not logs:
            capi.check(self.L.fvs2d_gpu_time_integration(float(t1), int(nsub), None, None, None))
            return None, None, None
        res = np.zeros((nsub, 4))
        ve = vxy = None
        if self.cfg.lvortex:
            ve, vxy = np.zeros((nsub, 14)), np.zeros((nsub, 2))
        capi.check(self.L.fvs2d_gpu_time_integration(float(t1), int(nsub), capi.ptr(res), capi.ptr(ve), capi.ptr(vxy)))
        return res, ve, vxy

    def compute_residual(self, time: float, want_ws: bool = False):
        resid = np.zeros((self.ncells, 4))
        ws = np.zeros(self.ncells) if want_ws else None
        capi.check(self.L.fvs2d_gpu_compute_residual(float(time), capi.ptr(resid), capi.ptr(ws)))
        return (resid, ws) if want_ws else resid

    def get_aux(self):
        """-> pvar[nc,4], grad[2,nc,4] (Fortran grad(ivar,ic,idim)), phi_lim[nc] of the last residual."""
        pv, gr, ph = np.zeros((self.ncells, 4)), np.zeros((2, self.ncells, 4)), np.zeros(self.ncells)
        capi.check(self.L.fvs2d_gpu_get_aux(capi.ptr(pv), capi.ptr(gr), capi.ptr(ph)))
        return pv, gr, ph

    def test_resid(self, corrected: bool = False):
        l2, li = np.zeros(4), np.zeros(4)
        capi.check(self.L.fvs2d_gpu_test_resid(int(corrected), capi.ptr(l2), capi.ptr(li)))
        return l2, li

    # -- output path (write_inst_ios / write_inst_cp_un numerics, src/io.f90:122-150, 340-449) --
    def interpolate_cell2node(self, select=(1, 1, 1, 1)) -> np.ndarray:
        """-> [nselected, nnodes]: the selected primitive variables (rho, u, v, p) of the current state at the
        nodes, inverse-distance weighted (src/interpolation.f90:62-123)."""
        sel = np.asarray([1 if s else 0 for s in select], dtype=np.int32)
        out = np.zeros((int(sel.sum()), self.sizes()["nnodes"]))
        if out.size:
            capi.check(self.L.fvs2d_gpu_interpolate_cell2node(capi.ptr(sel), capi.ptr(out)))
        return out

    def wall_values(self, ib: int) -> np.ndarray:
        """-> [nedges(ib), 4] = x_f, p_w, p_cell, u_n per edge of boundary ``ib`` (src/io.f90:340-449)."""
        if self.nranks == 1:
            bptr = capi.mesh_array("b_edge_ptr")
            n = int(bptr[ib + 1] - bptr[ib]) if 0 <= ib < len(bptr) - 1 else 0   # out of range: the library reports it
        else:   # one edge per listed cell (see include/fvs2d_gpu.h); entries of edges other ranks own stay zero
            n = self._bndry_ncells[ib] if 0 <= ib < len(self._bndry_ncells) else 0
        out = np.zeros((max(n, 1), 4))
        capi.check(self.L.fvs2d_gpu_wall_values(int(ib), capi.ptr(out)))
        return out[:n]

    # -- instrumentation ---------------------------------------------------------------------
    def set_option(self, key: str, value: int):
        capi.check(self.L.fvs2d_gpu_set_option(key.encode(), int(value)))

    def last_timing(self):
        ms = np.zeros(4)
        n = ctypes.c_long()
        capi.check(self.L.fvs2d_gpu_last_timing(capi.ptr(ms), ctypes.byref(n)))
        return dict(total_ms=ms[0], grad_ms=ms[1], flux_ms=ms[2], other_ms=ms[3], launches=int(n.value))

    def close(self):
        self.L.fvs2d_gpu_finalize()


def host_build(cfg: Fvs2dConfig, mesh: Mesh, rank: int = 0, nranks: int = 1) -> None:
    """CPU-only half of ``set_mesh`` (no GPU needed); inspect with :func:`capi.mesh_array`."""
    L = capi.lib()
    xy = np.ascontiguousarray(mesh.node_xy, dtype=np.float64)
    cptr, cnode = mesh.csr()
    bn, bt, bc = mesh.bc_arrays()
    capi.check(L.fvs2d_host_build(ctypes.byref(cfg), rank, nranks, mesh.nnodes, mesh.ntri, mesh.nquad, capi.ptr(xy),
                                  capi.ptr(cptr), capi.ptr(cnode), len(bn), capi.ptr(bn), capi.ptr(bt), capi.ptr(bc)))
