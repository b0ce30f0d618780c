"""Run configuration: the ``fvs2d.input`` / ``fvs2d.vortex`` parser and the C-ABI config struct.

Mirrors the reference's ``input_read`` (``src/input.f90:64-277``): positional, line-oriented,
list-directed reads -- everything after the values on a line is a comment.  The derived flags
(``lgrad_*``, ``limiter_type``, ``lface_reconst_*``) become the integer fields of
:class:`Fvs2dConfig`, which is laid out exactly like ``fvs2d_config`` in ``include/fvs2d_gpu.h``.
"""
from __future__ import annotations

import ctypes
import dataclasses
import math
import os


class Fvs2dConfig(ctypes.Structure):
    """ctypes mirror of ``fvs2d_config`` (include/fvs2d_gpu.h)."""
    _fields_ = [
        ("gamma", ctypes.c_double), ("dt", ctypes.c_double), ("cfl_user", ctypes.c_double),
        ("umuscl_cst", ctypes.c_double), ("lsq_pow", ctypes.c_double),
        ("grad_method", ctypes.c_int), ("lsq_stencil", ctypes.c_int), ("limiter", ctypes.c_int),
        ("recon", ctypes.c_int), ("flux", ctypes.c_int),
        ("rk_nstages", ctypes.c_int), ("rk_order", ctypes.c_int), ("ssprk", ctypes.c_int),
        ("steady", ctypes.c_int), ("lvortex", ctypes.c_int), ("ntstart", ctypes.c_int),
        ("pvar_inf", ctypes.c_double * 4),
        ("vortex_pos", ctypes.c_double * 2), ("vortex_kappa", ctypes.c_double),
        ("vortex_inf", ctypes.c_double * 4),
        ("mms_c", (ctypes.c_double * 4) * 4),
        ("ngpus", ctypes.c_int),
    ]


def _mms_table():
    # src/mms.f90:80-101 (decimal literals are doubles under -r8)
    pi = math.acos(-1.0)
    return [
        (1.12, 0.15, 3.12 * pi, 2.92 * pi),
        (1.32, 0.06, 2.09 * pi, 3.12 * pi),
        (1.18, 0.03, 2.15 * pi, 3.32 * pi),
        (1.62, 0.31, 3.79 * pi, 2.98 * pi),
    ]


def _flogical(tok: str) -> bool:
    t = tok.strip().strip(".").upper()
    if t[:1] == "T":
        return True
    if t[:1] == "F":
        return False
    raise ValueError(f"bad logical {tok!r}")


def _freal(tok: str) -> float:
    return float(tok.strip().replace("d", "e").replace("D", "E"))


def _tokens(line: str, n: int):
    """First n list-directed items of a line (comma and/or blank separated)."""
    out = []
    for chunk in line.replace("\t", " ").split(","):
        out.extend(chunk.split())
        if len(out) >= n:
            break
    if len(out) < n:
        raise ValueError(f"expected {n} values on line: {line!r}")
    return out[:n]


@dataclasses.dataclass
class RunInput:
    """Everything ``input_read`` produces (reference variable names kept)."""
    grid_base: str = "vortex"
    rey: float = 2.0e5
    mach_inf: float = 0.8
    aoa_inf_deg: float = 0.0
    gamma: float = 1.4
    dt: float = 0.01
    ntimes: int = 1
    nsaves: int = 1
    ntstart: int = 1
    lsteady: bool = False
    cfl_user: float = 1.25
    lvortex: bool = False
    lw: tuple = (True, True, False, False)
    cmach_inst: str = "s4"
    grad_cellcntr_imethd: int = 3
    grad_cellcntr_lsq_nghbr: str = "fn"
    grad_cellcntr_lsq_pow: float = 0.0
    grad_limiter_imethd: int = 0
    face_reconst_imethd: int = 2
    umuscl_cst: float = 0.0
    flux_inviscd_imethd: int = 1
    rk_nstages: int = 4
    rk_order: int = 4
    lSSPRK: bool = False
    # fvs2d.vortex (src/mms.f90:55-60)
    vortex_pos: tuple = (5.0, 5.0)
    vortex_kappa: float = 1.0
    vortex_inf: tuple = (1.0, 0.2, 0.0, 1.0)

    # -- derived exactly as input.f90:127-136
    def nsubsteps(self):
        if self.ntimes % self.nsaves == 0:
            return [self.ntimes // self.nsaves] * self.nsaves
        n = self.ntimes // self.nsaves + 1
        return [n] * (self.nsaves - 1) + [self.ntimes - n * (self.nsaves - 1)]

    def validate(self) -> None:
        """The ``stop`` conditions of ``input_read`` (src/input.f90:181-277) as exceptions."""
        if self.cmach_inst.lower() not in ("s4", "s8"):
            raise ValueError("format for regular output files must be either s4 or s8")
        if self.grad_cellcntr_imethd not in (1, 2, 3):
            raise ValueError("check cell-center gradient scheme in input file")
        if self.grad_cellcntr_imethd == 3 and self.grad_cellcntr_lsq_nghbr.lower() not in ("fn", "nn"):
            raise ValueError("check Least-Squares gradient scheme in input file")
        if self.grad_limiter_imethd not in (0, 1, 2, 3):
            raise ValueError("check gradient limiter scheme in input file")
        if self.face_reconst_imethd not in (1, 2, 3):
            raise ValueError("check face reconstruction scheme in input file")
        if self.flux_inviscd_imethd != 1:
            raise ValueError("check inviscid flux discretization scheme in input file")

    def to_config(self, ngpus: int = 1) -> Fvs2dConfig:
        self.validate()
        c = Fvs2dConfig()
        c.gamma, c.dt, c.cfl_user = self.gamma, self.dt, self.cfl_user
        # src/input.f90:248-254: recon 1 or 2 overwrite umuscl_cst with 0
        c.umuscl_cst = self.umuscl_cst if self.face_reconst_imethd == 3 else 0.0
        c.lsq_pow = self.grad_cellcntr_lsq_pow
        c.grad_method = self.grad_cellcntr_imethd
        c.lsq_stencil = 1 if self.grad_cellcntr_lsq_nghbr.lower() == "nn" else 0
        c.limiter = self.grad_limiter_imethd
        c.recon = self.face_reconst_imethd
        c.flux = self.flux_inviscd_imethd
        c.rk_nstages, c.rk_order = self.rk_nstages, self.rk_order
        c.ssprk, c.steady = int(self.lSSPRK), int(self.lsteady)
        # src/input.f90:140: ntstart==0 forces lvortex false
        c.lvortex = int(self.lvortex and self.ntstart != 0)
        c.ntstart = self.ntstart
        # src/data_solution.f90:55-63 (cosd/sind -> degrees)
        a = math.radians(self.aoa_inf_deg)
        cosd = 1.0 if self.aoa_inf_deg == 0.0 else math.cos(a)
        sind = 0.0 if self.aoa_inf_deg == 0.0 else math.sin(a)
        pinf = (1.0, self.mach_inf * cosd, self.mach_inf * sind, 1.0 / self.gamma)
        for i in range(4):
            c.pvar_inf[i] = pinf[i]
            c.vortex_inf[i] = self.vortex_inf[i]
        c.vortex_pos[0], c.vortex_pos[1] = self.vortex_pos
        c.vortex_kappa = self.vortex_kappa
        for i, row in enumerate(_mms_table()):
            for j, v in enumerate(row):
                c.mms_c[i][j] = v
        c.ngpus = ngpus
        return c


def read_input(path: str = "fvs2d.input", vortex_path: str | None = None) -> RunInput:
    """Parse ``fvs2d.input`` (+ ``fvs2d.vortex`` beside it when ``lvortex``), src/input.f90:84-120."""
    if not os.path.exists(path):
        raise FileNotFoundError(f'cannot find "{path}" file!')
    with open(path, "r") as f:
        L = f.read().split("\n")
    r = RunInput()
    # line 4: grid base = text before the first '-', trimmed (src/input.f90:166-175)
    g = L[3]
    r.grid_base = g.split("-")[0].strip()
    r.rey = _freal(_tokens(L[4], 1)[0])
    r.mach_inf = _freal(_tokens(L[5], 1)[0])
    r.aoa_inf_deg = _freal(_tokens(L[6], 1)[0])
    r.gamma = _freal(_tokens(L[7], 1)[0])
    r.dt = _freal(_tokens(L[9], 1)[0])
    r.ntimes = int(_tokens(L[10], 1)[0])
    r.nsaves = int(_tokens(L[11], 1)[0])
    r.ntstart = int(_tokens(L[12], 1)[0])
    t = _tokens(L[13], 2)
    r.lsteady, r.cfl_user = _flogical(t[0]), _freal(t[1])
    r.lvortex = _flogical(_tokens(L[14], 1)[0])
    r.lw = tuple(_flogical(x) for x in _tokens(L[16], 4))
    r.cmach_inst = _tokens(L[17], 1)[0][:2]
    t = _tokens(L[21], 3)
    r.grad_cellcntr_imethd = int(t[0])
    r.grad_cellcntr_lsq_nghbr = t[1].strip("'\"")[:2]
    r.grad_cellcntr_lsq_pow = _freal(t[2])
    r.grad_limiter_imethd = int(_tokens(L[22], 1)[0])
    t = _tokens(L[23], 2)
    r.face_reconst_imethd, r.umuscl_cst = int(t[0]), _freal(t[1])
    r.flux_inviscd_imethd = int(_tokens(L[24], 1)[0])
    r.rk_nstages = int(_tokens(L[26], 1)[0])
    r.rk_order = int(_tokens(L[27], 1)[0])
    r.lSSPRK = _flogical(_tokens(L[28], 1)[0])
    if r.ntstart == 0:
        r.lvortex = False
    r.validate()
    if r.lvortex:
        vp = vortex_path or os.path.join(os.path.dirname(os.path.abspath(path)), "fvs2d.vortex")
        if not os.path.exists(vp):
            raise FileNotFoundError(f'cannot find "{vp}" file!')
        with open(vp, "r") as f:
            V = f.read().split("\n")
        t = _tokens(V[0], 2)
        r.vortex_pos = (_freal(t[0]), _freal(t[1]))
        r.vortex_kappa = _freal(_tokens(V[1], 1)[0])
        r.vortex_inf = tuple(_freal(_tokens(V[i], 1)[0]) for i in range(2, 6))
    return r


_SEP = "=" * 139
_TIL = "~" * 139


def write_input(path: str, r: RunInput, write_vortex: bool = True) -> None:
    """Emit an ``fvs2d.input`` (and ``fvs2d.vortex``) the reference parser would accept."""
    tf = lambda b: "T" if b else "F"
    lines = [
        _SEP, "\tFlow & Simulation Parameters", _SEP,
        f"{r.grid_base}    --- Grid base file name without extension",
        f"{r.rey!r}      --- Reynolds number",
        f"{r.mach_inf!r}       --- Free-stream Mach number",
        f"{r.aoa_inf_deg!r}       --- Free-stream flow angle (deg)",
        f"{r.gamma!r}     --- gamma (cp/cv)",
        _TIL,
        f"{r.dt!r}      --- Time-step: dt",
        f"{r.ntimes}      --- Number of total time-steps: ntimes",
        f"{r.nsaves}        --- Number of output solutions: nsaves",
        f"{r.ntstart}         --- Starting time-step: ntstart",
        f"{tf(r.lsteady)},{r.cfl_user!r}\t\t--- compute steady flow, CFL number",
        f"{tf(r.lvortex)}         --- compute isentropic vortex: T/F",
        _TIL,
        ",".join(tf(b) for b in r.lw) + "\t\t--- Output variables(T/F): rho, u, v, p",
        f"{r.cmach_inst}\t\t\t\t--- Format for transient output files: s4/s8",
        _SEP, "\tTemporal & Spatial Discretization Schemes", _SEP,
        f"{r.grad_cellcntr_imethd},{r.grad_cellcntr_lsq_nghbr},{r.grad_cellcntr_lsq_pow!r}  --- cell-center gradient scheme",
        f"{r.grad_limiter_imethd}         --- gradient limiter scheme",
        f"{r.face_reconst_imethd},{r.umuscl_cst!r}  --- face reconstruction scheme",
        f"{r.flux_inviscd_imethd}         --- inviscid flux scheme",
        _TIL,
        f"{r.rk_nstages}         --- Number of stages for R-K time-integration",
        f"{r.rk_order}         --- Order of accuracy of R-K time-integration",
        f"{tf(r.lSSPRK)}         --- Strong stability preserving formulation for R-K: T/F",
    ]
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
    if write_vortex and r.lvortex:
        vp = os.path.join(os.path.dirname(os.path.abspath(path)), "fvs2d.vortex")
        with open(vp, "w") as f:
            f.write(f"{r.vortex_pos[0]!r},{r.vortex_pos[1]!r} --- x/y center of isentropic vortex\n")
            f.write(f"{r.vortex_kappa!r}     --- Strength of isentropic vortex (Gamma/2/pi)\n")
            for v, nm in zip(r.vortex_inf, ("density", "u-velocity", "v-velocity", "pressure")):
                f.write(f"{v!r}     --- freestream value: {nm}\n")
