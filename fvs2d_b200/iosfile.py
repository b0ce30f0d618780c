"""The reference's in-house "ios" files: a text header ``<base>.cd`` plus a direct-access binary ``<base>.s4`` /
``<base>.s8`` (src/ios_unstrc.f90).  Python mirror of the four routines the solver and its converter use:

    writecd  src/ios_unstrc.f90:141-290   -> :func:`write_cd`
    writed   src/ios_unstrc.f90:300-404   -> :func:`write_records`
    readcd   src/ios_unstrc.f90:438-562   -> :func:`read_cd`
    readd    src/ios_unstrc.f90:572-681   -> :func:`read_record`

Record k = (nt-1)*mp + np (1-based time level nt, parameter np) holds ``m1`` reals with no record markers; the data are
BIG-endian because the reference is built with ``-convert big_endian`` (CMakeLists.txt:28); ``.s4`` = real*4, ``.s8`` =
real*8.  ``m1`` is the header's "number of nodes" (for ``save.cd`` the reference deliberately stores ncells there,
src/io.f90:95-113).  Host-side file I/O only; nothing here touches the GPU.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np


@dataclass
class IosHeader:
    mnodes: int
    mcells: int
    mp: int
    mt: int
    itimes: list = field(default_factory=list)
    params: list = field(default_factory=list)   # mp parameter names
    info: list = field(default_factory=list)     # free-text info lines

    @property
    def m1(self) -> int:
        return self.mnodes


def _a72(s: str) -> str:
    return "   " + s[:72].ljust(72) + "\n"


def write_cd(base: str, h: IosHeader) -> None:
    """format 1010 / 1020 / 1030 / 1040 / 1050 of writecd (src/ios_unstrc.f90:239-257)."""
    with open(base + ".cd", "w") as f:
        f.write(f"     number of nodes = {h.mnodes}\n     number of cells = {h.mcells}\n"
                f"     number of parameters = {h.mp:5d}\n     number of timesteps  = {h.mt:5d}\n\n"
                f"     Information about file :   ({len(h.info):3d}  info lines )\n")
        for s in h.info:
            f.write(_a72(s))
        f.write("      Information about parameters :\n")
        for s in h.params[:h.mp]:
            f.write(_a72(s))
        f.write("  Numbers of timesteps :\n")
        t = h.itimes[:min(h.mt, 5000)]
        for i in range(0, len(t), 6):
            f.write("".join(f"  {v:10d}" for v in t[i:i + 6]) + "\n")


def read_cd(base: str) -> IosHeader:
    """readcd's fixed-column header (format 1010 at src/ios_unstrc.f90:491-492: 23x,i / 23x,i / 28x,i5 / 28x,i5 // 33x,i3)."""
    with open(base + ".cd", "r") as f:
        L = f.read().split("\n")
    mnodes, mcells = int(L[0][23:]), int(L[1][23:])
    mp, mt = int(L[2][28:33]), int(L[3][28:33])
    minf = int(L[5][33:36])
    info = [s[3:75].rstrip() for s in L[6:6 + minf]]
    o = 6 + minf + 1                                  # skip ' Information about parameters :'
    params = [s[3:75].rstrip() for s in L[o:o + mp]]
    o += mp + 1                                       # skip 'Numbers of timesteps :'
    itimes = [int(t) for t in " ".join(L[o:]).split()][:min(mt, 5000)]
    return IosHeader(mnodes, mcells, mp, mt, itimes, params, info)


def data_path(base: str) -> tuple[str, str]:
    """-> (path, numpy dtype) of the binary part: .s8 if present, else .s4 (mkfname's imach, src/ios_unstrc.f90:72-112)."""
    if os.path.exists(base + ".s8"):
        return base + ".s8", ">f8"
    if os.path.exists(base + ".s4"):
        return base + ".s4", ">f4"
    raise FileNotFoundError(f"{base}.s4 / {base}.s8")


def read_record(base: str, h: IosHeader, nt: int, np_: int) -> np.ndarray:
    """readd: time level nt (1-based), parameter np_ (1-based) -> m1 float64 values."""
    if not (1 <= nt <= h.mt and 1 <= np_ <= h.mp):
        raise IndexError(f"record (nt={nt}, np={np_}) outside (mt={h.mt}, mp={h.mp})")
    path, dt = data_path(base)
    k = (nt - 1) * h.mp + (np_ - 1)
    a = np.fromfile(path, dtype=dt, count=h.m1, offset=k * h.m1 * np.dtype(dt).itemsize)
    if a.size != h.m1:
        raise EOFError(f"{path}: record {k + 1} is short ({a.size} of {h.m1} values)")
    return a.astype(np.float64)


def write_records(base: str, records, double: bool = True, append: bool = False) -> None:
    """writed: consecutive records (iterable of float arrays) in big-endian real*8 (.s8) or real*4 (.s4)."""
    with open(base + (".s8" if double else ".s4"), "ab" if append else "wb") as f:
        for r in records:
            np.asarray(r, dtype=np.float64).astype(">f8" if double else ">f4").tofile(f)
