"""ios -> Tecplot / VTK converter: the non-interactive equivalent of ``utils/ios2tecplot/ios2tecplot.f90``.

The reference's converter reads ``<grid>.grid`` plus an ios pair written by the solver (``inst.cd`` + ``inst.s4|.s8``:
node-interpolated primitive variables, one record per variable per save, src/io.f90:122-150) and writes one binary
``.plt`` per time level through the vendor ``libtecio.so`` (utils/ios2tecplot/ios2tecplot.f90:1-129), file names
``<out>_itNNNNN.plt`` (:117-118), ``SOLUTIONTIME = it`` (:120), variables ``x,y,<names from the .cd>`` when grid and
solution go together (:69-87).  libtecio cannot be linked here, so the output is the ASCII Tecplot layout the
reference itself writes in src/test.f90:606-686 (``test_tecplot_mixed``: FEQUADRILATERAL zone, DATAPACKING=POINT, rows
in ``(n(e19.8,1x))``, triangles as quads with the third node repeated) -- readable by Tecplot's loader -- or legacy VTK
(big-endian binary, for ParaView).

    python -m fvs2d_b200.ios2tecplot GRID.grid IOSBASE OUT [--range MT1 MT2 SKIP] [--separate-grid] [--vtk]
"""
from __future__ import annotations

import argparse
import sys

import numpy as np

from . import iosfile
from .meshio import read_grid


def fortran_e(v: float, w: int = 19, d: int = 8) -> str:
    """Fortran Ew.d edit descriptor: 0.ddddddddE+ee, right-justified."""
    if v == 0.0 or not np.isfinite(v):
        s = "0." + "0" * d + "E+00"
    else:
        m, e = f"{abs(v):.{d - 1}E}".split("E")          # d significant digits: x.ddddddd
        s = ("-" if v < 0 else "") + "0." + m.replace(".", "") + f"E{int(e) + 1:+03d}"
    return s.rjust(w)


def _connectivity(tri: np.ndarray, quad: np.ndarray) -> np.ndarray:
    """1-based, triangles as degenerate quads (src/test.f90:676-682)."""
    t = np.concatenate([tri, tri[:, 2:3]], axis=1) if len(tri) else np.zeros((0, 4), np.int32)
    return np.concatenate([t, quad.reshape(-1, 4)], axis=0) + 1


def write_tecplot(path: str, xy: np.ndarray, tri: np.ndarray, quad: np.ndarray, names, var: np.ndarray,
                  sol_time: float, with_grid: bool = True) -> None:
    nn, nc = xy.shape[0], len(tri) + len(quad)
    cols = ([xy[:, 0], xy[:, 1]] if with_grid else []) + [var[:, k] for k in range(var.shape[1])]
    heads = (['"x"', '"y"'] if with_grid else []) + [f'"{n}"' for n in names]
    with open(path, "w") as f:
        f.write('TITLE ="grid_sol"\n')
        f.write("VARIABLES =" + ", ".join(heads) + "\n")
        f.write(f"ZONE NODES={nn} ELEMENTS={nc} DATAPACKING=POINT, ZONETYPE=FEQUADRILATERAL\n")
        f.write(f"STRANDID=1, SOLUTIONTIME={fortran_e(sol_time, 16, 8).strip()}\n")
        tab = np.stack(cols, axis=1)
        f.write("".join("".join(fortran_e(x) + " " for x in row) + "\n" for row in tab))
        if with_grid:
            f.write("".join(" ".join(f"{int(i):11d}" for i in row) + "\n" for row in _connectivity(tri, quad)))


def write_vtk(path: str, xy: np.ndarray, tri: np.ndarray, quad: np.ndarray, names, var: np.ndarray) -> None:
    nn, nt, nq = xy.shape[0], len(tri), len(quad)
    with open(path, "wb") as f:
        f.write(b"# vtk DataFile Version 3.0\nfvs2d ios2tecplot\nBINARY\nDATASET UNSTRUCTURED_GRID\n")
        f.write(f"POINTS {nn} double\n".encode())
        np.concatenate([xy, np.zeros((nn, 1))], axis=1).astype(">f8").tofile(f)
        f.write(f"\nCELLS {nt + nq} {4 * nt + 5 * nq}\n".encode())
        if nt:
            np.concatenate([np.full((nt, 1), 3), tri], axis=1).astype(">i4").tofile(f)
        if nq:
            np.concatenate([np.full((nq, 1), 4), quad], axis=1).astype(">i4").tofile(f)
        f.write(f"\nCELL_TYPES {nt + nq}\n".encode())
        np.concatenate([np.full(nt, 5), np.full(nq, 9)]).astype(">i4").tofile(f)
        f.write(f"\nPOINT_DATA {nn}\n".encode())
        for k, n in enumerate(names):
            f.write(f"SCALARS {n.replace(' ', '_')} double 1\nLOOKUP_TABLE default\n".encode())
            var[:, k].astype(">f8").tofile(f)
            f.write(b"\n")


def convert(grid: str, ios_base: str, out: str, mt_range=None, together: bool = True, vtk: bool = False) -> list:
    """-> list of files written.  Time levels mt1..mt2 step mt3 (1-based, utils/ios2tecplot/ios2tecplot.f90:49-56)."""
    xy, tri, quad = read_grid(grid)
    h = iosfile.read_cd(ios_base)
    if h.m1 != xy.shape[0]:
        raise SystemExit(f"{ios_base}.cd holds records of {h.m1} values but {grid} has {xy.shape[0]} nodes")
    mt1, mt2, mt3 = mt_range or (1, h.mt, 1)
    names = [n.strip().replace(" ", "") for n in h.params]                   # StripSpaces, :76
    written = []
    if not together and not vtk:
        p = f"{out}_grid.plt"
        write_tecplot(p, xy, tri, quad, [], np.zeros((xy.shape[0], 0)), 0.0, True)
        written.append(p)
    for it in range(mt1, mt2 + 1, mt3):
        var = np.stack([iosfile.read_record(ios_base, h, it, ip) for ip in range(1, h.mp + 1)], axis=1)
        if vtk:
            p = f"{out}_it{it:05d}.vtk"
            write_vtk(p, xy, tri, quad, names, var)
        else:
            p = f"{out}_it{it:05d}.plt"
            write_tecplot(p, xy, tri, quad, names, var, float(it), together)   # sol_time = dble(it), :120
        written.append(p)
    return written


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("grid")
    ap.add_argument("ios_base", help="ios file name without extension (e.g. inst)")
    ap.add_argument("out", help="output name without extension (_itNNNNN.plt is appended)")
    ap.add_argument("--range", type=int, nargs=3, metavar=("MT1", "MT2", "SKIP"))
    ap.add_argument("--separate-grid", action="store_true", help="grid in <out>_grid.plt, solution-only files per level")
    ap.add_argument("--vtk", action="store_true", help="legacy VTK (binary) instead of ASCII Tecplot")
    a = ap.parse_args(argv)
    for p in convert(a.grid, a.ios_base, a.out, a.range, not a.separate_grid, a.vtk):
        print(" written", p)
    return 0


if __name__ == "__main__":
    sys.exit(main())
