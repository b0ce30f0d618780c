"""Synthetic meshes of the benchmark configurations.

A background grid of ``nx x ny`` quads on ``[0,lx] x [0,ly]``; every quad outside ``quad_band`` is
split into two counter-clockwise triangles along the diagonal (i,j)->(i+1,j+1).  The diagonal is
flipped in the lower-right and upper-left corner quads so that no triangle owns two boundary edges
(the reference's boundary loop needs one boundary edge per boundary cell, src/residual.f90:112,125).
Interior nodes are jittered by U(-jitter*h, jitter*h) per coordinate with
``numpy.random.default_rng(seed)``.  Triangles are numbered first, then quads, as the ``.grid``
format requires (src/grid_procs.f90:98-111).  One boundary, listed in edge-walk order
(bottom, right, top, left).
"""
from __future__ import annotations

import numpy as np

from .meshio import Mesh


def make_mesh(nx: int, ny: int, lx: float = 20.0, ly: float = 10.0, quad_band=None, jitter: float = 0.2,
              seed: int = 12345, bc_type: str = "dirichlet") -> Mesh:
    if quad_band is not None:
        b0, b1 = quad_band
        if not (0 < b0 < b1 < nx):
            raise ValueError("quad_band must lie strictly inside (0, nx): a quad in a domain corner owns 2 boundary edges")
    else:
        b0 = b1 = 0
    if nx < 2 or ny < 2:
        raise ValueError("need at least 2x2 background quads")
    hx, hy = lx / nx, ly / ny
    ii, jj = np.meshgrid(np.arange(nx + 1, dtype=np.int32), np.arange(ny + 1, dtype=np.int32), indexing="xy")  # (ny+1, nx+1)
    xy = np.empty(((nx + 1) * (ny + 1), 2), dtype=np.float64)
    xy[:, 0] = (ii * hx).reshape(-1)
    xy[:, 1] = (jj * hy).reshape(-1)
    if jitter > 0:
        rng = np.random.default_rng(seed)
        d = rng.uniform(-jitter, jitter, size=xy.shape)
        d[:, 0] *= hx
        d[:, 1] *= hy
        interior = ((ii > 0) & (ii < nx) & (jj > 0) & (jj < ny)).reshape(-1)
        d *= interior[:, None]          # boundary nodes stay put (x + 0.0 == x: the same bits as adding to the interior only)
        xy += d
        del d, interior
    if (nx + 1) * (ny + 1) >= 2**31:
        raise ValueError("mesh too large for 32-bit node ids")
    del ii, jj
    # cells: only the quads of each kind are ever expanded (a 69 M-cell mesh is built in seconds); numbering = background
    # quads in row-major order, two triangles ("lower", "upper") per split quad
    q_all = np.arange(nx * ny, dtype=np.int64)
    qi_all = (q_all % nx).astype(np.int32)
    is_quad = (qi_all >= b0) & (qi_all < b1)
    tq, qq = np.flatnonzero(~is_quad), np.flatnonzero(is_quad)
    del q_all, qi_all, is_quad

    def corners(q):
        i, j = q % nx, q // nx
        n00 = (j * (nx + 1) + i).astype(np.int32)
        return n00, n00 + np.int32(1), n00 + np.int32(nx + 2), n00 + np.int32(nx + 1)
    n00, n10, n11, n01 = corners(tq)
    tri = np.empty((2 * tq.size, 3), dtype=np.int32)
    tri[0::2, 0], tri[0::2, 1], tri[0::2, 2] = n00, n10, n11      # lower: bottom, right
    tri[1::2, 0], tri[1::2, 1], tri[1::2, 2] = n00, n11, n01      # upper: top, left
    del n00, n10, n11, n01
    ncol_tri = nx - (b1 - b0)                                      # split quads per row

    def tri_rank(i, j):   # position of quad (i, j) among the split quads
        i = np.asarray(i, dtype=np.int64)
        return np.asarray(j, dtype=np.int64) * ncol_tri + np.where(i < b0, i, i - (b1 - b0))

    def quad_rank(i, j):
        return np.asarray(j, dtype=np.int64) * (b1 - b0) + (np.asarray(i, dtype=np.int64) - b0)
    flips = [(nx - 1, 0), (0, ny - 1)]
    for fi, fj in flips:                                            # the two corner quads with the other diagonal
        r = int(tri_rank(fi, fj))
        c00 = fj * (nx + 1) + fi
        c10, c11, c01 = c00 + 1, c00 + nx + 2, c00 + nx + 1
        tri[2 * r] = (c00, c10, c01)                                # lower: bottom, left
        tri[2 * r + 1] = (c10, c11, c01)                            # upper: right, top
    n00, n10, n11, n01 = corners(qq)
    quad = np.empty((qq.size, 4), dtype=np.int32)
    quad[:, 0], quad[:, 1], quad[:, 2], quad[:, 3] = n00, n10, n11, n01
    del n00, n10, n11, n01, tq, qq
    ntri = tri.shape[0]

    # which piece holds which side of the background quad
    #   unflipped: lo=(n00,n10,n11): bottom,right ; hi=(n00,n11,n01): top,left
    #   flipped:   lo=(n00,n10,n01): bottom,left  ; hi=(n10,n11,n01): right,top
    def side_cell(i, j, side):
        i, j = np.asarray(i, dtype=np.int64), np.asarray(j, dtype=np.int64)
        isq = (i >= b0) & (i < b1)
        lo = np.where(isq, ntri + quad_rank(i, j), 2 * tri_rank(i, j))
        hi = np.where(isq, ntri + quad_rank(i, j), 2 * tri_rank(i, j) + 1)
        f = ((i == nx - 1) & (j == 0)) | ((i == 0) & (j == ny - 1))
        if side == "bottom":
            return lo
        if side == "top":
            return hi
        if side == "right":
            return np.where(f, hi, lo)
        return np.where(f, lo, hi)  # left
    i_all, j_all = np.arange(nx), np.arange(ny)
    walk = np.concatenate([
        side_cell(i_all, np.zeros(nx, int), "bottom"),
        side_cell(np.full(ny, nx - 1), j_all, "right"),
        side_cell(i_all[::-1], np.full(nx, ny - 1), "top"),
        side_cell(np.zeros(ny, int), j_all[::-1], "left"),
    ]).astype(np.int32)
    return Mesh(xy, tri, quad, [bc_type], [walk])


def vortex_tri_mesh(nx: int, ny: int | None = None, **kw) -> Mesh:
    """C3 family: [0,20]x[0,10], nx x nx/2 split quads (nx=2000 -> 4.0 M triangles)."""
    return make_mesh(nx, ny if ny is not None else nx // 2, 20.0, 10.0, None, **kw)


def vortex_mixed_mesh(nx: int, ny: int | None = None, **kw) -> Mesh:
    """C4 family: the middle half of the columns stays quads (nx=9600 -> 46.08 M tri + 23.04 M quad)."""
    return make_mesh(nx, ny if ny is not None else nx // 2, 20.0, 10.0, (nx // 4, 3 * nx // 4), **kw)


def mms_mesh(n: int, **kw) -> Mesh:
    """C5 family: unit square, n x n split quads, all Dirichlet."""
    return make_mesh(n, n, 1.0, 1.0, None, **kw)
