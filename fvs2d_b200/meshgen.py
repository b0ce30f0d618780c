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
    xy = np.stack([ii * hx, jj * hy], axis=-1).astype(np.float64).reshape(-1, 2)
    if jitter > 0:
        rng = np.random.default_rng(seed)
        d = rng.uniform(-jitter, jitter, size=xy.shape)
        d[:, 0] *= hx
        d[:, 1] *= hy
        interior = ((ii > 0) & (ii < nx) & (jj > 0) & (jj < ny)).reshape(-1)
        xy[interior] += d[interior]
        del d, interior
    if (nx + 1) * (ny + 1) >= 2**31:
        raise ValueError("mesh too large for 32-bit node ids")
    node = lambda i, j: (j * np.int32(nx + 1) + i).astype(np.int32)
    del ii, jj
    qi, qj = np.meshgrid(np.arange(nx, dtype=np.int32), np.arange(ny, dtype=np.int32), indexing="xy")
    qi, qj = qi.reshape(-1), qj.reshape(-1)
    n00, n10, n11, n01 = node(qi, qj), node(qi + 1, qj), node(qi + 1, qj + 1), node(qi, qj + 1)
    is_quad = (qi >= b0) & (qi < b1)
    is_tri = ~is_quad
    flip = ((qi == nx - 1) & (qj == 0)) | ((qi == 0) & (qj == ny - 1))
    # two triangles per split quad: "lower" then "upper"
    t_lo = np.where(flip[:, None], np.stack([n00, n10, n01], 1), np.stack([n00, n10, n11], 1))
    t_hi = np.where(flip[:, None], np.stack([n10, n11, n01], 1), np.stack([n00, n11, n01], 1))
    tri = np.empty((2 * int(is_tri.sum()), 3), dtype=np.int32)
    tri[0::2] = t_lo[is_tri]
    tri[1::2] = t_hi[is_tri]
    del t_lo, t_hi
    quad = np.stack([n00[is_quad], n10[is_quad], n11[is_quad], n01[is_quad]], 1).astype(np.int32)
    del n00, n10, n11, n01
    ntri = tri.shape[0]
    # cell id of a background quad's pieces
    tri_rank = np.cumsum(is_tri, dtype=np.int32) - 1
    quad_rank = np.cumsum(is_quad, dtype=np.int32) - 1
    lo_id = np.where(is_tri, 2 * tri_rank, ntri + quad_rank)      # piece holding the bottom edge (and, unflipped, the right edge)
    hi_id = np.where(is_tri, 2 * tri_rank + 1, ntri + quad_rank)  # piece holding the top edge (and, unflipped, the left edge)
    Q = lambda i, j: j * nx + i
    # which piece holds which side of the background quad
    #   unflipped: lo=(n00,n10,n11): bottom,right ; hi=(n00,n11,n01): top,left
    #   flipped:   lo=(n00,n10,n01): bottom,left  ; hi=(n10,n11,n01): right,top
    def side_cell(i, j, side):
        q = Q(i, j)
        f = flip[q]
        if side == "bottom":
            return lo_id[q]
        if side == "top":
            return hi_id[q]
        if side == "right":
            return np.where(f, hi_id[q], lo_id[q])
        return np.where(f, lo_id[q], hi_id[q])  # left
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
