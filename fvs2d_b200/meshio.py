"""Readers/writers for the reference's ``.grid`` / ``.bc`` mesh files.

Format (reference ``src/grid_procs.f90:84-111`` and ``:146-159``):

``.grid``: line 1 comment; line 2 ``nnodes ncells_tri ncells_quad``; ``nnodes`` lines ``x y``;
then the triangles (3 node ids, 1-based), then the quads (4 ids).
``.bc``: line 1 ``nbndries``; one line ``ncells type`` per boundary; then every boundary's cell ids
(1-based), one per line.

In memory a mesh is a :class:`Mesh`: 0-based ids, CSR ``cell_ptr/cell_node`` with triangles first.
"""
from __future__ import annotations

import dataclasses
import io
import os

import numpy as np

BC_TYPES = {"freestream": 1, "slip_wall": 2, "solid_wall": 3, "dirichlet": 4}
BC_NAMES = {v: k for k, v in BC_TYPES.items()}


@dataclasses.dataclass
class Mesh:
    node_xy: np.ndarray      # (nnodes, 2) float64
    tri: np.ndarray          # (ntri, 3) int32, 0-based
    quad: np.ndarray         # (nquad, 4) int32, 0-based
    bndry_type: list         # list[str]
    bndry_cell: list         # list[np.ndarray int32], 0-based cell ids (triangles first numbering)

    @property
    def nnodes(self) -> int:
        return int(self.node_xy.shape[0])

    @property
    def ntri(self) -> int:
        return int(self.tri.shape[0])

    @property
    def nquad(self) -> int:
        return int(self.quad.shape[0])

    @property
    def ncells(self) -> int:
        return self.ntri + self.nquad

    def csr(self):
        """(cell_ptr[ncells+1], cell_node[3*ntri+4*nquad]) int32, triangles first."""
        assert 3 * self.ntri + 4 * self.nquad < 2**31
        ptr = np.empty(self.ncells + 1, dtype=np.int32)
        ptr[:self.ntri + 1] = np.arange(0, 3 * self.ntri + 1, 3, dtype=np.int32)
        ptr[self.ntri:] = np.arange(3 * self.ntri, 3 * self.ntri + 4 * self.nquad + 1, 4, dtype=np.int32)
        node = np.empty(3 * self.ntri + 4 * self.nquad, dtype=np.int32)
        node[:3 * self.ntri] = self.tri.reshape(-1)
        node[3 * self.ntri:] = self.quad.reshape(-1)
        return ptr, node

    def bc_arrays(self):
        """(bndry_ncells, bndry_type_enum, bndry_cell_concat) int32."""
        n = np.array([len(c) for c in self.bndry_cell], dtype=np.int32)
        t = np.array([BC_TYPES[s] for s in self.bndry_type], dtype=np.int32)
        c = (np.concatenate(self.bndry_cell) if len(self.bndry_cell) else np.zeros(0)).astype(np.int32)
        return n, t, c


def read_grid(path: str):
    """Parse a ``.grid`` file -> (node_xy, tri, quad), ids converted to 0-based."""
    with open(path, "r") as f:
        f.readline()
        nnodes, ntri, nquad = (int(t) for t in f.readline().replace(",", " ").split()[:3])
        body = f.read()
    # list-directed Fortran reads accept D exponents
    body = body.replace("D", "E").replace("d", "e")
    vals = np.array(body.split(), dtype=np.float64)
    need = 2 * nnodes + 3 * ntri + 4 * nquad
    if vals.size < need:
        raise ValueError(f"{path}: expected {need} numbers, found {vals.size}")
    xy = vals[:2 * nnodes].reshape(nnodes, 2).copy()
    o = 2 * nnodes
    tri = vals[o:o + 3 * ntri].astype(np.int32).reshape(ntri, 3) - 1
    o += 3 * ntri
    quad = vals[o:o + 4 * nquad].astype(np.int32).reshape(nquad, 4) - 1
    return xy, tri, quad


def read_bc(path: str):
    """Parse a ``.bc`` file -> (types list[str], cells list[int32 arrays, 0-based])."""
    with open(path, "r") as f:
        nb = int(f.readline().split()[0])
        counts, types = [], []
        for _ in range(nb):
            tok = f.readline().replace(",", " ").split()
            counts.append(int(tok[0]))
            types.append(tok[1].strip("'\""))
        cells = []
        for n in counts:
            c = np.array([int(f.readline().split()[0]) for _ in range(n)], dtype=np.int32) - 1
            cells.append(c)
    for t in types:
        if t not in BC_TYPES:
            raise ValueError(f"Boundary condition={t}  not implemented")
    return types, cells


def read_mesh(base: str) -> Mesh:
    """Read ``<base>.grid`` + ``<base>.bc`` (reference ``grid_read``/``grid_bc_read``)."""
    xy, tri, quad = read_grid(base + ".grid")
    types, cells = read_bc(base + ".bc")
    return Mesh(xy, tri, quad, types, cells)


def write_mesh(base: str, mesh: Mesh) -> None:
    """Write ``<base>.grid`` / ``<base>.bc`` in the reference's text format (1-based ids)."""
    with open(base + ".grid", "w") as f:
        f.write(" #nodes, #cells_tri, #cells_quad\n")
        f.write(f"{mesh.nnodes:12d}{mesh.ntri:12d}{mesh.nquad:12d}\n")
        buf = io.StringIO()
        np.savetxt(buf, mesh.node_xy, fmt="%25.17E")
        if mesh.ntri:
            np.savetxt(buf, mesh.tri + 1, fmt="%d")
        if mesh.nquad:
            np.savetxt(buf, mesh.quad + 1, fmt="%d")
        f.write(buf.getvalue())
    with open(base + ".bc", "w") as f:
        f.write(f"{len(mesh.bndry_type)}     !-- #boundaries\n")
        for i, (t, c) in enumerate(zip(mesh.bndry_type, mesh.bndry_cell)):
            f.write(f"{len(c)}   {t}     !-- boundary {i + 1}\n")
        for c in mesh.bndry_cell:
            for v in c:
                f.write(f"{int(v) + 1}\n")


def save_npz(path: str, mesh: Mesh) -> None:
    d = dict(node_xy=mesh.node_xy, tri=mesh.tri, quad=mesh.quad,
             bndry_type=np.array(mesh.bndry_type), nb=len(mesh.bndry_type))
    for i, c in enumerate(mesh.bndry_cell):
        d[f"bndry_cell_{i}"] = c
    np.savez_compressed(path, **d)


def load_npz(path: str) -> Mesh:
    z = np.load(path)
    nb = int(z["nb"])
    return Mesh(z["node_xy"], z["tri"].astype(np.int32).reshape(-1, 3), z["quad"].astype(np.int32).reshape(-1, 4),
                [str(s) for s in z["bndry_type"]], [z[f"bndry_cell_{i}"].astype(np.int32) for i in range(nb)])
