"""GPU: a plain C program (tests/c_abi/drive_c_abi.c) that includes include/fvs2d_gpu.h, links libfvs2d_gpu.so and drives
init -> set_mesh -> initialize_solution -> time_integration -> get_state -> compute_residual -> finalize without any
Python in between -- what the Fortran ISO_C_BINDING shim of INTEGRATION.md does.  Its outputs are compared with the CPU
oracle on the same mesh (rebuilt here from the same formulas)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _mesh(nx, ny):
    """the mesh drive_c_abi.c builds (same formulas, same numbering)"""
    from fvs2d_b200.meshio import Mesh
    i, j = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), indexing="xy")
    hx, hy = 20.0 / nx, 10.0 / ny
    x, y = i * hx, j * hy
    inner = (i > 0) & (i < nx) & (j > 0) & (j < ny)
    x = np.where(inner, x + 0.15 * hx * np.sin(1.3 * i + 0.7 * j), x)
    y = np.where(inner, y + 0.15 * hy * np.cos(0.9 * i - 1.1 * j), y)
    xy = np.stack([x, y], axis=-1).reshape(-1, 2)
    tri = np.zeros((2 * nx * ny, 3), dtype=np.int32)
    for jj in range(ny):
        for ii in range(nx):
            q, n00 = jj * nx + ii, jj * (nx + 1) + ii
            n10, n01 = n00 + 1, n00 + nx + 1
            n11 = n01 + 1
            if (ii == nx - 1 and jj == 0) or (ii == 0 and jj == ny - 1):
                tri[2 * q], tri[2 * q + 1] = (n00, n10, n01), (n10, n11, n01)
            else:
                tri[2 * q], tri[2 * q + 1] = (n00, n10, n11), (n00, n11, n01)
    b = [2 * ii for ii in range(nx)]
    b += [2 * (jj * nx + nx - 1) + (1 if jj == 0 else 0) for jj in range(ny)]
    b += [2 * ((ny - 1) * nx + ii) + 1 for ii in range(nx - 1, -1, -1)]
    b += [2 * (jj * nx) + (0 if jj == ny - 1 else 1) for jj in range(ny - 1, -1, -1)]
    return Mesh(xy, tri, np.zeros((0, 4), dtype=np.int32), ["dirichlet"], [np.array(b, dtype=np.int32)])


def test_c_program_over_the_c_abi(tmp_path):
    from fvs2d_b200 import config
    from oracle.oracle import Oracle
    exe = str(tmp_path / "drive_c_abi")
    lib_dir = os.path.join(ROOT, "fvs2d_b200", "csrc")
    cc = subprocess.run(["gcc", "-std=c99", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                         os.path.join(ROOT, "tests", "c_abi", "drive_c_abi.c"), "-o", exe, "-L", lib_dir, "-lfvs2d_gpu", "-lm",
                         f"-Wl,-rpath,{lib_dir}"], capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr
    nx, ny, nsteps = 96, 48, 12
    state, log = str(tmp_path / "state.bin"), str(tmp_path / "log_res.txt")
    run = subprocess.run([exe, str(nx), str(ny), str(nsteps), state, log], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stdout + run.stderr
    assert f"ncells {2 * nx * ny}" in run.stdout
    raw = np.fromfile(state).reshape(2, -1, 4)
    q, r = raw[0], raw[1]
    logs = np.loadtxt(log)
    mesh = _mesh(nx, ny)
    cfg = config.RunInput(grad_cellcntr_imethd=1, lvortex=True, dt=0.005).to_config()
    orc = Oracle(mesh, cfg)
    orc.initialize_solution()
    res_o, ve_o, _ = orc.time_integration(0.0, nsteps)
    assert float((np.abs(q - orc.cvar) / np.abs(orc.cvar).max(axis=0)).max()) <= 1e-10
    assert float((np.abs(logs[:, 1:5] - res_o) / np.abs(res_o)).max()) <= 1e-10
    assert float((np.abs(logs[:, 5] - ve_o[:, 3]) / np.abs(ve_o[:, 3])).max()) <= 1e-8
    r_o = orc.compute_residual(nsteps * 0.005)
    assert float((np.abs(r - r_o) / np.abs(r_o).max(axis=0)).max()) <= 1e-10
