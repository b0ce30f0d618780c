"""CPU: the C++ host program's input side (`fvs2d_gpu.exe --check`: fvs2d.input / .grid / .bc parsing, log.fvs2d, log.grid,
the binary grid image) -- everything `program fvs2d` does before the first kernel (src/fvs2d.f90:63-87), through
fvs2d_host_build, which touches no CUDA API."""
import os
import subprocess
import time

import numpy as np
import pytest

from conftest import ROOT, run_input

EXE = os.path.join(ROOT, "fvs2d_b200", "csrc", "fvs2d_gpu.exe")


def _check(d, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([EXE, "--check"], cwd=d, capture_output=True, text=True, timeout=300, env=e)


def _case(d, name, mesh, base):
    from fvs2d_b200 import config, meshio
    r = run_input(name)
    meshio.write_mesh(os.path.join(d, base), mesh)
    config.write_input(os.path.join(d, "fvs2d.input"), r)
    if r.lvortex:
        with open(os.path.join(d, "fvs2d.vortex"), "w") as f:
            f.write("5.0, 5.0\n1.0\n1.0\n0.2\n0.0\n1.0\n")
    return r


def test_check_mode_vortex_logs_and_grid_image(tmp_path, vortex_mesh):
    d = str(tmp_path)
    _case(d, "vortex", vortex_mesh, "vortex")
    out = _check(d)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "nodes=3734 cells=7226 (tri=7226 quad=0) edges=10959 (interior=10719 boundary=240)" in out.stdout
    assert "interior cells=6986 boundary cells=240" in out.stdout and "o.k. (check only)" in out.stdout
    # log.fvs2d: the echo of src/input.f90:283-409 (aN right-justifies, E16.8 prints 0.dddddddd)
    L = open(os.path.join(d, "log.fvs2d")).read().split("\n")
    assert L[0] == "=" * 139 and L[1] == "     FVM2D CODE                       "
    assert L[3] == "                   grid file name: vortex.grid" and L[4] == "                     bc file name: vortex.bc"
    assert L[6] == "                      Mach number:   0.80000000E+00"
    assert L[9] == "                        Time-step:   0.10000000E-01"
    assert L[12] == "Interval to output solution files: 80"
    assert L[15] == "                         t_final :   0.40000000E+02"
    assert " primative varialbes are initialized with isentropic vortex" in L
    assert "  Write out following variables in single precision:" in L and "                                          u-velocity" in L
    assert "          Cell-center gradient method: Unweigghted Least-Squeres based on face neighbor stencil" in L
    assert "                     Gradient limiter: not applied" in L and "  Inviscid flux discretization scheme: Roe" in L
    assert "      Runge-Kutta standard formualtion is employed" in L
    assert " Order of accuracy of Runge-Kutta time-integration: 4" in L
    lg = open(os.path.join(d, "log.grid")).read()
    assert " number of total edges: 10959" in lg and "Sum of the cell volumes via Green theorem: 2.00000000000E+02" in lg
    # the binary image: magic | version, counts | node_xy | cell_node (0-based)
    blob = open(os.path.join(d, "vortex.gridbin"), "rb").read()
    assert blob[:8] == b"FVS2DGRD" and list(np.frombuffer(blob[8:24], dtype="<i4")) == [1, 3734, 7226, 0]
    xy = np.frombuffer(blob[24:24 + 16 * 3734], dtype="<f8").reshape(-1, 2)
    np.testing.assert_array_equal(xy, vortex_mesh.node_xy)
    np.testing.assert_array_equal(np.frombuffer(blob[24 + 16 * 3734:], dtype="<i4").reshape(-1, 3), vortex_mesh.tri)
    # second run reads the image (the text file may even be unreadable garbage as long as it is older)
    open(os.path.join(d, "vortex.grid"), "w").write("garbage\n")
    os.utime(os.path.join(d, "vortex.grid"), (time.time() - 100, time.time() - 100))
    out2 = _check(d)
    assert out2.returncode == 0 and out2.stdout == out.stdout
    # ... but not when the image is disabled, and a newer text file invalidates it
    assert _check(d, {"FVS2D_NO_GRIDBIN": "1"}).returncode != 0
    os.utime(os.path.join(d, "vortex.grid"), (time.time() + 100, time.time() + 100))
    assert _check(d).returncode != 0


def test_check_mode_naca_steady_echo_and_list_directed_quirks(tmp_path, naca_mesh):
    d = str(tmp_path)
    r = _case(d, "naca", naca_mesh, "naca0012_omesh")
    # list-directed quirks: D exponents, commas, a comment after the counts (src/grid_procs.f90:84-111)
    p = os.path.join(d, "naca0012_omesh.grid")
    L = open(p).read().split("\n")
    L[1] = L[1] + "   ! nnodes, ntri, nquad"
    x0, y0 = L[2].split()
    L[2] = f"{float(x0):.16E}".replace("E", "D") + " , " + f"{float(y0):.16E}".replace("E", "d")
    open(p, "w").write("\n".join(L))
    out = _check(d)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "nodes=65792 cells=65536 (tri=0 quad=65536) edges=131328 (interior=130816 boundary=512)" in out.stdout
    log = open(os.path.join(d, "log.fvs2d")).read().split("\n")
    assert "          Steady flow is computed: local time-stepping is employed (input dt is ignored)" in log
    assert " Local dt is computed based on CFL=1.25" in log
    assert " primative varialbes are initialized with freestream values" in log
    assert "          Cell-center gradient method: Unweigghted Least-Squeres based on node neighbor stencil" in log
    assert "           Runge-Kutta SSP formualtion is employed" in log
    assert " Sum of the cell volumes via numerical cal: 6.96377" in open(os.path.join(d, "log.grid")).read()
    assert r.lsteady


def test_check_mode_stops_like_the_reference_on_bad_input(tmp_path, vortex_mesh):
    d = str(tmp_path)
    _case(d, "vortex", vortex_mesh, "vortex")
    bc = open(os.path.join(d, "vortex.bc")).read().replace("dirichlet", "solid_wall")
    open(os.path.join(d, "vortex.bc"), "w").write(bc)
    out = _check(d)
    assert out.returncode == 1 and "not implemented" in out.stdout          # src/residual.f90:206-216
    os.remove(os.path.join(d, "vortex.bc"))
    out = _check(d)
    assert out.returncode == 1 and "cannot find vortex.bc file!" in out.stdout
    os.remove(os.path.join(d, "fvs2d.input"))
    out = _check(d)
    assert out.returncode == 1 and 'cannot find "fvs2d.input" file!' in out.stdout
