"""GPU: the C++ host program fvs2d_gpu.exe (drop-in for `program fvs2d`): same input files in, the reference's
log / ios files out.  Compared with the CPU oracle driven through the same save loop."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import ROOT, run_input

pytestmark = pytest.mark.gpu
EXE = os.path.join(ROOT, "fvs2d_b200", "csrc", "fvs2d_gpu.exe")


def _run(workdir):
    out = subprocess.run([EXE, "0"], cwd=workdir, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    return out.stdout


def _read_be(path, n, dtype=">f8"):
    a = np.fromfile(path, dtype=dtype)
    return a.reshape(-1, n)


def test_vortex_example_files_and_restart(tmp_path, vortex_mesh):
    from fvs2d_b200 import config, meshio
    from oracle.oracle import Oracle
    r = run_input("vortex")
    r.ntimes, r.nsaves = 20, 2
    d = str(tmp_path)
    meshio.write_mesh(os.path.join(d, "vortex"), vortex_mesh)
    config.write_input(os.path.join(d, "fvs2d.input"), r)
    stdout = _run(d)
    assert "   10 time-steps done" in stdout and "   20 time-steps done" in stdout and "o.k." in stdout
    # oracle through the same save loop
    orc = Oracle(vortex_mesh, r.to_config())
    orc.initialize_solution()
    res_o, ve_o, vxy_o = orc.time_integration(0.0, 20)
    # log_res.plt: header + '(i0,1x,e16.8)' + 3 x '(e16.8)' per step (src/runge_kutta.f90:79,169-184)
    lines = open(os.path.join(d, "log_res.plt")).read().splitlines()
    assert lines[0].startswith('variables = "iteration"') and len(lines) == 21
    for i, ln in enumerate(lines[1:]):
        assert len(ln) == len(str(i + 1)) + 1 + 4 * 16
        vals = [float(ln[len(str(i + 1)) + 1 + 16 * k: len(str(i + 1)) + 1 + 16 * (k + 1)]) for k in range(4)]
        assert int(ln.split()[0]) == i + 1
        np.testing.assert_allclose(vals, res_o[i], rtol=6e-8)          # e16.8 prints 0.dddddddd: half a unit of the 8th digit
        assert "E-0" in ln and " 0." in ln                                # Fortran E format: 0.dddddddde-xx
    # log_vortex_err.plt: 6 header lines + '(14(e16.8,1x))' per step (src/mms.f90:283-294,357-361)
    vl = open(os.path.join(d, "log_vortex_err.plt")).read().splitlines()
    assert len(vl) == 6 + 20 and vl[0].startswith('variables = "t"')
    got = np.array([[float(x) for x in ln.split()] for ln in vl[6:]])
    np.testing.assert_allclose(got, ve_o, rtol=2e-7, atol=1e-30)
    xy = np.loadtxt(os.path.join(d, "log_vortex_err_xy.plt"), skiprows=1)
    np.testing.assert_allclose(xy[:, 1:], vxy_o, rtol=1e-14)
    # save.cd / save.s8: 4 big-endian real*8 records of ncells (src/io.f90:95-113,156-178)
    cd = open(os.path.join(d, "save.cd")).read()
    assert f"number of nodes = {vortex_mesh.ncells}" in cd and f"number of cells = {vortex_mesh.nnodes}" in cd  # deliberately swapped
    assert "number of parameters =     4" in cd and "rhoE" in cd and "#ncells and #nodes are replaced" in cd
    sv = _read_be(os.path.join(d, "save.s8"), vortex_mesh.ncells)
    assert sv.shape == (4, vortex_mesh.ncells)
    q_o = orc.cvar
    assert (np.abs(sv.T - q_o) / np.abs(q_o).max(axis=0)).max() <= 1e-10
    # inst.cd / inst.s4: rho and u at the nodes, one record per variable per save (src/io.f90:122-150)
    inst = _read_be(os.path.join(d, "inst.s4"), vortex_mesh.nnodes, ">f4")
    assert inst.shape == (2 * 2, vortex_mesh.nnodes)
    # against the oracle state interpolated to the nodes with inverse-distance weights (src/interpolation.f90:62-123)
    ptr, n2c = orc.array("n2c_ptr"), orc.array("n2c")
    xc, yc = orc.array("xc"), orc.array("yc")
    node = np.repeat(np.arange(vortex_mesh.nnodes), np.diff(ptr))
    w = 1.0 / np.hypot(xc[n2c] - vortex_mesh.node_xy[node, 0], yc[n2c] - vortex_mesh.node_xy[node, 1])
    w /= np.add.reduceat(w, ptr[:-1])[node]
    rho_o, u_o = q_o[:, 0], q_o[:, 1] / q_o[:, 0]
    np.testing.assert_allclose(inst[2], np.add.reduceat(w * rho_o[n2c], ptr[:-1]), rtol=3e-7)   # 2nd save, rho
    np.testing.assert_allclose(inst[3], np.add.reduceat(w * u_o[n2c], ptr[:-1]), rtol=3e-7, atol=1e-7)  # 2nd save, u
    icd = open(os.path.join(d, "inst.cd")).read()
    assert f"number of nodes = {vortex_mesh.nnodes}" in cd or f"number of nodes = {vortex_mesh.nnodes}" in icd
    assert "        10          20" in icd
    assert os.path.exists(os.path.join(d, "log.grid"))
    # the files the host program wrote parse with the readcd / readd mirror and convert like utils/ios2tecplot
    from fvs2d_b200 import ios2tecplot, iosfile
    hd = iosfile.read_cd(os.path.join(d, "inst"))
    assert (hd.mnodes, hd.mcells, hd.mp, hd.mt, hd.itimes) == (vortex_mesh.nnodes, vortex_mesh.ncells, 2, 2, [10, 20])
    np.testing.assert_array_equal(iosfile.read_record(os.path.join(d, "inst"), hd, 2, 1), inst[2].astype(np.float64))
    plt = ios2tecplot.convert(os.path.join(d, "vortex.grid"), os.path.join(d, "inst"), os.path.join(d, "vis"), (2, 2, 1))
    assert open(plt[0]).read().split("\n")[1] == 'VARIABLES ="x", "y", "rho", "u"'
    hs = iosfile.read_cd(os.path.join(d, "save"))
    assert hs.m1 == vortex_mesh.ncells and hs.mp == 4 and hs.itimes == [20]
    # restart: save -> cont, ntstart = 21, 10 more steps == a straight 30-step oracle run
    shutil.copy(os.path.join(d, "save.cd"), os.path.join(d, "cont.cd"))
    shutil.copy(os.path.join(d, "save.s8"), os.path.join(d, "cont.s8"))
    r.ntstart, r.ntimes, r.nsaves = 21, 10, 1
    config.write_input(os.path.join(d, "fvs2d.input"), r)
    _run(d)
    res2, _, _ = orc.time_integration(20 * r.dt, 10)
    sv2 = _read_be(os.path.join(d, "save.s8"), vortex_mesh.ncells)
    assert (np.abs(sv2.T - orc.cvar) / np.abs(orc.cvar).max(axis=0)).max() <= 1e-10
    first = open(os.path.join(d, "log_res.plt")).read().splitlines()[1]
    assert first.split()[0] == "21"                                      # icont + ntstart - 1


def test_mms_mode_appends_error_resid(tmp_path):
    """ntstart = 0: as shipped the reference runs test_resid and stops (src/fvs2d.f90:125-126); running the program
    once per grid level appends one row per level to error_resid.plt (src/test.f90:510-519)."""
    from fvs2d_b200 import config, meshgen, meshio
    from oracle.oracle import Oracle
    d = str(tmp_path)
    r = config.RunInput(grid_base="mms", grad_cellcntr_imethd=2, ntstart=0, lvortex=False)
    config.write_input(os.path.join(d, "fvs2d.input"), r)
    rows = []
    for n in (16, 32):
        mesh = meshgen.mms_mesh(n)
        meshio.write_mesh(os.path.join(d, "mms"), mesh)
        assert "ok" in _run(d)
        orc = Oracle(mesh, r.to_config())
        orc.initialize_solution()
        l2, li = orc.test_resid(False)
        rows.append(np.concatenate([[orc.scalars()["heff1"]], l2, li]))
    lines = open(os.path.join(d, "error_resid.plt")).read().splitlines()
    assert len(lines) == 3 and lines[0].startswith("variables")
    got = np.array([[float(x) for x in ln.split()] for ln in lines[1:]])
    np.testing.assert_allclose(got, np.array(rows), rtol=6e-9)


def test_naca_wall_postprocessing(tmp_path, naca_mesh):
    """C2 (limiter 0 as shipped): log_cp.plt / log_un.plt / log_clcd.plt (src/io.f90:340-449) against the same
    formulas evaluated on the oracle's pvar / grad at the oracle's state; log_res against the oracle."""
    from fvs2d_b200 import config, meshio
    from oracle.oracle import Oracle
    r = run_input("naca")
    r.ntimes, r.nsaves = 10, 1
    d = str(tmp_path)
    meshio.write_mesh(os.path.join(d, "naca0012_omesh"), naca_mesh)
    config.write_input(os.path.join(d, "fvs2d.input"), r)
    _run(d)
    orc = Oracle(naca_mesh, r.to_config())
    orc.initialize_solution()
    res_o, _, _ = orc.time_integration(0.0, 10)
    lines = open(os.path.join(d, "log_res.plt")).read().splitlines()[1:]
    got = np.array([[float(x) for x in ln.split()[1:]] for ln in lines])
    np.testing.assert_allclose(got, res_o, rtol=6e-8)
    # wall quantities from the oracle
    orc.compute_residual(10 * r.dt)
    pv, gr = orc.array("pvar").reshape(-1, 4), orc.array("grad").reshape(2, -1, 4)
    bptr, bedge, ec1 = orc.array("b_edge_ptr"), orc.array("b_edge"), orc.array("ec1")
    ex, ey, ea, enx, eny, xc, yc = (orc.array(n) for n in ("ex", "ey", "ea", "enx", "eny", "xc", "yc"))
    ib = naca_mesh.bndry_type.index("slip_wall")
    ie = bedge[bptr[ib]:bptr[ib + 1]]
    ic = ec1[ie]
    dx, dy = ex[ie] - xc[ic], ey[ie] - yc[ic]
    wall = lambda v: pv[ic, v] + dx * gr[0, ic, v] + dy * gr[1, ic, v]
    cp = 2.0 / r.mach_inf**2 * (wall(3) - 1.0 / r.gamma)
    cp1 = 2.0 / r.mach_inf**2 * (pv[ic, 3] - 1.0 / r.gamma)
    un = wall(1) * enx[ie] + wall(2) * eny[ie]
    cpl = open(os.path.join(d, "log_cp.plt")).read().splitlines()
    assert cpl[0] == 'TITLE     = "cp"' and cpl[2] == f"ZONE I={len(ie)} J=1" and cpl[3].startswith("STRANDID=1, SOLUTIONTIME=")
    tab = np.array([[float(x) for x in ln.split()] for ln in cpl[4:4 + len(ie)]])
    np.testing.assert_allclose(tab[:, 0], ex[ie], rtol=6e-8, atol=1e-12)
    np.testing.assert_allclose(tab[:, 1], cp, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(tab[:, 2], cp1, rtol=1e-6, atol=1e-7)
    unl = [float(x) for x in open(os.path.join(d, "log_un.plt")).read().splitlines()[1].split()]
    np.testing.assert_allclose(unl[1:], [np.abs(un).max(), np.sqrt((un**2).mean()), np.abs(un).mean()], rtol=1e-6)
    cl = [float(x) for x in open(os.path.join(d, "log_clcd.plt")).read().splitlines()[1].split()]
    np.testing.assert_allclose(cl[1:3], [(cp * eny[ie] * ea[ie]).sum(), (cp * enx[ie] * ea[ie]).sum()], rtol=1e-6, atol=1e-7)
