"""CPU: the ios file mirror (src/ios_unstrc.f90 writecd/writed/readcd/readd) and the ios -> Tecplot / VTK converter
(utils/ios2tecplot/ios2tecplot.f90), on the reference's own vortex mesh."""
import os

import numpy as np

from fvs2d_b200 import ios2tecplot, iosfile, meshio


def _make_inst(d, mesh, nt=3, names=("rho", "u")):
    base = os.path.join(d, "inst")
    h = iosfile.IosHeader(mesh.nnodes, mesh.ncells, len(names), nt, [10 * (i + 1) for i in range(nt)], list(names), [])
    iosfile.write_cd(base, h)
    rng = np.random.default_rng(3)
    recs = [rng.standard_normal(mesh.nnodes) * 10.0 ** rng.integers(-5, 5) for _ in range(nt * len(names))]
    iosfile.write_records(base, recs, double=False)
    return base, h, recs


def test_cd_header_is_the_fixed_column_layout_readcd_parses(tmp_path, vortex_mesh):
    base, h, _ = _make_inst(str(tmp_path), vortex_mesh)
    L = open(base + ".cd").read().split("\n")
    # readcd reads with format (23x,i / 23x,i / 28x,i5 / 28x,i5 // 33x,i3), src/ios_unstrc.f90:491-492
    assert L[0][:23] == "     number of nodes = " and int(L[0][23:]) == vortex_mesh.nnodes
    assert L[1][:23] == "     number of cells = " and int(L[1][23:]) == vortex_mesh.ncells
    assert L[2][:28] == "     number of parameters = " and L[2][28:33] == "    2"
    assert L[3][:28] == "     number of timesteps  = " and L[3][28:33] == "    3"
    assert L[4] == "" and L[5][:33] == "     Information about file :   (" and L[5][33:36] == "  0"
    assert L[6] == "      Information about parameters :"
    assert L[7] == "   rho".ljust(75) and L[8] == "   u".ljust(75)
    assert L[9] == "  Numbers of timesteps :" and L[10] == "          10          20          30"
    g = iosfile.read_cd(base)
    assert (g.mnodes, g.mcells, g.mp, g.mt, g.itimes, g.params, g.info) == (h.mnodes, h.mcells, 2, 3, [10, 20, 30], ["rho", "u"], [])


def test_records_are_big_endian_direct_access(tmp_path, vortex_mesh):
    base, h, recs = _make_inst(str(tmp_path), vortex_mesh)
    assert os.path.getsize(base + ".s4") == 4 * vortex_mesh.nnodes * 6
    raw = np.fromfile(base + ".s4", dtype=">f4").reshape(6, -1)
    for nt in (1, 2, 3):
        for ip in (1, 2):
            k = (nt - 1) * 2 + ip - 1                       # record number of readd, src/ios_unstrc.f90:572-611
            a = iosfile.read_record(base, h, nt, ip)
            np.testing.assert_array_equal(a, raw[k].astype(np.float64))
            np.testing.assert_allclose(a, recs[k], rtol=1e-6)
    # real*8 round trip is exact; save.cd stores ncells in the node slot (src/io.f90:95-113)
    sbase = os.path.join(str(tmp_path), "save")
    hs = iosfile.IosHeader(vortex_mesh.ncells, vortex_mesh.nnodes, 4, 1, [4000], ["rho", "rhou", "rhov", "rhoE"],
                           ["number of time-step computed = 4000", "conservatve variables are saved in cell centers",
                            "#ncells and #nodes are replaced", " "])
    iosfile.write_cd(sbase, hs)
    q = np.random.default_rng(1).standard_normal((4, vortex_mesh.ncells))
    iosfile.write_records(sbase, q, double=True)
    g = iosfile.read_cd(sbase)
    assert g.m1 == vortex_mesh.ncells and len(g.info) == 4 and g.info[2] == "#ncells and #nodes are replaced"
    for v in range(4):
        np.testing.assert_array_equal(iosfile.read_record(sbase, g, 1, v + 1), q[v])


def test_fortran_e_descriptor():
    f = ios2tecplot.fortran_e
    assert f(1.0) == "     0.10000000E+01" and f(-0.000123456789) == "    -0.12345679E-03" and f(0.0) == "     0.00000000E+00"
    assert f(9.99999999) == "     0.10000000E+02" and f(12345.678, 16, 8) == "  0.12345678E+05"
    for x in np.random.default_rng(0).standard_normal(200) * 1e3:
        assert abs(float(f(x)) - x) <= 0.5e-8 * 10 ** np.ceil(np.log10(abs(x))) * 1.0000001


def test_ios2tecplot_ascii_and_vtk(tmp_path, vortex_mesh):
    d = str(tmp_path)
    meshio.write_mesh(os.path.join(d, "vortex"), vortex_mesh)
    base, h, recs = _make_inst(d, vortex_mesh)
    out = ios2tecplot.convert(os.path.join(d, "vortex.grid"), base, os.path.join(d, "sol"), (1, 3, 2))
    assert [os.path.basename(p) for p in out] == ["sol_it00001.plt", "sol_it00003.plt"]      # <out>_it<i5.5>.plt
    L = open(out[1]).read().split("\n")
    nn, nc = vortex_mesh.nnodes, vortex_mesh.ncells
    assert L[0] == 'TITLE ="grid_sol"' and L[1] == 'VARIABLES ="x", "y", "rho", "u"'
    assert L[2] == f"ZONE NODES={nn} ELEMENTS={nc} DATAPACKING=POINT, ZONETYPE=FEQUADRILATERAL"
    assert L[3] == "STRANDID=1, SOLUTIONTIME=0.30000000E+01"
    tab = np.array([[float(x) for x in ln.split()] for ln in L[4:4 + nn]])
    assert all(len(ln) == 4 * 20 for ln in L[4:4 + nn])                                        # 4(e19.8,1x)
    np.testing.assert_allclose(tab[:, :2], vortex_mesh.node_xy, rtol=6e-8, atol=1e-300)
    raw = np.fromfile(base + ".s4", dtype=">f4").reshape(6, -1).astype(np.float64)
    np.testing.assert_allclose(tab[:, 2], raw[4], rtol=6e-8)                                  # level 3, rho
    np.testing.assert_allclose(tab[:, 3], raw[5], rtol=6e-8)
    conn = np.array([[int(x) for x in ln.split()] for ln in L[4 + nn:4 + nn + nc]])
    np.testing.assert_array_equal(conn[:, :3], vortex_mesh.tri + 1)
    np.testing.assert_array_equal(conn[:, 3], vortex_mesh.tri[:, 2] + 1)                       # triangle = quad with node 3 repeated
    # separate grid file + solution-only levels
    out2 = ios2tecplot.convert(os.path.join(d, "vortex.grid"), base, os.path.join(d, "s2"), None, together=False)
    assert [os.path.basename(p) for p in out2] == ["s2_grid.plt", "s2_it00001.plt", "s2_it00002.plt", "s2_it00003.plt"]
    assert open(out2[1]).read().split("\n")[1] == 'VARIABLES ="rho", "u"'
    # VTK: header, point count, cell types
    (v,) = ios2tecplot.convert(os.path.join(d, "vortex.grid"), base, os.path.join(d, "v"), (2, 2, 1), vtk=True)
    blob = open(v, "rb").read()
    assert blob.startswith(b"# vtk DataFile Version 3.0") and f"POINTS {nn} double".encode() in blob
    assert f"CELLS {nc} {4 * nc}".encode() in blob and b"SCALARS rho double 1" in blob
    o = blob.index(b"SCALARS u double 1\nLOOKUP_TABLE default\n") + len(b"SCALARS u double 1\nLOOKUP_TABLE default\n")
    np.testing.assert_array_equal(np.frombuffer(blob[o:o + 8 * nn], dtype=">f8"), raw[3])      # level 2, u
    assert ios2tecplot.main([os.path.join(d, "vortex.grid"), base, os.path.join(d, "cli"), "--range", "1", "1", "1"]) == 0
