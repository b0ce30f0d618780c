"""CPU: pins of the oracle.  The reference ships no numeric goldens, so the oracle is pinned by
the reference's own analytic self-checks and by invariants of the scheme:
  1. LSQ linear exactness, |grad(2x+y) - (2,1)| <= 1e-10 on every cell   (src/gradient_lsq.f90:490-529)
  2. sum(vol) by the triangle formula == Green-theorem volume            (src/grid_procs.f90:824-840)
  3. boundary cell / edge counts agree with the .bc file                 (src/grid_procs.f90:722-728,785-791)
  4. the MMS residual-error table has the reference's shape (density row stalls because of the
     src/mms.f90:169 typo; the corrected source converges)               (src/test.f90:498-519)
  5. vortex error magnitudes / log_res of the shipped example match the survey's independent numpy
     restatement to the printed digits.
"""
import numpy as np
import pytest

from conftest import run_input


def _oracle(mesh, cfg):
    from oracle.oracle import Oracle
    return Oracle(mesh, cfg)


def test_mesh_counts_and_volumes(vortex_mesh, naca_mesh):
    v = _oracle(vortex_mesh, run_input("vortex").to_config())
    s, sc = v.sizes(), v.scalars()
    assert (s["nnodes"], s["ncells"], s["nedges"], s["nedges_intr"], s["nedges_bndr"]) == (3734, 7226, 10959, 10719, 240)
    assert s["ncells_bndr"] == 240 and s["ncells_intr"] == 7226 - 240
    assert abs(sc["vol_sum"] - 200.0) < 1e-10 and abs(sc["vol_green"] - 200.0) < 1e-10
    assert sc["lsq_verified"] == 1.0 and sc["lsq_verify_err"] <= 1e-10
    n = _oracle(naca_mesh, run_input("naca").to_config())
    s, sc = n.sizes(), n.scalars()
    assert (s["nnodes"], s["ncells"], s["nedges"], s["nedges_intr"], s["nedges_bndr"]) == (65792, 65536, 131328, 130816, 512)
    assert abs(sc["vol_sum"] - sc["vol_green"]) / sc["vol_sum"] < 1e-12
    assert abs(sc["vol_sum"] - 69637.786) < 1e-2
    assert sc["lsq_verified"] == 1.0
    # c1 < c2 always, c1 never -1, boundary <=> c2 == -1
    c1, c2 = n.array("ec1"), n.array("ec2")
    assert (c1 >= 0).all() and ((c2 > c1) | (c2 < 0)).all()
    # every boundary cell of both shipped meshes has exactly one boundary edge
    assert s["nedges_bndr"] == s["ncells_bndr"]


def test_vortex_example_matches_survey_band(vortex_mesh):
    """SURVEY Appendix D (independent numpy restatement): log_res and rho errors of the shipped example."""
    o = _oracle(vortex_mesh, run_input("vortex").to_config())
    o.initialize_solution()
    res, ve, vxy = o.time_integration(0.0, 100)
    np.testing.assert_allclose(res[0], [1.8714e-05, 3.3647e-05, 5.6717e-05, 6.6516e-05], rtol=2e-4)
    np.testing.assert_allclose(res[9], [4.4743e-06, 3.2869e-05, 5.6427e-05, 1.7389e-05], rtol=2e-4)
    np.testing.assert_allclose(res[99], [4.3587e-06, 3.2669e-05, 5.6414e-05, 1.7100e-05], rtol=2e-4)
    np.testing.assert_allclose(ve[0][1:4], [1.095e-04, 6.68e-06, 1.854e-05], rtol=2e-3)
    np.testing.assert_allclose(ve[99][1:4], [4.126e-04, 2.466e-05, 6.710e-05], rtol=2e-3)
    assert abs(ve[99][0] - 1.0) < 1e-12


def test_roe_flux_invariants():
    from oracle.oracle import roe_flux
    rng = np.random.default_rng(0)
    g = 1.4
    for _ in range(50):
        U = np.array([rng.uniform(0.5, 2), rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(0.5, 2)])
        V = U * rng.uniform(0.8, 1.2, 4)
        th = rng.uniform(0, 2 * np.pi)
        nx, ny = np.cos(th), np.sin(th)
        # consistency: F(U,U,n) = physical flux . n
        f, ws = roe_flux(g, U, U, nx, ny)
        r, u, v, p = U
        un = u * nx + v * ny
        H = g / (g - 1) * p / r + 0.5 * (u * u + v * v)
        np.testing.assert_allclose(f, [r * un, r * un * u + p * nx, r * un * v + p * ny, r * un * H], rtol=1e-13, atol=1e-14)
        assert abs(ws - 0.5 * (abs(un) + np.sqrt(g * p / r))) < 1e-14
        # antisymmetry: F(L,R,n) = -F(R,L,-n)
        f1, _ = roe_flux(g, U, V, nx, ny)
        f2, _ = roe_flux(g, V, U, -nx, -ny)
        np.testing.assert_allclose(f1, -f2, rtol=1e-12, atol=1e-13)


def test_freestream_preservation_and_slip_wall():
    from fvs2d_b200 import config, meshgen
    mesh = meshgen.make_mesh(24, 12, bc_type="freestream")
    for grad in (1, 2, 3):
        cfg = config.RunInput(grad_cellcntr_imethd=grad, lvortex=False, mach_inf=0.5, aoa_inf_deg=0.0).to_config()
        o = _oracle(mesh, cfg)
        o.initialize_solution()
        r = o.compute_residual(0.0)
        assert np.abs(r).max() < 5e-12  # sum n*a = 0 per cell
    # slip wall: zero mass / energy flux through the wall (un_R = -un_L)
    from oracle.oracle import roe_flux
    L = np.array([1.1, 0.3, -0.2, 0.9])
    nx, ny = 0.6, 0.8
    un = L[1] * nx + L[2] * ny
    R = L.copy(); R[1] -= 2 * un * nx; R[2] -= 2 * un * ny
    f, _ = roe_flux(1.4, L, R, nx, ny)
    assert abs(f[0]) < 1e-15 and abs(f[3]) < 1e-15


def test_gradients_exactness():
    """GG gradients annihilate constants; LSQ is exact for linear fields (implied invariants)."""
    from fvs2d_b200 import config, meshgen
    mesh = meshgen.vortex_mixed_mesh(24)
    for grad, st in ((1, "fn"), (2, "fn"), (3, "fn"), (3, "nn")):
        cfg = config.RunInput(grad_cellcntr_imethd=grad, grad_cellcntr_lsq_nghbr=st, lvortex=True).to_config()
        o = _oracle(mesh, cfg)
        xc, yc = o.array("xc"), o.array("yc")
        # primitive field rho = 2 + 0.1x - 0.05y, u = 0.3, v = 0.1 + 0.02x, p = 1 + 0.03y  -> conserved
        rho, u, v, p = 2 + 0.1 * xc - 0.05 * yc, 0.3 + 0 * xc, 0.1 + 0.02 * xc, 1 + 0.03 * yc
        q = np.stack([rho, rho * u, rho * v, p / 0.4 + 0.5 * rho * (u * u + v * v)], 1)
        o.set_state(q)
        o.compute_residual(0.0)
        g = o.array("grad").reshape(2, -1, 4)
        intr = o.array("cell_intr")
        if grad == 3:
            np.testing.assert_allclose(g[0][:, 0], 0.1, atol=1e-11)
            np.testing.assert_allclose(g[1][:, 0], -0.05, atol=1e-11)
            np.testing.assert_allclose(g[1][:, 3], 0.03, atol=1e-11)
        np.testing.assert_allclose(g[0][intr, 1], 0.0, atol=1e-11)  # constant u
        np.testing.assert_allclose(g[1][intr, 1], 0.0, atol=1e-11)


def test_mms_table_shape():
    """C5 on the CPU: GGNB, one residual per level; density row stalls with the reference source, converges
    with the corrected one; momentum/energy rows decrease."""
    from fvs2d_b200 import config, meshgen
    rows, fixed = [], []
    for n in (16, 32, 64):
        mesh = meshgen.mms_mesh(n)
        cfg = config.RunInput(grad_cellcntr_imethd=2, ntstart=0).to_config()
        o = _oracle(mesh, cfg)
        o.initialize_solution()
        rows.append(o.test_resid(False)[0])
        fixed.append(o.test_resid(True)[0])
    rows, fixed = np.array(rows), np.array(fixed)
    assert rows[-1, 0] > 0.8 * rows[0, 0]           # typo: no convergence of the continuity residual error
    assert fixed[-1, 0] < 0.55 * fixed[0, 0]          # corrected source converges
    assert (rows[-1, 1:] < rows[0, 1:]).all()
    np.testing.assert_allclose(rows[:, 1:], fixed[:, 1:], rtol=0, atol=0)  # the typo only touches the density row


def test_rk4_order_on_linear_decay():
    """RK4 table (order 4) integrates q' = R(q) with 4th-order accuracy: refine dt on the vortex problem and
    compare with a dt/4 reference solution."""
    from fvs2d_b200 import config, meshgen
    mesh = meshgen.vortex_tri_mesh(16)
    sols = {}
    for nst in (2, 4, 16):
        cfg = config.RunInput(grad_cellcntr_imethd=1, lvortex=True, dt=0.08 / nst).to_config()
        o = _oracle(mesh, cfg)
        o.initialize_solution()
        o.time_integration(0.0, nst)
        sols[nst] = o.cvar.copy()
    e2 = np.abs(sols[2] - sols[16]).max()
    e4 = np.abs(sols[4] - sols[16]).max()
    assert e2 / e4 > 10.0  # ~16 for 4th order


def test_oracle_builds_agree_fast_vs_strict(vortex_mesh):
    """-Ofast timing build vs the strict parity build: same trajectory to ~1e-12 over 20 smooth steps."""
    from oracle.oracle import Oracle
    cfg = run_input("vortex").to_config()
    a, b = Oracle(vortex_mesh, cfg), Oracle(vortex_mesh, cfg, fast=True)
    a.initialize_solution(); b.initialize_solution()
    ra, _, _ = a.time_integration(0.0, 20)
    rb, _, _ = b.time_integration(0.0, 20)
    assert np.abs(a.cvar - b.cvar).max() < 1e-11
    assert (np.abs(ra - rb) / np.abs(ra)).max() < 1e-9


@pytest.mark.parametrize("case", ["vortex", "naca"])
def test_oracle_all_cores_variant_agrees(case, vortex_mesh, naca_mesh):
    """The OpenMP gather variant (bench context number, not the reference algorithm) differs from the oracle only by
    the order in which a cell's face fluxes are summed."""
    from oracle.oracle import Oracle
    mesh = vortex_mesh if case == "vortex" else naca_mesh
    cfg = run_input(case).to_config()
    a, b = Oracle(mesh, cfg, fast=True), Oracle(mesh, cfg, fast="omp")
    a.initialize_solution(); b.initialize_solution()
    ra, va, _ = a.time_integration(0.0, 5)
    rb, vb, _ = b.time_integration(0.0, 5)
    assert np.abs(a.cvar - b.cvar).max() < 1e-10
    assert (np.abs(ra - rb) / np.abs(ra).clip(1e-300)).max() < 1e-8
    if cfg.lvortex:
        assert np.abs(va - vb).max() < 1e-12


def test_output_path_interpolation_and_wall_values(vortex_mesh, naca_mesh):
    """The oracle's restatement of the output numerics (write_inst_ios src/io.f90:122-150 + src/interpolation.f90:62-123,
    write_inst_cp_un src/io.f90:340-449) against their defining properties: the inverse-distance weights of a node sum
    to one (a uniform field interpolates to itself), the interpolant is the independent numpy evaluation of the same
    formula, and on a uniform freestream the wall values are p_w = p_cell = p_inf, u_n = u_inf . n (zero gradients)."""
    r = run_input("vortex")
    orc = _oracle(vortex_mesh, r.to_config())
    orc.initialize_solution()
    ptr, n2c = orc.array("n2c_ptr"), orc.array("n2c")
    xc, yc = orc.array("xc"), orc.array("yc")
    node = np.repeat(np.arange(vortex_mesh.nnodes), np.diff(ptr))
    w = 1.0 / np.hypot(xc[n2c] - vortex_mesh.node_xy[node, 0], yc[n2c] - vortex_mesh.node_xy[node, 1])
    w /= np.add.reduceat(w, ptr[:-1])[node]
    q = orc.cvar
    pv = np.stack([q[:, 0], q[:, 1] / q[:, 0], q[:, 2] / q[:, 0],
                   (r.gamma - 1.0) * (q[:, 3] - 0.5 * (q[:, 1]**2 + q[:, 2]**2) / q[:, 0])], axis=1)
    for v in range(4):
        np.testing.assert_allclose(orc.interpolate_cell2node(v), np.add.reduceat(w * pv[n2c, v], ptr[:-1]), rtol=1e-13, atol=1e-15)
    orc.set_state(np.tile([1.0, 0.3, -0.1, 2.5], (vortex_mesh.ncells, 1)))
    np.testing.assert_allclose(orc.interpolate_cell2node(0), 1.0, rtol=1e-14)
    np.testing.assert_allclose(orc.interpolate_cell2node(1), 0.3, rtol=1e-14)
    # NACA: freestream state -> zero gradients (LSQ differences vanish identically)
    rn = run_input("naca")
    on = _oracle(naca_mesh, rn.to_config())
    on.initialize_solution()
    ib = naca_mesh.bndry_type.index("slip_wall")
    wv = on.wall_values(ib)
    bptr, bedge = on.array("b_edge_ptr"), on.array("b_edge")
    ie = bedge[bptr[ib]:bptr[ib + 1]]
    assert wv.shape == (256, 4)
    np.testing.assert_array_equal(wv[:, 0], on.array("ex")[ie])
    np.testing.assert_allclose(wv[:, 1], 1.0 / rn.gamma, rtol=1e-15)
    np.testing.assert_allclose(wv[:, 2], 1.0 / rn.gamma, rtol=1e-15)
    np.testing.assert_allclose(wv[:, 3], rn.mach_inf * on.array("enx")[ie], rtol=1e-14, atol=1e-16)
    # closed wall: sum(n a) = 0, so the freestream pressure integrates to zero force
    assert abs((on.array("enx")[ie] * on.array("ea")[ie]).sum()) < 1e-12
