"""GPU parity: the CUDA path (through the C-ABI) against the CPU oracle on the same inputs.

Tolerance (BASELINE.json north_star): density/momentum/energy fields and the L2 residual history agree to
<= 1e-10 relative after N steps (N stated per test); relative = max|a-b| / max|b| per variable.
"""
import numpy as np
import pytest

from conftest import run_input

pytestmark = pytest.mark.gpu

TOL = 1e-10


def _rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    scale = np.abs(b).reshape(-1, b.shape[-1]).max(axis=0) if b.ndim > 1 else np.abs(b).max()
    return float((np.abs(a - b) / np.maximum(scale, 1e-300)).max())


def _pair(mesh, cfg):
    from fvs2d_b200 import solver
    from oracle.oracle import Oracle
    gpu = solver.Fvs2dGpu(cfg, device=0)
    gpu.set_mesh(mesh)
    orc = Oracle(mesh, cfg)
    return gpu, orc


def _run_steps(mesh, cfg, nsteps, t1=0.0):
    gpu, orc = _pair(mesh, cfg)
    gpu.initialize_solution()
    orc.initialize_solution()
    assert _rel(gpu.get_state(), orc.cvar) < 1e-14  # identical initial condition
    out_g = gpu.time_integration(t1, nsteps)
    out_o = orc.time_integration(t1, nsteps)
    q_g, q_o = gpu.get_state(), orc.cvar
    gpu.close()
    return out_g, out_o, q_g, q_o


def test_c1_vortex_example_100_steps(vortex_mesh):
    """C1: examples/isentropic_vortex as shipped (LSQ-fn, no limiter, upwind-2nd, Roe, RK4, dt=0.01), N=100."""
    cfg = run_input("vortex").to_config()
    (res, ve, vxy), (res_o, ve_o, vxy_o), q, q_o = _run_steps(vortex_mesh, cfg, 100)
    assert _rel(q, q_o) <= TOL
    assert float((np.abs(res - res_o) / np.abs(res_o)).max()) <= TOL
    assert float((np.abs(ve - ve_o) / np.maximum(np.abs(ve_o), 1e-300)).max()) <= 1e-8  # error norms are differences of O(1) values
    assert np.abs(vxy - vxy_o).max() == 0.0  # same argmax cell


def test_c2_naca_ssprk_steady_100_steps(naca_mesh):
    """C2 as shipped (limiter 0): NACA0012 O-mesh, LSQ-nn, SSPRK(4,2) steady local-dt CFL 1.25, N=100 from freestream."""
    cfg = run_input("naca").to_config()
    (res, _, _), (res_o, _, _), q, q_o = _run_steps(naca_mesh, cfg, 100)
    assert _rel(q, q_o) <= TOL
    assert float((np.abs(res - res_o) / np.abs(res_o)).max()) <= 1e-9


@pytest.mark.parametrize("limiter", [1, 2])
def test_c2_naca_limited(naca_mesh, limiter):
    """C2 with Venkatakrishnan / Barth: the impulsively started, limited transonic run amplifies a last-bit
    difference about tenfold per time step (measured: 2e-15 after 1 step, 1e-9 after 10, O(1e-2) after 20 --
    scripts/diag_limiter.py; two builds of the CPU oracle diverge the same way).  So the trajectory is
    compared at N=5 (<= 1e-10), and at the developed state after N=100 the GPU residual, limiter and
    gradients are compared with the oracle evaluated on the SAME state (single-evaluation parity)."""
    from fvs2d_b200 import solver
    from oracle.oracle import Oracle
    r = run_input("naca")
    r.grad_limiter_imethd = limiter
    cfg = r.to_config()
    gpu, orc = _pair(naca_mesh, cfg)
    gpu.initialize_solution()
    orc.initialize_solution()
    res, _, _ = gpu.time_integration(0.0, 5)
    res_o, _, _ = orc.time_integration(0.0, 5)
    assert _rel(gpu.get_state(), orc.cvar) <= TOL
    assert float((np.abs(res - res_o) / np.abs(res_o)).max()) <= 1e-9
    gpu.time_integration(5 * r.dt, 95)
    q = gpu.get_state()
    assert np.isfinite(q).all()
    orc.set_state(q)
    resid_o = orc.compute_residual(0.0)
    resid, ws = gpu.compute_residual(0.0, want_ws=True)
    pv, gr, ph = gpu.get_aux()
    g_o = orc.array("grad").reshape(2, -1, 4)
    ph_o = orc.array("phi_lim")
    assert (ph_o < 0.999).sum() > 500  # the limiter is really active
    assert _rel(resid, resid_o) <= 1e-11
    assert _rel(ws, orc.array("ws_nrml")) <= 1e-13
    # the limited gradient is what enters the fluxes (Barth's phi itself is noise where grad ~ 0)
    assert np.abs(ph[None, :, None] * gr - ph_o[None, :, None] * g_o).max() / np.abs(g_o).max() <= 1e-11
    if limiter == 1:
        assert np.abs(ph - ph_o).max() <= 1e-10
    gpu.close()


@pytest.mark.parametrize("grad,stencil", [(1, "fn"), (2, "fn"), (3, "fn"), (3, "nn")])
@pytest.mark.parametrize("mixed", [False, True])
def test_residual_all_gradients(grad, stencil, mixed):
    """compute_residual on jittered tri and mixed tri/quad meshes for GGCB / GGNB / LSQ-fn / LSQ-nn:
    resid, ws_nrml, pvar, grad, phi against the oracle."""
    from fvs2d_b200 import config, meshgen
    mesh = (meshgen.vortex_mixed_mesh if mixed else meshgen.vortex_tri_mesh)(48)
    r = config.RunInput(grad_cellcntr_imethd=grad, grad_cellcntr_lsq_nghbr=stencil, lvortex=True, dt=0.005)
    cfg = r.to_config()
    gpu, orc = _pair(mesh, cfg)
    gpu.initialize_solution()
    orc.initialize_solution()
    resid, ws = gpu.compute_residual(0.3, want_ws=True)
    resid_o = orc.compute_residual(0.3)
    pv, gr, ph = gpu.get_aux()
    assert _rel(pv, orc.array("pvar").reshape(-1, 4)) <= 1e-14
    g_o = orc.array("grad").reshape(2, -1, 4)
    assert np.abs(gr - g_o).max() / np.abs(g_o).max() <= 1e-12
    assert _rel(resid, resid_o) <= 1e-11
    assert _rel(ws, orc.array("ws_nrml")) <= 1e-13
    assert np.array_equal(ph, orc.array("phi_lim"))
    gpu.close()


@pytest.mark.parametrize("limiter", [1, 2, 3])
def test_limiters(limiter):
    """Venkatakrishnan / Barth-Jespersen / van Albada on a vortex: limited gradients and the residual.
    van Albada as coded in the reference (src/gradient_limiter.f90:121-127) divides by (b+eps2) and can return
    huge negative phi, so NaNs appear in the residual of both implementations: same NaN pattern required."""
    from fvs2d_b200 import config, meshgen
    mesh = meshgen.vortex_tri_mesh(48)
    r = config.RunInput(grad_cellcntr_imethd=3, grad_cellcntr_lsq_nghbr="nn", grad_limiter_imethd=limiter, lvortex=True, dt=0.005)
    cfg = r.to_config()
    gpu, orc = _pair(mesh, cfg)
    gpu.initialize_solution()
    orc.initialize_solution()
    resid = gpu.compute_residual(0.0)
    resid_o = orc.compute_residual(0.0)
    _, gr, ph = gpu.get_aux()
    ph_o = orc.array("phi_lim")
    g_o = orc.array("grad").reshape(2, -1, 4)
    lim_g, lim_o = ph[None, :, None] * gr, ph_o[None, :, None] * g_o
    assert np.abs(lim_g - lim_o).max() / np.abs(lim_o).max() <= (1e-9 if limiter != 2 else 1e-6)  # Barth: phi is noise where grad ~ 1e-8
    if limiter == 1:
        assert np.abs(ph - ph_o).max() <= 1e-9
    nan_g, nan_o = ~np.isfinite(resid), ~np.isfinite(resid_o)
    assert np.array_equal(nan_g, nan_o)
    ok = ~nan_o
    scale = np.abs(np.where(ok, resid_o, 0)).max(axis=0)
    assert (np.abs(np.where(ok, resid - resid_o, 0)) / scale).max() <= (1e-8 if limiter != 2 else 1e-6)
    gpu.close()


@pytest.mark.parametrize("recon,kappa", [(1, 0.0), (3, 1.0 / 3.0), (3, -1.0)])
def test_reconstructions(recon, kappa):
    """first-order upwind and UMUSCL (kappa = 1/3, -1)."""
    from fvs2d_b200 import config, meshgen
    mesh = meshgen.vortex_mixed_mesh(40)
    r = config.RunInput(grad_cellcntr_imethd=1, face_reconst_imethd=recon, umuscl_cst=kappa, lvortex=True, dt=0.005)
    cfg = r.to_config()
    (res, _, _), (res_o, _, _), q, q_o = _run_steps(mesh, cfg, 10)
    assert _rel(q, q_o) <= TOL
    assert float((np.abs(res - res_o) / np.abs(res_o)).max()) <= TOL


@pytest.mark.parametrize("ssprk,steady,order", [(0, 0, 1), (0, 0, 2), (0, 0, 3), (1, 0, 2), (0, 1, 4), (1, 1, 2)])
def test_integrators(ssprk, steady, order):
    """all four integrators (RK / SSPRK x unsteady / steady local-dt) and the RK coefficient tables."""
    from fvs2d_b200 import config, meshgen
    mesh = meshgen.vortex_tri_mesh(40)
    r = config.RunInput(grad_cellcntr_imethd=1, lvortex=True, dt=0.005, rk_order=order, lSSPRK=bool(ssprk), lsteady=bool(steady), cfl_user=0.8)
    cfg = r.to_config()
    (res, _, _), (res_o, _, _), q, q_o = _run_steps(mesh, cfg, 10)
    assert _rel(q, q_o) <= TOL
    assert float((np.abs(res - res_o) / np.abs(res_o)).max()) <= TOL


def test_c5_mms_sweep_ggnb():
    """C5: MMS residual-error table (GGNB, one compute_residual(0) per level), reference source (typo kept,
    src/mms.f90:169) and corrected source; oracle and GPU must print the same table."""
    from fvs2d_b200 import config, meshgen
    rows = []
    for n in (16, 32, 64):
        mesh = meshgen.mms_mesh(n)
        r = config.RunInput(grad_cellcntr_imethd=2, ntstart=0, lvortex=False)
        cfg = r.to_config()
        gpu, orc = _pair(mesh, cfg)
        gpu.initialize_solution()
        orc.initialize_solution()
        for corrected in (False, True):
            l2, li = gpu.test_resid(corrected)
            l2_o, li_o = orc.test_resid(corrected)
            assert np.abs(l2 - l2_o).max() / np.abs(l2_o).max() <= 1e-10
            assert np.abs(li - li_o).max() / np.abs(li_o).max() <= 1e-10
            rows.append((n, corrected, l2))
        gpu.close()
    # the reference's density row does not converge (typo); the corrected one does
    typo = [r[2][0] for r in rows if not r[1]]
    fixed = [r[2][0] for r in rows if r[1]]
    assert typo[-1] > 0.5 * typo[0]
    assert fixed[-1] < 0.6 * fixed[0]


def test_determinism_bitwise():
    """no floating-point atomics: two runs give bit-identical states and logs."""
    from fvs2d_b200 import config, meshgen, solver
    mesh = meshgen.vortex_mixed_mesh(64)
    cfg = config.RunInput(grad_cellcntr_imethd=1, lvortex=True, dt=0.005).to_config()
    outs = []
    for _ in range(2):
        gpu = solver.Fvs2dGpu(cfg, device=0)
        gpu.set_mesh(mesh)
        gpu.initialize_solution()
        res, ve, _ = gpu.time_integration(0.0, 7)
        outs.append((gpu.get_state().copy(), res.copy(), ve.copy()))
        gpu.close()
    for a, b in zip(outs[0], outs[1]):
        assert np.array_equal(a, b)


def test_freestream_preservation_and_restart():
    """uniform state + freestream/Dirichlet-constant boundaries keeps resid at round-off; set_state/get_state
    round-trips bit-exactly and a split run (restart path, ntstart>1) equals a straight run."""
    from fvs2d_b200 import config, meshgen, solver
    mesh = meshgen.make_mesh(32, 16, bc_type="freestream")
    cfg = config.RunInput(grad_cellcntr_imethd=1, lvortex=False, dt=0.005, mach_inf=0.5).to_config()
    gpu = solver.Fvs2dGpu(cfg, device=0)
    gpu.set_mesh(mesh)
    gpu.initialize_solution()
    q0 = gpu.get_state().copy()
    resid = gpu.compute_residual(0.0)
    assert np.abs(resid).max() <= 1e-11
    gpu.set_state(q0)
    assert np.array_equal(gpu.get_state(), q0)
    gpu.close()
    # split run on a vortex
    mesh = meshgen.vortex_tri_mesh(40)
    cfg = config.RunInput(grad_cellcntr_imethd=1, lvortex=True, dt=0.005).to_config()
    g1 = solver.Fvs2dGpu(cfg, device=0); g1.set_mesh(mesh); g1.initialize_solution()
    g1.time_integration(0.0, 6)
    qa = g1.get_state().copy()
    g1.initialize_solution()
    g1.time_integration(0.0, 3)
    mid = g1.get_state().copy()
    g1.set_state(mid)
    g1.time_integration(3 * 0.005, 3)
    qb = g1.get_state().copy()
    g1.close()
    assert np.array_equal(qa, qb)


def test_full_size_properties_c3():
    """C3 at full size (4.0 M triangles, GGCB): size-independent checks -- discrete conservation
    (sum vol*dq over cells = boundary flux only -> total mass change tiny for a vortex far from the
    boundary), bitwise determinism of the logs, and log_res equal to the norm recomputed from states."""
    from fvs2d_b200 import capi, config, meshgen, solver
    mesh = meshgen.vortex_tri_mesh(2000)
    # vortex in the domain centre: its tail at the inflow boundary (exp(-50)) carries no measurable momentum in
    cfg = config.RunInput(grad_cellcntr_imethd=1, lvortex=True, dt=0.002, vortex_pos=(10.0, 5.0)).to_config()
    gpu = solver.Fvs2dGpu(cfg, device=0)
    gpu.set_mesh(mesh)
    gpu.initialize_solution()
    vol = capi.mesh_array("vol")
    q0 = gpu.get_state().copy()
    res, ve, _ = gpu.time_integration(0.0, 3)
    q1 = gpu.get_state().copy()
    res2, _, _ = gpu.time_integration(3 * 0.002, 1)
    q2 = gpu.get_state()
    # log_res of the 4th step equals sqrt(sum((q2-q1)^2)/nc)
    ref = np.sqrt(((q2 - q1) ** 2).sum(axis=0) / mesh.ncells)
    assert np.abs(res2[0] - ref).max() / ref.max() <= 1e-12
    # conservation: interior fluxes cancel exactly up to summation round-off
    dm = (vol[:, None] * (q1 - q0)).sum(axis=0)
    tot = (vol[:, None] * np.abs(q0)).sum(axis=0)
    assert np.all(np.abs(dm) / tot <= 1e-9)
    # vortex error stays at the discretisation level (reference diagnostic, src/mms.f90:357-361)
    assert ve[-1, 3] < 1e-3
    gpu.close()
