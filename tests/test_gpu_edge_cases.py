"""GPU: edge cases of the C-ABI -- error behaviour (the reference prints and stops; here: non-zero return +
fvs2d_gpu_last_error), ragged / tiny inputs (meshes smaller than one 128-cell tile, partial last tiles), zero-length
calls, and the fallback kernel."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((np.abs(a - b) / np.maximum(np.abs(b).max(axis=0), 1e-300)).max())


def test_call_order_and_scheme_errors():
    from fvs2d_b200 import capi, config, meshgen, solver
    L = capi.lib()
    L.fvs2d_gpu_finalize()
    # no context yet
    assert L.fvs2d_gpu_time_integration(0.0, 1, None, None, None) != 0
    assert b"no state" in L.fvs2d_gpu_last_error()
    # limiter needs the LSQ stencil (src/gradient_limiter.f90:54-58): GGCB + Venkatakrishnan is rejected at init
    bad = config.RunInput(grad_cellcntr_imethd=1, grad_limiter_imethd=1).to_config()
    assert L.fvs2d_gpu_init(ctypes.byref(bad), 0) != 0
    assert b"limiter" in L.fvs2d_gpu_last_error()
    # SSPRK only for (4 stages, order 2) (src/runge_kutta.f90:56); only 4 stages are coded (:36)
    for kw in (dict(lSSPRK=True, rk_order=4), dict(rk_nstages=3)):
        cfg = config.RunInput(**kw).to_config()
        assert L.fvs2d_gpu_init(ctypes.byref(cfg), 0) != 0
    # a solid_wall boundary is "not implemented yet" in the reference (src/residual.f90:206-208)
    gpu = solver.Fvs2dGpu(config.RunInput(grad_cellcntr_imethd=1).to_config(), device=0)
    mesh = meshgen.make_mesh(8, 8, bc_type="solid_wall")
    with pytest.raises(capi.Fvs2dError, match="not implemented"):
        gpu.set_mesh(mesh)
    # state calls before a mesh / state exist
    with pytest.raises(capi.Fvs2dError):
        gpu.get_state(np.zeros((4, 4)))
    mesh = meshgen.make_mesh(8, 8)
    gpu.set_mesh(mesh)
    with pytest.raises(capi.Fvs2dError, match="no state"):
        gpu.time_integration(0.0, 1)
    # a cell with two boundary edges breaks the LSQ-fn stencil rule only when > 2 (src/gradient_lsq.f90:128-131); a
    # .bc file that lists too few cells is caught like in grid_data (src/grid_procs.f90:722-728)
    import copy
    m2 = copy.deepcopy(mesh)
    m2.bndry_cell = [m2.bndry_cell[0][:-2]]
    with pytest.raises(capi.Fvs2dError, match="boundary cells"):
        gpu.set_mesh(m2)
    gpu.close()


@pytest.mark.parametrize("nx,ny,band", [(2, 2, None), (3, 2, None), (5, 3, None), (6, 4, (2, 4)), (13, 9, (4, 9)), (17, 11, None)])
@pytest.mark.parametrize("tile", [2, 0])
def test_tiny_and_ragged_meshes(nx, ny, band, tile):
    """meshes of 8 ... 374 cells: less than one tile, exactly-not-a-multiple of 32 / 128, mixed; pipeline and
    fallback kernels; 6 steps against the oracle."""
    from fvs2d_b200 import config, meshgen, solver
    from oracle.oracle import Oracle
    mesh = meshgen.make_mesh(nx, ny, 20.0, 10.0, band)
    cfg = config.RunInput(grad_cellcntr_imethd=1, lvortex=True, dt=0.02).to_config()
    gpu = solver.Fvs2dGpu(cfg, device=0)
    gpu.set_option("tile", tile)
    gpu.set_mesh(mesh)
    gpu.initialize_solution()
    orc = Oracle(mesh, cfg)
    orc.initialize_solution()
    res, ve, vxy = gpu.time_integration(0.0, 6)
    res_o, ve_o, vxy_o = orc.time_integration(0.0, 6)
    assert _rel(gpu.get_state(), orc.cvar) <= 1e-10
    assert float((np.abs(res - res_o) / np.abs(res_o)).max()) <= 1e-10
    if orc.sizes()["ncells_intr"] > 0:
        assert float((np.abs(ve - ve_o) / np.maximum(np.abs(ve_o), 1e-300)).max()) <= 1e-8
        assert np.abs(vxy - vxy_o).max() == 0.0
    gpu.close()


def test_zero_steps_and_repeated_calls():
    """nsub = 0 is a no-op; 1+2+4 steps in three calls == 7 steps in one (eager first step + graph replay)."""
    from fvs2d_b200 import config, meshgen, solver
    mesh = meshgen.vortex_mixed_mesh(24)
    cfg = config.RunInput(grad_cellcntr_imethd=3, grad_cellcntr_lsq_nghbr="nn", grad_limiter_imethd=1, lvortex=True, dt=0.01).to_config()
    gpu = solver.Fvs2dGpu(cfg, device=0)
    gpu.set_mesh(mesh)
    gpu.initialize_solution()
    q0 = gpu.get_state().copy()
    res, ve, _ = gpu.time_integration(0.0, 0)
    assert res.shape == (0, 4) and np.array_equal(gpu.get_state(), q0)
    r7, v7, _ = gpu.time_integration(0.0, 7)
    q7 = gpu.get_state().copy()
    gpu.set_state(q0)
    parts = []
    t = 0.0
    for n in (1, 2, 4):
        r, v, _ = gpu.time_integration(t, n)
        parts.append(r)
        t += n * 0.01
    assert np.array_equal(gpu.get_state(), q7)
    assert np.array_equal(np.concatenate(parts), r7)
    # graph replay off gives the same bits
    gpu.set_option("graph", 0)
    gpu.set_state(q0)
    r7b, _, _ = gpu.time_integration(0.0, 7)
    assert np.array_equal(r7b, r7) and np.array_equal(gpu.get_state(), q7)
    gpu.close()


def test_fallback_kernel_matches_pipeline_bitwise():
    """the direct-gather kernel (any numbering) and the shared-memory pipeline evaluate the same faces in the same
    order: identical bits."""
    from fvs2d_b200 import config, meshgen, solver
    mesh = meshgen.vortex_mixed_mesh(40)
    cfg = config.RunInput(grad_cellcntr_imethd=2, face_reconst_imethd=3, umuscl_cst=1.0 / 3.0, lvortex=True, dt=0.01).to_config()
    out = []
    for tile in (2, 0):
        gpu = solver.Fvs2dGpu(cfg, device=0)
        gpu.set_option("tile", tile)
        gpu.set_mesh(mesh)
        gpu.initialize_solution()
        r, v, _ = gpu.time_integration(0.0, 5)
        out.append((gpu.get_state().copy(), r, v))
        gpu.close()
    for a, b in zip(*out):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("stencil,limiter", [("fn", 0), ("nn", 1)])
def test_set_lsq_uses_the_callers_table(stencil, limiter):
    """fvs2d_gpu_set_lsq: the reference's public lsq(:) table -- here the oracle's, with its own kd-tree
    choices for the boundary stencils -- replaces the library's; residual, gradients and limiter equal the oracle's, and a
    deliberately different table (every weight doubled, coefficient halved: the same operator) gives the same numbers."""
    from fvs2d_b200 import config, meshgen, solver
    from oracle.oracle import Oracle
    mesh = meshgen.vortex_mixed_mesh(40)
    cfg = config.RunInput(grad_cellcntr_imethd=3, grad_cellcntr_lsq_nghbr=stencil, grad_cellcntr_lsq_pow=1.0, grad_limiter_imethd=limiter,
                          lvortex=True, dt=0.005).to_config()
    orc = Oracle(mesh, cfg)
    orc.initialize_solution()
    r_o = orc.compute_residual(0.2).copy()
    g_o, ph_o = orc.array("grad").reshape(2, -1, 4), orc.array("phi_lim")
    ptr, cell, w, coef = orc.array("lsq_ptr"), orc.array("lsq_cell"), orc.array("lsq_w"), orc.array("lsq_coef").reshape(-1, 2)
    gpu = solver.Fvs2dGpu(cfg, device=0)
    gpu.set_mesh(mesh)
    gpu.initialize_solution()
    r_lib = gpu.compute_residual(0.2)
    for ww, cc in ((w, coef), (2.0 * w, 0.5 * coef)):
        gpu.set_lsq(ptr, cell, ww, cc)
        gpu.initialize_solution()
        r = gpu.compute_residual(0.2)
        _, gr, ph = gpu.get_aux()
        scale = np.abs(r_o).max(axis=0)
        assert (np.abs(r - r_o) / scale).max() <= 1e-11
        assert np.abs(gr - g_o).max() / np.abs(g_o).max() <= 1e-12
        assert np.abs(ph - ph_o).max() <= 1e-9
        assert (np.abs(r - r_lib) / scale).max() <= 1e-11
    res, _, _ = gpu.time_integration(0.0, 3)       # the time loop runs on the supplied operator as well
    res_o, _, _ = orc.time_integration(0.0, 3)
    assert float((np.abs(res - res_o) / np.abs(res_o)).max()) <= 1e-9
    # a table that is not linearly exact is rejected like the reference's grad_lsq_verify does (src/gradient_lsq.f90:490-529)
    from fvs2d_b200 import capi
    with pytest.raises(capi.Fvs2dError, match="LSQ coefficients"):
        gpu.set_lsq(ptr, cell, w, 1.1 * coef)
    gpu.close()


def test_device_hilbert_sort_equals_host_sort(monkeypatch):
    """set-up: the Hilbert keys and their stable radix sort run on the device in fvs2d_gpu_set_mesh; the
    permutation must be the host's (fvs2d_host_build / FVS2D_HOST_SORT=1), bit for bit, because every rank derives its
    partition from it."""
    from fvs2d_b200 import capi, config, meshgen, solver
    mesh = meshgen.vortex_mixed_mesh(96)
    cfg = config.RunInput(grad_cellcntr_imethd=1, lvortex=True, dt=0.005).to_config()
    gpu = solver.Fvs2dGpu(cfg, device=0)
    gpu.set_mesh(mesh)
    dev = capi.mesh_array("perm").copy()
    monkeypatch.setenv("FVS2D_HOST_SORT", "1")
    gpu.set_mesh(mesh)
    host = capi.mesh_array("perm").copy()
    gpu.close()
    assert np.array_equal(np.sort(dev), np.arange(mesh.ncells)) and np.array_equal(dev, host)


@pytest.mark.parametrize("case", ["c1-vortex-lsqfn", "c2-naca-venkat-steady", "naca-ggcb-rk4", "mixed-ggnb-umuscl", "tri-first-order", "mixed-lsqnn-barth"])
def test_two_threads_per_cell_kernels_match_one_thread_kernels(case, vortex_mesh, naca_mesh):
    """Small meshes run both passes with two threads per cell (option "pair", automatic up to 1024 cells per SM): the state,
    gradients and limiter must be bitwise those of the one-thread-per-cell kernels; log_res / vortex errors agree to 1e-13
    (the norm partials are grouped per 64 cells instead of per persistent CTA)."""
    from fvs2d_b200 import config, meshgen, solver
    from conftest import run_input
    if case == "c1-vortex-lsqfn":
        mesh, run = vortex_mesh, run_input("vortex")
    elif case == "c2-naca-venkat-steady":
        mesh, run = naca_mesh, run_input("naca")
        run.grad_limiter_imethd = 1
    elif case == "naca-ggcb-rk4":
        mesh, run = naca_mesh, config.RunInput(grad_cellcntr_imethd=1, dt=1e-4, mach_inf=0.5)
    elif case == "mixed-ggnb-umuscl":
        mesh, run = meshgen.vortex_mixed_mesh(47), config.RunInput(grad_cellcntr_imethd=2, face_reconst_imethd=3, umuscl_cst=1.0 / 3.0, lvortex=True, dt=0.005)
    elif case == "mixed-lsqnn-barth":
        mesh, run = meshgen.vortex_mixed_mesh(40), config.RunInput(grad_cellcntr_imethd=3, grad_cellcntr_lsq_nghbr="nn", grad_limiter_imethd=2, lvortex=True, dt=0.005)
    else:
        mesh, run = meshgen.vortex_tri_mesh(40), config.RunInput(grad_cellcntr_imethd=1, face_reconst_imethd=1, lvortex=True, dt=0.005)
    cfg = run.to_config()
    gpu = solver.Fvs2dGpu(cfg, device=0)
    gpu.set_mesh(mesh)
    out = {}
    for pair in (0, 1):
        gpu.set_option("pair", pair)
        gpu.set_option("fuse", 0)
        gpu.initialize_solution()
        r1, v1, x1 = gpu.time_integration(0.0, 7)
        r2, v2, x2 = gpu.time_integration(7 * run.dt, 5)          # a second call: graph replay with the same kernels
        resid = gpu.compute_residual(12 * run.dt)
        aux = gpu.get_aux()
        out[pair] = (gpu.get_state().copy(), np.concatenate([r1, r2]), None if v1 is None else np.concatenate([v1, v2]),
                     None if x1 is None else np.concatenate([x1, x2]), resid, aux)
    gpu.close()
    q0, r0, v0, x0, res0, aux0 = out[0]
    q1, r1, v1, x1, res1, aux1 = out[1]
    assert np.array_equal(q0, q1) and np.array_equal(res0, res1, equal_nan=True)
    for a, b in zip(aux0, aux1):
        assert np.array_equal(a, b, equal_nan=True)
    assert np.allclose(r1, r0, rtol=1e-13, atol=0.0)
    if v0 is not None:
        assert np.allclose(v1, v0, rtol=1e-12, atol=0.0) and np.array_equal(x0, x1)


@pytest.mark.parametrize("case", ["c1-vortex-lsqfn", "c2-naca-venkat-steady", "mixed-ggcb-fused-264k"])
def test_programmatic_dependent_launch_changes_nothing(case, vortex_mesh, naca_mesh):
    """option "pdl" (default on, one GPU): the kernels of a step are launched with programmatic stream serialisation and
    wait for their predecessor at their first instruction -- state and logs must be bitwise those of plain launches, on the
    launch-bound examples and on a mesh whose stage kernels run several waves of tiles."""
    from fvs2d_b200 import config, meshgen, solver
    from conftest import run_input
    if case == "c1-vortex-lsqfn":
        mesh, run, n = vortex_mesh, run_input("vortex"), 40
    elif case == "c2-naca-venkat-steady":
        mesh, run, n = naca_mesh, run_input("naca"), 12
        run.grad_limiter_imethd = 1
    else:
        mesh, run, n = meshgen.vortex_mixed_mesh(420), config.RunInput(grad_cellcntr_imethd=1, lvortex=True, dt=0.002), 12
    cfg = run.to_config()
    gpu = solver.Fvs2dGpu(cfg, device=0)
    gpu.set_mesh(mesh)
    out = {}
    for pdl in (0, 1):
        gpu.set_option("pdl", pdl)
        gpu.initialize_solution()
        r1, v1, _ = gpu.time_integration(0.0, n)
        r2, v2, _ = gpu.time_integration(n * run.dt, n)
        out[pdl] = (gpu.get_state().copy(), np.concatenate([r1, r2]), None if v1 is None else np.concatenate([v1, v2]))
    gpu.close()
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    if out[0][2] is not None:
        assert np.array_equal(out[0][2], out[1][2])
