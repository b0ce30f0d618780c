"""GPU: the one-kernel-per-stage schedule (k_stage_fused; option "fuse": -1 automatic = default, 2 a single launch per
stage, 0 off) gives the bits of the two-pass path and agrees with the CPU oracle; schemes it does not cover (limiter,
kappa != 0, first order) run the two-pass path whatever the option says."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(gpu, fuse, nsteps):
    gpu.set_option("fuse", fuse)
    gpu.initialize_solution()
    res, ve, vxy = gpu.time_integration(0.0, nsteps)
    return gpu.get_state().copy(), res, ve, gpu.last_timing()["launches"]


@pytest.mark.parametrize("case", ["tri-ggcb-rk4", "mixed-lsqfn-ssprk", "mixed-ggcb-rk4-steady", "mixed-lsqnn-rk4"])
def test_fused_bitwise_and_oracle(case):
    from fvs2d_b200 import config, meshgen, solver
    from oracle.oracle import Oracle
    if case == "tri-ggcb-rk4":
        mesh, kw = meshgen.vortex_tri_mesh(44), dict(grad_cellcntr_imethd=1, lvortex=True, dt=0.01)
    elif case == "mixed-lsqfn-ssprk":
        mesh, kw = meshgen.vortex_mixed_mesh(36), dict(grad_cellcntr_imethd=3, grad_cellcntr_lsq_nghbr="fn", lvortex=True, dt=0.01,
                                                       rk_order=2, lSSPRK=True)
    elif case == "mixed-ggcb-rk4-steady":
        mesh, kw = meshgen.vortex_mixed_mesh(64), dict(grad_cellcntr_imethd=1, lvortex=True, dt=0.01, lsteady=True, cfl_user=0.8)
    else:
        mesh, kw = meshgen.vortex_mixed_mesh(32), dict(grad_cellcntr_imethd=3, grad_cellcntr_lsq_nghbr="nn", lvortex=True, dt=0.01)
    cfg = config.RunInput(**kw).to_config()
    n = 8
    gpu = solver.Fvs2dGpu(cfg, device=0)
    gpu.set_mesh(mesh)
    q0, r0, v0, l0 = _run(gpu, 0, n)
    for fuse in (2, -1):
        q, r, v, l = _run(gpu, fuse, n)
        assert np.array_equal(q, q0) and np.allclose(r, r0, rtol=1e-13, atol=0.0), f"fuse={fuse}"
        assert np.allclose(v, v0, rtol=1e-12, atol=0.0)
        assert l <= l0 and (fuse != 2 or l < l0), "one launch per stage (two where the tiles are split by shared-memory need) instead of two passes"
    gpu.close()
    orc = Oracle(mesh, cfg)
    orc.initialize_solution()
    r_o, _, _ = orc.time_integration(0.0, n)
    scale = np.abs(orc.cvar).max(axis=0)
    assert float((np.abs(q0 - orc.cvar) / scale).max()) <= 1e-10      # tolerance of BASELINE.json's north_star
    assert float((np.abs(r0 - r_o) / np.abs(r_o)).max()) <= 1e-10


def test_fuse_is_ignored_where_it_does_not_apply():
    """limiter / kappa != 0 / first order: the option is accepted and the two-pass path runs."""
    from fvs2d_b200 import config, meshgen, solver
    mesh = meshgen.vortex_tri_mesh(24)
    cfg = config.RunInput(grad_cellcntr_imethd=3, grad_cellcntr_lsq_nghbr="nn", grad_limiter_imethd=1, lvortex=True, dt=0.01).to_config()
    gpu = solver.Fvs2dGpu(cfg, device=0)
    gpu.set_mesh(mesh)
    q0, r0, _, l0 = _run(gpu, 0, 4)
    q1, r1, _, l1 = _run(gpu, 2, 4)
    assert np.array_equal(q0, q1) and np.array_equal(r0, r1) and l0 == l1
    gpu.close()
