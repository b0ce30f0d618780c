"""GPU: the output path on the device -- fvs2d_gpu_interpolate_cell2node and
fvs2d_gpu_wall_values through the C-ABI against the oracle's restatement of write_inst_ios (src/io.f90:122-150,
src/interpolation.f90:62-123) and write_inst_cp_un (src/io.f90:340-449) on the same state.

Tolerance: 1e-12 relative to the variable's magnitude -- the interpolation sums the node's cells in the reference's
order, so only FMA contraction and the 1e-10-bounded state difference after the steps separate the two.
"""
import numpy as np
import pytest

from conftest import run_input

pytestmark = pytest.mark.gpu


def _pair(mesh, cfg):
    from fvs2d_b200 import solver
    from oracle.oracle import Oracle
    gpu = solver.Fvs2dGpu(cfg, device=0)
    gpu.set_mesh(mesh)
    orc = Oracle(mesh, cfg)
    gpu.initialize_solution()
    orc.initialize_solution()
    return gpu, orc


def _relv(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.mark.parametrize("case", ["vortex", "mixed"])
def test_interpolate_cell2node_matches_oracle(case, vortex_mesh):
    from fvs2d_b200 import config, meshgen
    if case == "vortex":
        mesh, r = vortex_mesh, run_input("vortex")
    else:
        mesh = meshgen.make_mesh(36, 18, 20.0, 10.0, (9, 27))
        r = config.RunInput(grad_cellcntr_imethd=1, lvortex=True, dt=0.01)
    gpu, orc = _pair(mesh, r.to_config())
    for nsteps in (0, 5):
        if nsteps:
            gpu.time_integration(0.0, nsteps)
            orc.time_integration(0.0, nsteps)
        fn = gpu.interpolate_cell2node((1, 1, 1, 1))
        assert fn.shape == (4, mesh.nnodes)
        for v in range(4):
            assert _relv(fn[v], orc.interpolate_cell2node(v)) <= 1e-12, (nsteps, v)
        # a subset comes back packed in variable order (lw_inst of fvs2d.input line 17)
        sub = gpu.interpolate_cell2node((0, 1, 0, 1))
        assert sub.shape == (2, mesh.nnodes)
        np.testing.assert_array_equal(sub[0], fn[1])
        np.testing.assert_array_equal(sub[1], fn[3])
        assert gpu.interpolate_cell2node((0, 0, 0, 0)).shape == (0, mesh.nnodes)
    # the state itself is untouched by the output path
    assert _relv(gpu.get_state(), orc.cvar) <= 1e-10
    gpu.close()


@pytest.mark.parametrize("limiter", [0, 1])
def test_wall_values_naca(naca_mesh, limiter):
    """C2: slip-wall pressure / normal velocity after 10 steady SSPRK steps (LSQ-nn; the wall gradient is unlimited also
    when the run itself is limited).  The limited transonic trajectory amplifies last-bit differences tenfold per step
    (tests/test_gpu_parity.py::test_c2_naca_limited), so the oracle evaluates the wall values on the GPU's state:
    single-evaluation parity."""
    r = run_input("naca")
    r.grad_limiter_imethd = limiter
    gpu, orc = _pair(naca_mesh, r.to_config())
    gpu.time_integration(0.0, 10)
    q = gpu.get_state()
    orc.set_state(q)
    ib = naca_mesh.bndry_type.index("slip_wall")
    wg, wo = gpu.wall_values(ib), orc.wall_values(ib)
    assert wg.shape == wo.shape == (256, 4)
    np.testing.assert_allclose(wg[:, 0], wo[:, 0], rtol=1e-15, atol=1e-18)
    assert np.abs(wo[:, 1] - wo[:, 2]).max() > 1e-6            # the flow has developed: wall and cell pressure differ
    for k in (1, 2, 3):
        assert _relv(wg[:, k], wo[:, k]) <= 1e-12, k
    # the other boundary (freestream) works the same way; an out-of-range index is an error
    assert _relv(gpu.wall_values(1 - ib)[:, 1], orc.wall_values(1 - ib)[:, 1]) <= 1e-12
    from fvs2d_b200.capi import Fvs2dError
    with pytest.raises(Fvs2dError):
        gpu.wall_values(2)
    # the output calls leave the state alone, and the time loop continues unaffected (gradients are recomputed
    # every stage): 5 more steps equal 5 steps restarted from the same state (set_state recomputes the primitive
    # variables in another kernel, so allow the last-bit difference and its growth over 5 limited steps)
    np.testing.assert_array_equal(gpu.get_state(), q)
    res_a, _, _ = gpu.time_integration(10 * r.dt, 5)
    qa = gpu.get_state()
    gpu.set_state(q)
    res_b, _, _ = gpu.time_integration(10 * r.dt, 5)
    np.testing.assert_allclose(res_a, res_b, rtol=1e-8)
    assert _relv(qa, gpu.get_state()) <= 1e-9
    gpu.close()


def test_wall_values_first_order_still_uses_the_gradient():
    """gradient_cellcntr_1var (src/gradient.f90:74-96) evaluates the selected gradient scheme even when the
    reconstruction is first order (compute_gradient_cellcntr returns early, :49)."""
    from fvs2d_b200 import config, meshgen
    mesh = meshgen.make_mesh(24, 12, 20.0, 10.0, (6, 18), bc_type="slip_wall")
    r = config.RunInput(grad_cellcntr_imethd=1, face_reconst_imethd=1, lvortex=False, dt=0.005, mach_inf=0.3)
    gpu, orc = _pair(mesh, r.to_config())
    # a non-uniform state so the gradients are not zero
    rng = np.random.default_rng(7)
    q = orc.cvar * (1.0 + 0.05 * rng.standard_normal(orc.cvar.shape))
    gpu.set_state(q)
    orc.set_state(q)
    wg, wo = gpu.wall_values(0), orc.wall_values(0)
    assert np.abs(wo[:, 1] - wo[:, 2]).max() > 1e-4          # the extrapolation really uses a gradient
    for k in (1, 2, 3):
        assert _relv(wg[:, k], wo[:, k]) <= 1e-12, k
    gpu.close()
