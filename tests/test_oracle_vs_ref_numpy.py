"""CPU: the C oracle against fixtures produced by an INDEPENDENT numpy transcription of the Fortran sources
(tests/golden/ref_numpy.py -> tests/golden/ref_*.npz, generator tests/golden/make_ref_fixtures.py).

What this pins, and what it does not: the reference holds no numeric goldens and cannot be built in this image (no Fortran
compiler), so nothing here is an output of the reference itself -- parity stays "unpinned by the reference".  The fixtures
were written from src/*.f90 without looking at oracle/, and use different formulations on purpose (edge-based Green-Gauss,
node-interpolated GGNB, pseudo-inverse least squares, sorted node-pair edge detection, scipy's kd-tree): two independent
transcriptions agreeing to round-off on fields, gradients, limiters, logs and MMS rows is the strongest pin available.

Tolerances: 1e-12 relative to the column / field maximum (round-off of different summation orders and of lstsq-vs-normal
equations, amplified over <= 100 steps); 1e-10 on the limited NACA fields after 8 steps (see the test for why not 20)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN


def _load(name):
    z = np.load(os.path.join(GOLDEN, name))
    return z, json.loads(str(z["case"]))


def _run_input(c):
    from fvs2d_b200 import config
    return config.RunInput(gamma=c["gamma"], dt=c["dt"], cfl_user=c["cfl_user"], umuscl_cst=c["umuscl_cst"],
                           grad_cellcntr_lsq_pow=c["lsq_pow"], grad_cellcntr_imethd=c["grad_method"], grad_cellcntr_lsq_nghbr=c["lsq_stencil"],
                           grad_limiter_imethd=c["limiter"], face_reconst_imethd=c["recon"], rk_order=c["rk_order"], lSSPRK=bool(c["ssprk"]),
                           lsteady=bool(c["steady"]), lvortex=bool(c["lvortex"]), ntstart=c["ntstart"], mach_inf=c["pvar_inf"][1],
                           vortex_pos=tuple(c["vortex_pos"]), vortex_kappa=c["vortex_kappa"], vortex_inf=tuple(c["vortex_inf"]))


def _oracle(mesh, c):
    from oracle.oracle import Oracle
    return Oracle(mesh, _run_input(c).to_config())


def _rel(a, b, axis=0):
    """max |a-b| relative to the largest |b| of each column"""
    return float((np.abs(a - b) / np.maximum(np.abs(b).max(axis=axis, keepdims=True), 1e-300)).max())


def test_c1_vortex_100_steps(vortex_mesh):
    z, case = _load("ref_c1_vortex.npz")
    o = _oracle(vortex_mesh, case["cfg"])
    o.initialize_solution()
    dt, done = case["cfg"]["dt"], 0
    res, ve, vxy = [], [], []
    for upto in case["steps"]:
        r, e, xy = o.time_integration(done * dt, upto - done)
        res.append(r); ve.append(e); vxy.append(xy)
        done = upto
        assert _rel(o.cvar, z[f"cvar_{upto}"]) <= 1e-12, f"state after {upto} steps"
    res, ve, vxy = np.concatenate(res), np.concatenate(ve), np.concatenate(vxy)
    assert np.abs(res / z["log_res"] - 1.0).max() <= 1e-10
    assert np.abs(ve / z["vortex_err"] - 1.0).max() <= 1e-9       # errors ~1e-5 of O(1) fields: 1e-9 relative = 1e-14 absolute
    np.testing.assert_array_equal(vxy, z["vortex_xy"])            # the same cell attains the largest density error at every step
    assert case["lsq_verify"] <= 1e-10 and o.scalars()["lsq_verify_err"] <= 1e-10


@pytest.mark.parametrize("lim", [1, 0])
def test_c2_naca_8_and_20_steps(naca_mesh, lim):
    """Unlimited: 20 steps to 1e-12.  Venkatakrishnan: the impulsive start is ill-conditioned -- from step 9 on a last-bit
    difference grows ~10x per step (measured between the two transcriptions: 7e-12 at step 8, 2e-9 at step 10, 2e-2 at
    step 20, confined to ~1 % of the cells) -- so the tight comparison is made after 8 steps and the 20-step state is only
    required to agree on 97 % of the sampled cells."""
    z, case = _load("ref_c2_naca.npz")
    c = dict(case["cfg"], limiter=lim)
    o = _oracle(naca_mesh, c)
    o.initialize_solution()
    st = case["stride"]
    r8, _, _ = o.time_integration(0.0, 8)
    assert _rel(o.cvar[::st], z[f"cvar8_lim{lim}"]) <= (1e-10 if lim else 1e-12)
    assert np.abs(r8 / z[f"log_res_lim{lim}"][:8] - 1.0).max() <= 1e-10
    o.compute_residual(0.0)
    assert np.abs(o.array("phi_lim")[::st] - z[f"phi8_lim{lim}"]).max() <= 1e-8
    r20, _, _ = o.time_integration(0.0, 12)
    d = (np.abs(o.cvar[::st] - z[f"cvar_lim{lim}"]) / np.abs(z[f"cvar_lim{lim}"]).max(axis=0)).max(axis=1)
    if lim:
        assert np.quantile(d, 0.97) <= 1e-8 and d.max() <= 0.2
    else:
        assert d.max() <= 1e-12 and np.abs(r20 / z[f"log_res_lim{lim}"][8:] - 1.0).max() <= 1e-10


def test_single_residuals_every_scheme():
    from fvs2d_b200 import meshgen
    z, case = _load("ref_resid_mixed.npz")
    mesh = meshgen.vortex_mixed_mesh(24)
    for k, kw in case["cases"].items():
        o = _oracle(mesh, dict(case["base"], **kw))
        o.initialize_solution()
        o.set_state(z[f"{k}_cvar0"])
        R = o.compute_residual(case["time"])
        nc = mesh.ncells
        g = o.array("grad").reshape(2, nc, 4)         # Fortran grad(ivar, ic, idim)
        # van Albada as coded returns |phi| up to 1e3 (src/gradient_limiter.f90:127-128): same values, looser absolute scale
        tol = 1e-9 if k == "lsqnn_albada" else 1e-12
        assert _rel(R, z[f"{k}_resid"]) <= tol, k
        assert _rel(o.array("ws_nrml"), z[f"{k}_ws"]) <= tol, k
        assert _rel(g[0], z[f"{k}_gx"]) <= 1e-12 and _rel(g[1], z[f"{k}_gy"]) <= 1e-12, k
        assert _rel(o.array("phi_lim"), z[f"{k}_phi"]) <= tol, k


def test_integrators():
    from fvs2d_b200 import meshgen
    z, case = _load("ref_integrators.npz")
    mesh = meshgen.vortex_mixed_mesh(16)
    for k, kw in case["cases"].items():
        o = _oracle(mesh, dict(case["base"], **kw))
        o.initialize_solution()
        r, e, _ = o.time_integration(0.0, case["steps"])
        assert _rel(o.cvar, z[f"{k}_cvar"]) <= 1e-12, k
        assert np.abs(r / z[f"{k}_log_res"] - 1.0).max() <= 1e-10, k
        assert np.abs(e / z[f"{k}_verr"] - 1.0).max() <= 1e-9, k


def test_mms_rows_and_observed_order():
    """C5 (src/test.f90:481-519): the error_resid.plt rows of the oracle equal the fixture's for n = 16..128, with the
    reference's source (src/mms.f90:169 typo) and the corrected one.  What the table shows (and what "the reference's order"
    therefore is on SURVEY C5's jittered split-quad meshes, whose irregularity does not vanish with h): the residual
    (truncation) error of the second-order scheme converges at ~0.85-0.97 between n = 16 and 32 and then stalls towards
    order 0 (0.2-0.3 between n = 64 and 128) -- the classical O(1) truncation error of finite volumes on irregular meshes;
    with the typo the continuity row does not converge at all (order -0.01)."""
    from fvs2d_b200 import meshgen
    z, case = _load("ref_mms.npz")
    rows = z["rows"]
    for i, n in enumerate(case["n"]):
        o = _oracle(meshgen.mms_mesh(n), case["cfg"])
        o.initialize_solution()
        assert abs(o.scalars()["heff1"] / z["heff"][i] - 1.0) <= 1e-13
        for j, corr in enumerate((False, True)):
            l2, li = o.test_resid(corr)
            assert np.abs(l2 / rows[i, j, 0] - 1.0).max() <= 1e-10, (n, corr)
            assert np.abs(li / rows[i, j, 1] - 1.0).max() <= 1e-10, (n, corr)
    order = np.log(rows[:-1, :, 0] / rows[1:, :, 0]) / np.log(z["heff"][:-1] / z["heff"][1:])[:, None, None]
    assert (order[0, 1] > 0.8).all() and (order[-1, 1] > 0.1).all() and (order[-1, 1] < 0.4).all(), order
    assert (np.abs(order[:, 0, 0]) < 0.05).all(), order  # typo kept: the continuity residual error never converges
