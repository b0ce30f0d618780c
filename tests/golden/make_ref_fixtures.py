"""Generates tests/golden/ref_*.npz from the independent numpy restatement (tests/golden/ref_numpy.py).

    python tests/golden/make_ref_fixtures.py            (numpy + scipy only; ~2 min; no GPU, no oracle, no product code
                                                         except the mesh generator / fixture meshes, which are input data)

Every fixture holds the inputs needed to repeat the run (a case dictionary) and the outputs of ref_numpy:
  ref_c1_vortex.npz        C1: the shipped isentropic-vortex example (LSQ-fn, RK4), 100 steps: cvar at steps 1, 10, 100,
                           log_res[100,4], vortex_err[100,14], vortex_xy[100,2]
  ref_c2_naca.npz          C2: NACA 0012 o-grid, LSQ-nn, SSPRK steady CFL 1.25, Venkatakrishnan and unlimited: every 16th cell
                           of cvar after 8 and 20 steps, phi_lim of the state after 8 steps, log_res[20,4].  (The limited
                           impulsive start is ill-conditioned: a last-bit difference grows ~10x per step from step 9 on --
                           two correct implementations agree to 1e-11 at step 8 and to 1e-2 at step 20; the unlimited run
                           agrees to 2e-15 at step 20.)
  ref_resid_mixed.npz      one compute_residual on a 24x12 mixed tri/quad vortex mesh for GGCB, GGNB, LSQ-fn, LSQ-nn,
                           LSQ-nn + each limiter, UMUSCL kappa=1/3, first order: resid, ws_nrml, grad, phi
  ref_integrators.npz      RK order 1-4, SSPRK, steady variants, 6 steps on a 16x8 mixed mesh: cvar, log_res
  ref_mms.npz              C5: test_resid rows (typo kept / corrected) for GGNB on n = 16, 32, 64, 128
These are NOT outputs of the reference (no Fortran compiler in this image): they pin the C oracle against a second,
independently written transcription of the same Fortran sources.  tests/test_oracle_pins.py compares oracle == fixture.
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import ref_numpy as rn  # noqa: E402
from fvs2d_b200 import meshgen, meshio  # noqa: E402  (mesh generator + fixture meshes = input data)

BASE = dict(gamma=1.4, dt=0.01, cfl_user=1.25, umuscl_cst=0.0, lsq_pow=0.0, grad_method=3, lsq_stencil="fn", limiter=0, recon=2,
            rk_order=4, ssprk=0, steady=0, lvortex=1, ntstart=1, pvar_inf=[1.0, 0.8, 0.0, 1.0 / 1.4],
            vortex_pos=[5.0, 5.0], vortex_kappa=1.0, vortex_inf=[1.0, 0.2, 0.0, 1.0])


def cfg(**kw):
    c = dict(BASE)
    c.update(kw)
    return c


def ref_mesh(m):
    return rn.RefMesh(m.node_xy, m.tri, m.quad, m.bndry_type, m.bndry_cell)


def save(name, case, **arrays):
    np.savez_compressed(os.path.join(HERE, name), case=json.dumps(case), **arrays)
    print(f"{name}: {os.path.getsize(os.path.join(HERE, name)) / 1e3:.0f} kB", flush=True)


def main():
    t0 = time.time()
    # ---- C1
    vm = meshio.load_npz(os.path.join(HERE, "vortex_mesh.npz"))
    c = cfg()
    s = rn.RefSolver(ref_mesh(vm), c)
    s.initialize_solution()
    snaps, res, ve, vxy, done = {}, [], [], [], 0
    for upto in (1, 10, 100):
        r, e, xy = s.time_integration(done * c["dt"], upto - done)
        res.append(r); ve.append(e); vxy.append(xy)
        done = upto
        snaps[f"cvar_{upto}"] = s.cvar.copy()
    save("ref_c1_vortex.npz", dict(mesh="vortex_mesh.npz", cfg=c, steps=[1, 10, 100], lsq_verify=s.lsq.verify()),
         log_res=np.concatenate(res), vortex_err=np.concatenate(ve), vortex_xy=np.concatenate(vxy), **snaps)
    # ---- C2
    nm = meshio.load_npz(os.path.join(HERE, "naca_mesh.npz"))
    rm = ref_mesh(nm)
    out = {}
    for lim in (1, 0):
        c = cfg(lsq_stencil="nn", limiter=lim, rk_order=2, ssprk=1, steady=1, lvortex=0)
        s = rn.RefSolver(rm, c)
        s.initialize_solution()
        r8, _, _ = s.time_integration(0.0, 8)
        out[f"cvar8_lim{lim}"] = s.cvar[::16].copy()
        s.compute_residual(0.0)
        out[f"phi8_lim{lim}"] = s.phi[::16].copy()
        r20, _, _ = s.time_integration(0.0, 12)
        out[f"cvar_lim{lim}"] = s.cvar[::16].copy()
        out[f"log_res_lim{lim}"] = np.concatenate([r8, r20])
    save("ref_c2_naca.npz", dict(mesh="naca_mesh.npz", cfg=cfg(lsq_stencil="nn", rk_order=2, ssprk=1, steady=1, lvortex=0), steps=20,
                                 stride=16, limiters=[1, 0]), **out)
    # ---- single residuals, every gradient / limiter / reconstruction
    mm = meshgen.vortex_mixed_mesh(24)
    rm = ref_mesh(mm)
    cases = {"ggcb": dict(grad_method=1), "ggnb": dict(grad_method=2), "lsqfn": dict(grad_method=3, lsq_stencil="fn"),
             "lsqnn": dict(grad_method=3, lsq_stencil="nn"), "lsqnn_p1": dict(grad_method=3, lsq_stencil="nn", lsq_pow=1.0),
             "lsqnn_venk": dict(grad_method=3, lsq_stencil="nn", limiter=1), "lsqnn_barth": dict(grad_method=3, lsq_stencil="nn", limiter=2),
             "lsqnn_albada": dict(grad_method=3, lsq_stencil="nn", limiter=3), "ggnb_umuscl": dict(grad_method=2, recon=3, umuscl_cst=1.0 / 3.0),
             "first_order": dict(grad_method=1, recon=1)}
    out = {}
    for k, kw in cases.items():
        s = rn.RefSolver(rm, cfg(**kw))
        s.initialize_solution()
        # off the exact solution so that the limiters are active; van Albada as coded in the reference (phi/(b+eps2),
        # src/gradient_limiter.f90:127-128) returns large negative phi and NaNs at 2 %: it gets 0.5 %
        amp = 0.005 if k == "lsqnn_albada" else 0.02
        s.cvar = s.cvar * (1.0 + amp * np.sin(3.0 * rm.xc + 2.0 * rm.yc))[:, None]
        out[f"{k}_cvar0"] = s.cvar.copy()
        R = s.compute_residual(0.37)
        out[f"{k}_resid"], out[f"{k}_ws"], out[f"{k}_gx"], out[f"{k}_gy"], out[f"{k}_phi"] = R, s.ws_nrml, s.gx, s.gy, s.phi
    save("ref_resid_mixed.npz", dict(mesh="vortex_mixed_mesh(24)", time=0.37, cases=cases, base=BASE, perturbation="cvar *= 1 + amp*sin(3 xc + 2 yc), amp 0.02 (albada 0.005)"), **out)
    # ---- integrators
    im = meshgen.vortex_mixed_mesh(16)
    rm = ref_mesh(im)
    icases = {"rk1": dict(rk_order=1), "rk2": dict(rk_order=2), "rk3": dict(rk_order=3), "rk4": dict(rk_order=4),
              "ssprk": dict(rk_order=2, ssprk=1), "rk4_steady": dict(rk_order=4, steady=1, cfl_user=0.8),
              "ssprk_steady": dict(rk_order=2, ssprk=1, steady=1, cfl_user=0.8)}
    out = {}
    for k, kw in icases.items():
        s = rn.RefSolver(rm, cfg(grad_method=1, **kw))
        s.initialize_solution()
        r, e, _ = s.time_integration(0.0, 6)
        out[f"{k}_cvar"], out[f"{k}_log_res"], out[f"{k}_verr"] = s.cvar.copy(), r, e
    save("ref_integrators.npz", dict(mesh="vortex_mixed_mesh(16)", steps=6, cases=icases, base=cfg(grad_method=1)), **out)
    # ---- C5 MMS rows
    ns = [16, 32, 64, 128]
    rows = np.zeros((len(ns), 2, 2, 4))
    heff = np.zeros(len(ns))
    for i, n in enumerate(ns):
        rm = ref_mesh(meshgen.mms_mesh(n))
        s = rn.RefSolver(rm, cfg(grad_method=2, lvortex=0, ntstart=0))
        s.initialize_solution()
        heff[i] = rm.heff()
        for j, corr in enumerate((False, True)):
            rows[i, j, 0], rows[i, j, 1] = s.test_resid(corr)
    save("ref_mms.npz", dict(mesh="mms_mesh(n)", n=ns, cfg=cfg(grad_method=2, lvortex=0, ntstart=0), layout="[n][typo|corrected][l2|linf][4]"),
         rows=rows, heff=heff)
    print(f"done in {time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()
