"""diagnostic: growth of GPU-vs-oracle difference for the limited NACA run, and single-evaluation mismatch at a developed state"""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
from conftest import run_input
from fvs2d_b200 import meshio, solver
from oracle.oracle import Oracle
mesh = meshio.load_npz("tests/golden/naca_mesh.npz")
for lim in (1, 2):
    r = run_input("naca"); r.grad_limiter_imethd = lim
    cfg = r.to_config()
    gpu = solver.Fvs2dGpu(cfg, device=0); gpu.set_mesh(mesh); gpu.initialize_solution()
    orc = Oracle(mesh, cfg); orc.initialize_solution()
    done = 0
    for n in (1, 2, 5, 10, 20, 40, 70, 100):
        gpu.time_integration(done * r.dt, n - done); orc.time_integration(done * r.dt, n - done); done = n
        q, qo = gpu.get_state(), orc.cvar
        d = np.abs(q - qo) / np.abs(qo).max(axis=0)
        # single evaluation at the GPU's state
        orc2 = Oracle(mesh, cfg); orc2.set_state(q)
        ro = orc2.compute_residual(0.0); rg = gpu.compute_residual(0.0)
        pv, gr, ph = gpu.get_aux()
        pho = orc2.array("phi_lim")
        dr = np.abs(rg - ro) / np.abs(ro).max(axis=0)
        print(f"lim {lim} step {n:4d} state diff {d.max():.3e}  single-eval: phi diff {np.abs(ph-pho).max():.3e} at {np.abs(ph-pho).argmax()} "
              f"resid diff {dr.max():.3e}  min phi {pho.min():.3f}  #phi<1: {(pho<0.999).sum()}")
    gpu.close()
