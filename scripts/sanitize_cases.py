"""Small hot-path runs for compute-sanitizer (memcheck / racecheck / synccheck), sized so that every persistent CTA of the
pipeline kernels wraps its shared-memory ring (option ctas=1 on the 65 536-cell NACA mesh: 512 tiles over 148 CTAs):
    compute-sanitizer --tool racecheck python scripts/sanitize_cases.py
Each case is also compared with the CPU oracle (1e-10) so that a sanitizer-clean run is a correct run."""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from fvs2d_b200 import config, meshgen, meshio, solver  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

naca = meshio.load_npz(os.path.join(ROOT, "tests", "golden", "naca_mesh.npz"))
steady = dict(lsteady=True, cfl_user=1.25, rk_order=2, lSSPRK=True, mach_inf=0.8)
CASES = [
    ("fused stage kernel, naca ggcb ssprk steady, ring wraps", naca, dict(grad_cellcntr_imethd=1, **steady), [("ctas", 1)], 2),
    ("two-pass pipeline, naca lsq-nn venkatakrishnan steady, ring wraps", naca,
     dict(grad_cellcntr_imethd=3, grad_cellcntr_lsq_nghbr="nn", grad_limiter_imethd=1, **steady), [("ctas", 1), ("pair", 0)], 2),
    ("two threads per cell, naca lsq-nn venkatakrishnan steady", naca,
     dict(grad_cellcntr_imethd=3, grad_cellcntr_lsq_nghbr="nn", grad_limiter_imethd=1, **steady), [("pair", 1)], 2),
    ("fused stage kernel, mixed mesh lsq-fn rk4 + vortex errors", meshgen.vortex_mixed_mesh(36),
     dict(grad_cellcntr_imethd=3, grad_cellcntr_lsq_nghbr="fn", lvortex=True, dt=0.01), [], 3),
    ("two-pass, mixed mesh ggnb umuscl rk4", meshgen.vortex_mixed_mesh(24),
     dict(grad_cellcntr_imethd=2, face_reconst_imethd=3, umuscl_cst=1.0 / 3.0, lvortex=True, dt=0.01), [], 3),
]
only = os.environ.get("CASE")
ok = True
for k, (name, mesh, kw, opts, nsteps) in enumerate(CASES):
    if only is not None and int(only) != k:
        continue
    cfg = config.RunInput(**kw).to_config()
    gpu = solver.Fvs2dGpu(cfg, device=0)
    for o, v in opts:
        gpu.set_option(o, v)
    gpu.set_mesh(mesh)
    gpu.initialize_solution()
    res, _, _ = gpu.time_integration(0.0, nsteps)
    q = gpu.get_state()
    launches = gpu.last_timing()["launches"]
    gpu.close()
    orc = Oracle(mesh, cfg)
    orc.initialize_solution()
    res_o, _, _ = orc.time_integration(0.0, nsteps)
    dq = float((np.abs(q - orc.cvar) / np.abs(orc.cvar).max(axis=0)).max())
    dr = float((np.abs(res - res_o) / np.abs(res_o)).max())
    good = dq <= 1e-10 and dr <= 1e-10
    ok = ok and good
    print(f"CASE {k} {name}: cells {mesh.ncells} steps {nsteps} launches {launches} state {dq:.2e} log_res {dr:.2e} {'ok' if good else 'FAILED'}", flush=True)
sys.exit(0 if ok else 1)
