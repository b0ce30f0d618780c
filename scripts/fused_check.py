"""Option "fuse" (one kernel per RK stage, k_stage_fused) against the two-pass production path on the GPU:
    python scripts/fused_check.py parity          small meshes, 10 steps: state / logs of fused vs two-pass
    python scripts/fused_check.py time c3 [c4]    ms per step of both paths on the bench meshes
Appends one line per result to gpurun_out/fused_check.txt (flushed as it goes).  numpy + ctypes only (no torch)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from fvs2d_b200 import config, meshgen, meshio, solver  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out", "fused_check.txt")
FUSE = tuple(int(x) for x in os.environ.get("FUSE", "-1,2").split(","))
os.makedirs(os.path.dirname(OUT), exist_ok=True)


def say(*a):
    line = " ".join(str(x) for x in a)
    print(line, flush=True)
    with open(OUT, "a") as f:
        f.write(line + "\n")


def run(gpu, nsteps, fuse, t0=0.0):
    gpu.set_option("fuse", fuse)
    gpu.initialize_solution()
    res, ve, _ = gpu.time_integration(t0, nsteps)
    return gpu.get_state(), res, ve


def parity():
    naca = meshio.load_npz(os.path.join(ROOT, "tests", "golden", "naca_mesh.npz"))
    cases = [
        ("tri ggcb rk4", meshgen.vortex_tri_mesh(40), dict(grad_cellcntr_imethd=1, lvortex=True, dt=0.01)),
        ("mixed ggcb rk4 (ragged last tile)", meshgen.vortex_mixed_mesh(36), dict(grad_cellcntr_imethd=1, lvortex=True, dt=0.01)),
        ("tri lsq-fn rk4", meshgen.vortex_tri_mesh(50), dict(grad_cellcntr_imethd=3, grad_cellcntr_lsq_nghbr="fn", lvortex=True, dt=0.01)),
        ("mixed lsq-nn ssprk", meshgen.vortex_mixed_mesh(32), dict(grad_cellcntr_imethd=3, grad_cellcntr_lsq_nghbr="nn", lvortex=True, dt=0.01,
                                                                   rk_order=2, lSSPRK=True)),
        ("naca ggcb ssprk steady (slip wall + freestream)", naca, dict(grad_cellcntr_imethd=1, lsteady=True, cfl_user=1.25, rk_order=2,
                                                                       lSSPRK=True, mach_inf=0.8)),
    ]
    ok = True
    for name, mesh, kw in cases:
        try:
            cfg = config.RunInput(**kw).to_config()
        except TypeError as e:  # option names differ: report and go on
            say("PARITY", name, "SKIPPED", e)
            continue
        gpu = solver.Fvs2dGpu(cfg, device=0)
        gpu.set_mesh(mesh)
        q0, r0, v0 = run(gpu, 10, 0)
        for fuse in FUSE:
            q1, r1, v1 = run(gpu, 10, fuse)
            used = gpu.last_timing()["launches"]
            dq = float(np.abs(q1 - q0).max() / np.abs(q0).max())
            dr = float(np.abs(r1 - r0).max() / np.abs(r0).max())
            good = dq <= 1e-12 and dr <= 1e-10 and np.isfinite(q1).all()
            ok = ok and good
            say("PARITY", name, f"cells {mesh.ncells} fuse={fuse} vs two-pass: state {dq:.3e} log_res {dr:.3e} bitwise "
                f"{bool(np.array_equal(q0, q1))} launches/10 steps {used}", "ok" if good else "FAILED")
        gpu.close()
    say("PARITY_OK" if ok else "PARITY_FAILED")
    return ok


def timing(which):
    for w in which:
        t = time.time()
        if w == "c3":
            mesh, dt = meshgen.vortex_tri_mesh(2000), 0.002
        elif w == "c4":
            mesh, dt = meshgen.make_mesh(9600, 600, 20.0, 1.25, (2400, 7200)), 4e-4
        else:
            mesh, dt = meshgen.vortex_tri_mesh(int(w)), 0.002
        cfg = config.RunInput(grad_cellcntr_imethd=1, lvortex=True, dt=dt).to_config()
        gpu = solver.Fvs2dGpu(cfg, device=0)
        gpu.set_mesh(mesh)
        say("TIME", w, f"cells {mesh.ncells} mesh+setup {time.time() - t:.1f}s")
        gpu.initialize_solution()
        for fuse in (0,) + FUSE + (0,) + FUSE:
            gpu.set_option("fuse", fuse)
            gpu.set_option("timing", 0)
            gpu.time_integration(0.0, 5, logs=False)
            time.sleep(1.0)                       # start every measurement from an idle GPU (power state)
            gpu.time_integration(0.0, 20, logs=False)
            ms = gpu.last_timing()["total_ms"] / 20
            gpu.set_option("timing", 1)   # event pair around every launch (eager path)
            gpu.time_integration(0.0, 5, logs=False)
            tm = gpu.last_timing()
            say("TIME", w, f"fuse {fuse}: {ms:.4f} ms/step = {mesh.ncells * 4 / ms / 1e6:.3f} G cell-stages/s; per launch: grad "
                f"{tm['grad_ms'] / 20:.4f} flux {tm['flux_ms'] / 20:.4f} ms; launches/step {tm['launches'] / 5:.0f}")
        gpu.close()


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "parity"
    if mode == "parity":
        sys.exit(0 if parity() else 1)
    timing(sys.argv[2:] or ["c3"])
