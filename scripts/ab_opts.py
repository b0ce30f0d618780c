"""A/B of library options on the bench meshes (device time per step of the default, fused path), with a bitwise comparison
of the state after 10 steps between the option sets:
    python scripts/ab_opts.py "nreg=0" "nreg=1" -- c3 c4
Appends its lines to gpurun_out/ab_opts.txt.  numpy + ctypes only (no torch)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from fvs2d_b200 import config, meshgen, solver  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out", "ab_opts.txt")
os.makedirs(os.path.dirname(OUT), exist_ok=True)


def say(*a):
    line = " ".join(str(x) for x in a)
    print(line, flush=True)
    with open(OUT, "a") as f:
        f.write(line + "\n")


def main():
    args = sys.argv[1:]
    cut = args.index("--") if "--" in args else len(args)
    sets = [[kv.split("=") for kv in a.split(",") if kv] for a in args[:cut]] or [[]]
    which = args[cut + 1:] or ["c3", "c4"]
    for w in which:
        if w == "c3":
            mesh, dt = meshgen.vortex_tri_mesh(2000), 0.002
        elif w == "c4":
            mesh, dt = meshgen.make_mesh(9600, 600, 20.0, 1.25, (2400, 7200)), 4e-4
        else:
            mesh, dt = meshgen.vortex_tri_mesh(int(w)), 0.002
        cfg = config.RunInput(grad_cellcntr_imethd=1, lvortex=True, dt=dt).to_config()
        gpu = solver.Fvs2dGpu(cfg, device=0)
        gpu.set_mesh(mesh)
        states = []
        for rep in range(2):
            for si, opts in enumerate(sets):
                for k, v in opts:
                    gpu.set_option(k, int(v))
                gpu.set_option("timing", 0)
                if rep == 0:
                    gpu.initialize_solution()
                    gpu.time_integration(0.0, 10, logs=False)
                    states.append(gpu.get_state())
                gpu.time_integration(0.0, 5, logs=False)
                time.sleep(1.0)                       # every measurement starts from an idle GPU (power state)
                gpu.time_integration(0.0, 20, logs=False)
                ms = gpu.last_timing()["total_ms"] / 20
                gpu.set_option("timing", 1)           # event pair around every launch (eager path)
                gpu.time_integration(0.0, 5, logs=False)
                tm = gpu.last_timing()
                say("AB", w, f"cells {mesh.ncells}", ",".join("=".join(o) for o in opts) or "(default)",
                    f"{ms:.4f} ms/step = {mesh.ncells * 4 / ms / 1e6:.3f} G cell-stages/s; stage kernels {tm['flux_ms'] / 20:.4f} ms; "
                    f"launches/step {tm['launches'] / 5:.0f}")
        for si in range(1, len(states)):
            say("AB", w, f"state after 10 steps, set {si} vs set 0: bitwise {bool(np.array_equal(states[0], states[si]))}, "
                f"max rel {float(np.abs(states[si] - states[0]).max() / np.abs(states[0]).max()):.2e}, finite {bool(np.isfinite(states[si]).all())}")
        gpu.close()


if __name__ == "__main__":
    main()
