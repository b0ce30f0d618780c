"""The reference's two shipped examples, exactly as shipped (examples/isentropic_vortex: 4 000 RK4 steps, 50 saves;
examples/naca0012_ogrid: 50 000 steady SSPRK steps, 100 saves), end to end through the drop-in host program
fvs2d_gpu.exe -- input files in, log_*.plt / inst.* / save.* files out -- with the wall-clock of the whole program, next to
the CPU oracle's time for the same step counts (its measured per-step time x steps; one core, -Ofast).
    python scripts/run_examples.py [workdir]      -> gpurun_out/run_examples.txt"""
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import run_input  # noqa: E402
from fvs2d_b200 import config, meshio  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

EXE = os.path.join(ROOT, "fvs2d_b200", "csrc", "fvs2d_gpu.exe")
OUT = os.path.join(ROOT, "gpurun_out", "run_examples.txt")
os.makedirs(os.path.dirname(OUT), exist_ok=True)


def say(*a):
    line = " ".join(str(x) for x in a)
    print(line, flush=True)
    with open(OUT, "a") as f:
        f.write(line + "\n")


base = sys.argv[1] if len(sys.argv) > 1 else tempfile.mkdtemp(prefix="fvs2d_examples_")
for name, grid in (("vortex", "vortex"), ("naca", "naca0012_ogrid")):
    r = run_input(name)
    mesh = meshio.load_npz(os.path.join(ROOT, "tests", "golden", f"{name}_mesh.npz"))
    d = os.path.join(base, name)
    os.makedirs(d, exist_ok=True)
    meshio.write_mesh(os.path.join(d, r.grid_base), mesh)
    config.write_input(os.path.join(d, "fvs2d.input"), r)
    t0 = time.perf_counter()
    out = subprocess.run([EXE, "0"], cwd=d, capture_output=True, text=True, timeout=1500)
    wall = time.perf_counter() - t0
    ok = out.returncode == 0 and "o.k." in out.stdout
    files = sorted(f for f in os.listdir(d) if f.startswith(("log_", "inst.", "save.", "cont.")))
    tail = [ln for ln in out.stdout.splitlines() if "time" in ln.lower() or "min" in ln.lower()][-4:]
    # CPU oracle: per-step time on a few steps of the same run
    orc = Oracle(mesh, r.to_config(), fast=True)
    orc.initialize_solution()
    nprobe = 20
    orc.time_integration(0.0, 2)
    t1 = time.perf_counter()
    orc.time_integration(2 * r.dt, nprobe)
    cpu_step = (time.perf_counter() - t1) / nprobe
    say(f"EXAMPLE {name}: {mesh.ncells} cells, {r.ntimes} steps, {r.nsaves} saves: fvs2d_gpu.exe {'ok' if ok else 'FAILED'} in {wall:.2f} s wall "
        f"(whole program: read, set-up, time loop, output files {files}); CPU oracle {cpu_step * 1e3:.2f} ms per step -> "
        f"{cpu_step * r.ntimes:.0f} s for the time loop alone; ratio {cpu_step * r.ntimes / wall:.0f}x")
    for ln in tail:
        say("   ", ln.strip())
    if not ok:
        say(out.stdout[-2000:], out.stderr[-2000:])
