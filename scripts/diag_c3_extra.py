"""Diagnostic: production-path step time of C3 as a function of warm-up length and call sequence."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from fvs2d_b200 import solver

mesh, run, desc, _ = bench.make_workload("c3", 1)
g = solver.Fvs2dGpu(run.to_config(1), device=0)
g.set_mesh(mesh)
nc = mesh.ncells
g.initialize_solution()
t = 0.0
for n in (3, 20, 20, 20, 40, 20):
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    g.time_integration(t, n, logs=False)
    torch.cuda.synchronize()
    w = time.perf_counter() - w0
    tm = g.last_timing()
    t += n * run.dt
    print(f"nsub {n:3d}  total_ms/step {tm['total_ms'] / n:.3f}  wall/step {w * 1e3 / n:.3f}  G/s {nc * 4 * n / tm['total_ms'] / 1e6:.2f}", flush=True)
time.sleep(8.0)
for n in (3, 20, 20):
    g.time_integration(t, n, logs=False)
    tm = g.last_timing()
    t += n * run.dt
    print(f"after 8 s idle: nsub {n:3d}  total_ms/step {tm['total_ms'] / n:.3f}", flush=True)
g.close()
