"""torchrun diagnostic: per-call wall time of the C-ABI calls under a communicator."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from fvs2d_b200 import config, meshgen, solver
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0: uid.copy_(torch.frombuffer(bytearray(solver.Fvs2dGpu.comm_unique_id()), dtype=torch.uint8))
dist.broadcast(uid, 0)
mesh = meshgen.make_mesh(2400, 600 * world, 20.0, 10.0, (600, 1800))
run = config.RunInput(grad_cellcntr_imethd=1, lvortex=True, dt=1e-3)
gpu = solver.Fvs2dGpu(run.to_config(world), device=local, comm=(rank, world, uid.cpu().numpy().tobytes()))
gpu.set_mesh(mesh); gpu.initialize_solution()
n_own = gpu.sizes()["ncells_own"]
q = torch.empty((n_own, 4), dtype=torch.float64).pin_memory()
def T(name, fn, n=5):
    for i in range(n):
        torch.cuda.synchronize(); dist.barrier(); t = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t
        if rank == 0: print(f"{name:40s} call {i}: {dt*1e3:9.3f} ms", flush=True)
T("get_state_local", lambda: gpu.get_state_local(q))
T("set_state_local", lambda: gpu.set_state_local(q))
T("time_integration(1, logs=False)", lambda: gpu.time_integration(0.0, 1, logs=False))
T("time_integration(1, logs=True)", lambda: gpu.time_integration(0.0, 1, logs=True))
T("time_integration(10, logs=True)", lambda: gpu.time_integration(0.0, 10, logs=True), 3)
print(rank, gpu.last_timing(), flush=True)
gpu.close(); dist.destroy_process_group()
