"""Run under torchrun (one rank per GPU): the domain-decomposed run (NCCL halo exchange) against the CPU oracle
and, bit for bit, against ... itself is not possible across partitionings, so the reference is the oracle.
Prints PARITY_OK on rank 0.  FUSE=<n> in the environment selects the library's "fuse" option before set_mesh (the fused
stage kernel on several ranks needs the deep ghost layers, which are decided when the mesh is set)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from fvs2d_b200 import config, meshgen, meshio, solver  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    def new_comm():  # one NCCL unique id per communicator
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(solver.Fvs2dGpu.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        return (rank, world, uid.cpu().numpy().tobytes())
    ok = True
    cases = [
        ("tri ggcb rk4", meshgen.vortex_tri_mesh(64), config.RunInput(grad_cellcntr_imethd=1, lvortex=True, dt=0.005), 10),
        ("mixed lsq-nn venk ssprk steady", meshgen.vortex_mixed_mesh(64),
         config.RunInput(grad_cellcntr_imethd=3, grad_cellcntr_lsq_nghbr="nn", grad_limiter_imethd=1, lvortex=True, dt=0.005,
                         rk_order=2, lSSPRK=True, lsteady=True, cfl_user=0.8), 6),
        ("mixed ggnb umuscl", meshgen.vortex_mixed_mesh(48),
         config.RunInput(grad_cellcntr_imethd=2, face_reconst_imethd=3, umuscl_cst=1.0 / 3.0, lvortex=True, dt=0.005), 6),
    ]
    cases.append(("tri lsq-fn rk4 (whole-mesh pre-processing: boundary stencils need the global nearest-centroid search)",
                  meshgen.vortex_tri_mesh(48), config.RunInput(grad_cellcntr_imethd=3, grad_cellcntr_lsq_nghbr="fn", lvortex=True, dt=0.005), 6))
    naca = meshio.load_npz(os.path.join(ROOT, "tests", "golden", "naca_mesh.npz"))
    cases.append(("naca ggcb ssprk steady (slip wall + freestream; wall values)", naca,
                  config.RunInput(grad_cellcntr_imethd=1, lsteady=True, cfl_user=1.25, rk_order=2, lSSPRK=True, mach_inf=0.8), 5))
    if os.environ.get("BIG"):  # the C4 sibling bench.py's parity block uses: >= 4 tiles per persistent CTA on 2 ranks
        cases.append(("c4 sibling 540k mixed ggcb rk4", meshgen.make_mesh(2400, 150, 20.0, 10.0, (600, 1800)),
                      config.RunInput(grad_cellcntr_imethd=1, lvortex=True, dt=1.6e-3), 10))
    for name, mesh, run, nsteps in cases:
        cfg = run.to_config(world)
        gpu = solver.Fvs2dGpu(cfg, device=local, comm=new_comm())
        if os.environ.get("FUSE"):
            gpu.set_option("fuse", int(os.environ["FUSE"]))
        gpu.set_mesh(mesh)
        gpu.initialize_solution()
        res, ve, vxy = gpu.time_integration(0.0, nsteps)
        q = np.zeros((mesh.ncells, 4))
        gpu.get_state(q)                      # fills the owned cells only
        qt = torch.from_numpy(q).cuda()
        dist.all_reduce(qt)                   # disjoint ownership -> sum assembles the global state
        q = qt.cpu().numpy()
        sizes = gpu.sizes()
        launches = gpu.last_timing()["launches"]
        # output path under the communicator: the ranks' shares of the node values / their own wall edges add up
        fn = torch.from_numpy(gpu.interpolate_cell2node((1, 1, 1, 1))).cuda()
        dist.all_reduce(fn)
        fn = fn.cpu().numpy()
        walls = {}
        for ib, bt in enumerate(mesh.bndry_type):
            if bt == "slip_wall":
                w = torch.from_numpy(gpu.wall_values(ib)).cuda()
                dist.all_reduce(w)
                walls[ib] = w.cpu().numpy()
        gpu.close()
        if rank == 0:
            from oracle.oracle import Oracle
            orc = Oracle(mesh, cfg)
            orc.initialize_solution()
            res_o, ve_o, vxy_o = orc.time_integration(0.0, nsteps)
            q_o = orc.cvar
            eq = float((np.abs(q - q_o) / np.abs(q_o).max(axis=0)).max())
            er = float((np.abs(res - res_o) / np.abs(res_o)).max())
            if ve is not None:
                ev = float((np.abs(ve - ve_o) / np.maximum(np.abs(ve_o), 1e-300)).max())
                exy = float(np.abs(vxy - vxy_o).max())
            else:
                ev = exy = 0.0
            en = max(float(np.abs(fn[v] - orc.interpolate_cell2node(v)).max() / np.abs(fn[v]).max()) for v in range(4))
            ew = max([float(np.abs(w - orc.wall_values(ib)).max() / np.abs(orc.wall_values(ib)).max()) for ib, w in walls.items()] + [0.0])
            good = eq <= 1e-10 and er <= 1e-9 and ev <= 1e-8 and exy == 0.0 and en <= 1e-12 and ew <= 1e-10
            ok = ok and good
            print(f"[{world} ranks fuse={os.environ.get('FUSE', 'default')}] {name}: cells {mesh.ncells} own(rank0) {sizes['ncells_own']} local {sizes['ncells_local']} launches {launches} "
                  f"state {eq:.2e} log_res {er:.2e} vortex {ev:.2e} xy {exy:.1e} nodes {en:.1e} wall {ew:.1e} -> {'ok' if good else 'FAIL'}", flush=True)
    if rank == 0:
        print("PARITY_OK" if ok else "PARITY_FAIL", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
