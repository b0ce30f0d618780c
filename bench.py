#!/usr/bin/env python
"""bench.py -- cell-RK-stage updates per second of the fvs2d hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c4|c3|naca|vortex]

One "step" = one full RK time step (4 stages: gradient -> Roe flux gather -> residual -> RK update, plus
the per-step residual / vortex-error norms) of `time_integration` over the whole mesh.
Metric = ncells * 4 * K / seconds (BASELINE.json: "cell-RK-stage updates/sec").
Besides the contract's keys the line carries `parity` (GPU on all ranks vs the CPU oracle on a sibling mesh, run before
the timed region), `state_check` (isfinite + conservation of the timed state) and `sustained` (>= --sustain-s seconds).

Workloads:
  c4     (default) synthetic mixed tri/quad vortex mesh, 8.64 M cells per GPU (9600 x 600*N background quads;
         N=8 is the 69.12 M-cell C4 mesh), GGCB, RK4, dt=4e-4 -- weak scaling
  c3     synthetic 4.0 M-triangle vortex mesh (2000 x 1000 split quads), GGCB, RK4, dt=0.002 (single GPU)
  naca   tests/golden/naca_mesh.npz, LSQ-nn + Venkatakrishnan, SSPRK steady (C2)
  vortex tests/golden/vortex_mesh.npz as shipped (C1)
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from fvs2d_b200 import config as fcfg  # noqa: E402
from fvs2d_b200 import meshgen, meshio  # noqa: E402

# algorithmic bytes per cell-stage, SURVEY.md section 8(d): (pass A, pass B)
B_ALG = {"tri_ggcb": (180, 332), "quad_ggcb": (200, 360), "tri_lsqfn": (156, 332), "quad_lsqnn_venk_steady": (336, 376)}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def workload_desc(name: str, ngpus: int, scale: float = 1.0) -> str:
    if name == "c4":
        nx, ny = int(round(9600 * scale)), int(round(600 * scale)) * ngpus
        return f"C4 synthetic mixed tri/quad vortex mesh {nx}x{ny} background quads, GGCB, upwind-2nd, Roe, RK4"
    if name == "c3":
        nx = int(round(2000 * scale))
        return f"C3 synthetic triangle vortex mesh {nx}x{nx // 2} split quads, GGCB, upwind-2nd, Roe, RK4"
    return {"naca": "C2 naca0012_ogrid, LSQ-nn + Venkatakrishnan, SSPRK(4,2) steady CFL 1.25",
            "vortex": "C1 isentropic_vortex example as shipped, LSQ-fn, RK4"}[name]


def make_workload(name: str, ngpus: int, scale: float = 1.0):
    """-> (mesh, RunInput, description, (bytes_passA, bytes_passB) per cell-stage)"""
    golden = os.path.join(ROOT, "tests", "golden")
    if name == "c4":
        nx = int(round(9600 * scale))
        ny = int(round(600 * scale)) * ngpus
        mesh = meshgen.make_mesh(nx, ny, 20.0, 10.0, (nx // 4, 3 * nx // 4))
        run = fcfg.RunInput(grad_cellcntr_imethd=1, lvortex=True, dt=4e-4 / scale)
        ft = mesh.ntri / mesh.ncells
        ba = ft * B_ALG["tri_ggcb"][0] + (1 - ft) * B_ALG["quad_ggcb"][0]
        bb = ft * B_ALG["tri_ggcb"][1] + (1 - ft) * B_ALG["quad_ggcb"][1]
        return mesh, run, workload_desc(name, ngpus, scale), (ba, bb)
    if name == "c3":
        nx = int(round(2000 * scale))
        mesh = meshgen.vortex_tri_mesh(nx)
        run = fcfg.RunInput(grad_cellcntr_imethd=1, lvortex=True, dt=0.002 / scale)
        return mesh, run, workload_desc(name, ngpus, scale), B_ALG["tri_ggcb"]
    if name.startswith("x:"):  # experiment: x:nx,ny,b0,b1 (b0==b1==0: all triangles)
        nx, ny, b0, b1 = (int(t) for t in name[2:].split(","))
        mesh = meshgen.make_mesh(nx, ny, 20.0, 10.0, (b0, b1) if b1 > b0 else None)
        run = fcfg.RunInput(grad_cellcntr_imethd=1, lvortex=True, dt=0.1 * min(20.0 / nx, 10.0 / ny))
        ft = mesh.ntri / mesh.ncells
        return mesh, run, f"experiment {name}", (ft * 180 + (1 - ft) * 200, ft * 332 + (1 - ft) * 360)
    inp = json.load(open(os.path.join(golden, "inputs.json")))
    if name == "naca":
        mesh = meshio.load_npz(os.path.join(golden, "naca_mesh.npz"))
        d = inp["naca"]
        run = fcfg.RunInput(**{k: (tuple(v) if isinstance(v, list) else v) for k, v in d.items()})
        run.grad_limiter_imethd = 1
        return mesh, run, workload_desc(name, ngpus, scale), B_ALG["quad_lsqnn_venk_steady"]
    if name == "vortex":
        mesh = meshio.load_npz(os.path.join(golden, "vortex_mesh.npz"))
        d = inp["vortex"]
        run = fcfg.RunInput(**{k: (tuple(v) if isinstance(v, list) else v) for k, v in d.items()})
        return mesh, run, workload_desc(name, ngpus, scale), B_ALG["tri_lsqfn"]
    raise SystemExit(f"unknown workload {name}")


class ClockSampler:
    """SM clock / power / throttle reasons DURING the timed region (B200_PROFILING.md recipe), polled through NVML
    every few ms (an nvidia-smi subprocess would return one sample per ~100 ms region); nvidia-smi is the fallback."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int):
        self.rows, self.stop, self.index, self.max_mhz = [], False, index, None
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            while not self.stop:
                self.rows.append((float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)), pynvml.nvmlDeviceGetPowerUsage(h) / 1e3,
                                  int(get_reasons(h))))
                time.sleep(0.004)
            return
        except Exception:
            pass
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        bits = [0x8, 0x40, 0x20, 0x4]
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                t = [x.strip() for x in out.split(",")]
                self.max_mhz = float(t[1])
                self.rows.append((float(t[0]), float(t[2]), sum(b for b, x in zip(bits, t[3:7]) if x.lower().startswith("active"))))
            except Exception:
                pass
            time.sleep(0.05)

    def __enter__(self):
        self.th.start()
        time.sleep(0.02)
        return self

    def __exit__(self, *a):
        self.stop = True
        self.th.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock samples"]}
        sm = sorted(r[0] for r in self.rows)
        mask = 0
        for r in self.rows:
            mask |= r[2]
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.max_mhz, "samples": len(self.rows),
                "power_w_max": max(r[1] for r in self.rows), "reasons": [n for b, n in self.REASONS.items() if mask & b]}


def cpu_baseline(name: str, run, steps: int, warmup: int = 0, full: bool = False, scale: float = 1.0):
    """The CPU oracle (C restatement of the reference algorithm, -Ofast, ONE thread: the reference's only
    multi-thread mode is racy, src/residual.f90:65) on a bounded sample of the workload: the 1/16-size sibling
    (cpu_baseline leg of the GPU arm, ~10 s) or, full=True (the --impl reference arm), the whole single-GPU mesh."""
    from oracle.oracle import Oracle, build
    build()
    if name in ("c4", "c3") and not full:
        mesh, run_s, desc, _ = make_workload(name, 1, scale=0.25)
        sample = f"1/16-size sibling of the workload ({mesh.ncells} cells, same generator/seed/scheme), {steps} RK4 steps"
    else:
        mesh, run_s, desc, _ = make_workload(name, 1, scale=scale)
        sample = f"the full single-GPU {name} mesh ({mesh.ncells} cells), {steps} steps"
    orc = Oracle(mesh, run_s.to_config(), fast=True)
    orc.initialize_solution()
    if warmup:
        orc.time_integration(0.0, warmup)
    orc.reset_timers()
    t0 = time.perf_counter()
    orc.time_integration(warmup * run_s.dt, steps)
    dt = time.perf_counter() - t0
    tm = orc.timers()
    out = {"value": mesh.ncells * 4 * steps / dt, "unit": "cell-RK-stage updates/s", "cores": 1, "kind": "port",
           "sample": sample, "seconds": dt, "host_cores_available": os.cpu_count(),
           "buckets_s": {k: round(v, 4) for k, v in tm.items()}}
    del orc
    # context only: a race-free all-cores variant -- edge-parallel flux + cell gather under OpenMP.  It is
    # NOT the reference algorithm (the reference's threaded flux loop races and its GGCB/GGNB/limiter/RK loops are serial).
    try:
        nthr = len(os.sched_getaffinity(0))
        os.environ["OMP_NUM_THREADS"] = str(nthr)
        omp = Oracle(mesh, run_s.to_config(), fast="omp")
        omp.initialize_solution()
        omp.time_integration(0.0, 1)
        so = max(2, min(steps, 6)) if full else steps
        t0 = time.perf_counter()
        omp.time_integration(run_s.dt, so)
        dto = time.perf_counter() - t0
        out["all_cores_variant"] = {"value": mesh.ncells * 4 * so / dto, "cores": nthr, "seconds": dto,
                                    "note": "OpenMP gather variant of the oracle; not the reference algorithm"}
    except Exception as e:  # the context number must never break the bench line
        out["all_cores_variant"] = {"unavailable": str(e)[:200]}
    return out, dt / steps * 1e3


def bind_to_gpu_cpus(local_rank: int):
    """Several ranks on one host: run this rank on the CPUs next to its GPU (NVML's ideal-CPU mask, as a job launcher's NUMA
    binding would), so that its pinned host buffers are first touched on the memory node the GPU's PCIe link hangs off.
    Without it the e2e leg (553 MB over PCIe per rank and step) crosses the socket interconnect for about half of the ranks.
    Never fatal: returns the CPU list it bound to, or None when NVML gives no usable mask inside the allowed CPU set."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(local_rank)
        try:
            h = pynvml.nvmlDeviceGetHandleByPciBusId(f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0")
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        allowed = os.sched_getaffinity(0)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (max(max(allowed) + 1, os.cpu_count() or 1) + 63) // 64)
        ideal = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        use = ideal & allowed
        if not use or use == allowed:
            return None
        os.sched_setaffinity(0, use)
        return sorted(use)
    except Exception:
        return None


def parity_block(workload: str, world: int, rank: int, local_rank: int, new_comm, opts):
    """GPU (all `world` ranks) against the CPU oracle (parity build, rank 0) on a sibling of the workload small enough for
    the oracle, same generator / seed / scheme, state and log_res after `steps` RK4 steps.  c4 / c3: the 1/16-size sibling
    (540 000 / 250 000 cells; >= 4 tiles per persistent CTA on one GPU), 1/4-size from 4 ranks on."""
    import torch
    import torch.distributed as dist
    from fvs2d_b200 import solver
    if workload in ("c4", "c3"):
        scale, steps = (0.25, 10) if world < 4 else (0.5, 6)
    else:
        scale, steps = 1.0, 10
    mesh, run_s, desc, _ = make_workload(workload, 1, scale=scale)
    cfg = run_s.to_config(world)
    gpu = solver.Fvs2dGpu(cfg, device=local_rank, comm=new_comm())
    for k, v in opts:
        gpu.set_option(k, v)
    gpu.set_mesh(mesh)
    gpu.initialize_solution()
    res, _, _ = gpu.time_integration(0.0, steps)
    q = np.zeros((mesh.ncells, 4))
    gpu.get_state(q)                       # several ranks: fills the owned cells only
    launches = gpu.last_timing()["launches"]
    gpu.close()
    if world > 1:
        qt = torch.from_numpy(q).cuda()
        dist.all_reduce(qt)                # disjoint ownership: the sum assembles the global state
        q = qt.cpu().numpy()
    out = None
    if rank == 0:
        from oracle.oracle import Oracle
        orc = Oracle(mesh, cfg)
        orc.initialize_solution()
        res_o, _, _ = orc.time_integration(0.0, steps)
        q_o = orc.cvar
        out = {"max_rel_state": float((np.abs(q - q_o) / np.abs(q_o).max(axis=0)).max()),
               "max_rel_log_res": float((np.abs(res - res_o) / np.abs(res_o)).max()), "ranks": world, "steps": steps,
               "ncells": mesh.ncells, "finite": bool(np.isfinite(q).all()), "launches": launches,
               "against": "CPU oracle (oracle/liboracle.so, -O2 -ffp-contract=off) on the same mesh and inputs; tolerance 1e-10",
               "mesh": desc}
        out["ok"] = out["finite"] and out["max_rel_state"] <= 1e-10 and out["max_rel_log_res"] <= 1e-10
    if world > 1:
        dist.barrier()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4")
    ap.add_argument("--scale", type=float, default=1.0, help="mesh refinement factor relative to the named workload")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the additional C3 measurement at N=1")
    ap.add_argument("--opt", action="append", default=[], help="library tuning option key=value (fvs2d_gpu_set_option)")
    ap.add_argument("--sustain-s", type=float, default=2.0, help="length of the second, sustained timed region in seconds (0: skip)")
    ap.add_argument("--no-parity", action="store_true", help="skip the GPU-vs-oracle parity block")
    ap.add_argument("--mesh-for-gpus", type=int, default=0, help="build the weak-scaling mesh of this many GPUs whatever --gpus says "
                    "(c4: --mesh-for-gpus 8 --gpus 1 runs the full 69.12 M-cell C4 mesh on one GPU: the strong-scaling denominator)")
    args = ap.parse_args()
    K, W = args.steps, max(args.warmup, 0)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    metric = "cell-RK-stage updates/sec"

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        desc = workload_desc(args.workload, max(args.gpus, 1), args.scale)
        # the whole single-GPU mesh as long as K + W steps of it fit ~3 minutes of one host core (8.64 M cells: ~6 s per RK4
        # step, i.e. up to 30 steps); beyond that each step is a bounded sample -- the 1/4- or 1/16-size sibling of the mesh
        # (same generator, seed and scheme; the per-cell rate is size-independent to a few per cent)
        ref_scale = args.scale
        if args.workload in ("c4", "c3"):
            est_s = (K + W) * (8.64e6 if args.workload == "c4" else 4.0e6) * args.scale ** 2 * 4 / 5.7e6
            ref_scale = args.scale * (1.0 if est_s <= 190 else 0.5 if est_s <= 4 * 190 else 0.25)
        cb, ms_step = cpu_baseline(args.workload, None, max(K, 1), W, full=True, scale=ref_scale)
        if ref_scale != args.scale:
            cb["sample"] += f" -- a {ref_scale / args.scale:g}x-refinement sibling of the workload, chosen so that {K + W} steps end within a few minutes"
        line = {"metric": metric, "value": cb["value"], "unit": "cell-stage updates/s", "n_gpus": args.gpus, "steps": K, "warmup": W,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "impl": "reference",
                "config": {"workload": desc, "parallelism": "1 host thread (the reference's OpenMP flux loop races, src/residual.f90:65)",
                           "note": "reference algorithm on the host (C restatement in oracle/, -Ofast; the Fortran reference cannot be "
                                   "compiled in this image); each step is one RK4 step over " + cb["sample"] +
                                   (" (a bounded sample)" if ref_scale != args.scale else " = the workload itself" if args.gpus <= 1
                                    else " = one GPU's share of the workload (a bounded sample)")},
                "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "cell-stage updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ our arm (GPU)
    import torch
    import torch.distributed as dist
    from fvs2d_b200 import solver

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    cpu_binding = bind_to_gpu_cpus(local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def new_comm():  # one NCCL unique id per library context
        if world == 1:
            return None
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(solver.Fvs2dGpu.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        return (rank, world, bytes(uid.cpu().numpy().tobytes()))
    ngpus = world
    opts = [(kv.split("=")[0], int(kv.split("=")[1])) for kv in args.opt]

    # ---- parity first (its own library context): the configuration that is about to be timed, on a mesh the oracle can do
    parity = None
    if not args.no_parity and args.scale == 1.0:
        parity = parity_block(args.workload, world, rank, local_rank, new_comm, opts)
    comm = new_comm()

    t_setup = time.perf_counter()
    mesh, run, desc, (bA, bB) = make_workload(args.workload, args.mesh_for_gpus or ngpus, args.scale)
    t_meshgen = time.perf_counter() - t_setup         # the synthetic-mesh generator (numpy; stands in for reading a .grid file)
    cfg = run.to_config(ngpus)
    gpu = solver.Fvs2dGpu(cfg, device=local_rank, comm=comm)
    for k, v in opts:
        gpu.set_option(k, v)
    t_lib = time.perf_counter()
    gpu.set_mesh(mesh)
    ncells = mesh.ncells
    del mesh                                          # the library holds its own copy
    gpu.initialize_solution()
    sizes, scal = gpu.sizes(), gpu.scalars()
    t_lib = time.perf_counter() - t_lib               # fvs2d_gpu_set_mesh + fvs2d_gpu_initialize_solution (the library's share)
    t_setup = time.perf_counter() - t_setup
    from fvs2d_b200 import capi
    vol_own = capi.mesh_array("lvol")[:sizes["ncells_own"]]

    def conserved_totals():
        """sum over the owned cells of vol * (rho, rho u, rho v, rho E), summed over the ranks; isfinite of the state"""
        qo = np.zeros((sizes["ncells_own"], 4))
        gpu.get_state_local(qo)
        t = torch.tensor(np.concatenate([(vol_own[:, None] * qo).sum(axis=0), [float(np.isfinite(qo).all())]]), dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t)
        t = t.cpu().numpy()
        return t[:4], bool(t[4] == world)
    tot0, _ = conserved_totals()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    dt = run.dt
    t_sim = 0.0
    # warm-up (untimed)
    if W:
        gpu.time_integration(t_sim, W, logs=False)
        t_sim += W * dt
    # ---- timed region: exactly K steps of the production path (state resident in HBM; at N=1 steps 2..K replay a
    # CUDA graph), device time from CUDA events on the library's stream recorded around the K steps inside
    # fvs2d_gpu_time_integration, max over ranks
    with ClockSampler(local_rank) as clk:
        barrier()
        w0 = time.perf_counter()
        gpu.time_integration(t_sim, K, logs=False)
        barrier()
        wall = time.perf_counter() - w0
        tm = gpu.last_timing()
        t_sim += K * dt
    # ---- what the timed region left behind: finite, and conservative (interior fluxes cancel bit for bit -- the change of
    # the totals is the boundary flux of a vortex far from the boundary plus summation round-off)
    tot1, finite = conserved_totals()
    state_check = {"finite": finite, "steps_from_initial_state": W + K,
                   "conserved_totals_rel_drift": [float(x) for x in np.abs(tot1 - tot0) / np.abs(tot0).max()]}
    # ---- sustained: the same call repeated for >= --sustain-s seconds (the shipped examples run 4 000-50 000 steps; the
    # fp64-heavy kernels reach the power cap of this pool after ~100 ms and the SM clock settles lower)
    sustained = None
    if args.sustain_s > 0:
        ks = max(K, 20)
        # (every rank must make the same number of calls: the count comes from the max-over-ranks time, not the local one)
        reps = max(1, int(np.ceil(args.sustain_s / (max(max_over_ranks(tm["total_ms"]), 1e-3) * 1e-3 * ks / K))))
        with ClockSampler(local_rank) as clk_s:
            barrier()
            ms_s = 0.0
            for r in range(reps):
                gpu.time_integration(t_sim, ks, logs=False)
                ms_s += gpu.last_timing()["total_ms"]
                t_sim += ks * dt
            barrier()
        ms_s = max_over_ranks(ms_s)
        sustained = {"value": ncells * 4 * ks * reps / (ms_s * 1e-3), "ms_per_step": ms_s / (ks * reps), "steps": ks * reps,
                     "seconds": ms_s * 1e-3, "clocks": clk_s.summary()}
    # ---- second pass of K steps with a CUDA event pair around every kernel launch (no graph): the average launch
    # durations of pass A / pass B for the roofline.  It starts after a short idle: on this pool the fp64-heavy
    # pass B trips sw_power_cap after ~100 ms of sustained load, so a second pass run back to back would time the
    # kernels at lower clocks than the timed pass itself saw.  Its own clock samples are reported next to it.
    gpu.set_option("timing", 1)
    barrier()
    time.sleep(2.0)
    with ClockSampler(local_rank) as clk2:
        barrier()
        gpu.time_integration(t_sim, K, logs=False)
        barrier()
    tk = gpu.last_timing()
    gpu.set_option("timing", 0)
    t_sim += K * dt
    dev_ms = max_over_ranks(tm["total_ms"])
    value = ncells * 4 * K / (dev_ms * 1e-3)
    flux_ms = max_over_ranks(tk["flux_ms"]) / (4 * K)   # pass-B kernel time per stage (N>1: interior + boundary launch)
    grad_ms = max_over_ranks(tk["grad_ms"]) / (4 * K)   # pass-A kernel time per stage
    launches = tm["launches"]

    # ---- end to end through the C-ABI with HOST buffers: every step uploads cvar(4,ncells) from pinned
    # host memory, runs one time step, downloads cvar and the 4 residual norms
    e2e = None
    if not args.no_e2e:
        # N=1: the reference-facing global arrays cvar(4,ncells) in the original numbering;
        # N>1: every rank moves the cells it owns (fvs2d_gpu_{set,get}_state_local)
        n_mine = ncells if world == 1 else sizes["ncells_own"]
        q_host = torch.empty((n_mine, 4), dtype=torch.float64).pin_memory()
        put = gpu.set_state if world == 1 else gpu.set_state_local
        get = gpu.get_state if world == 1 else gpu.get_state_local
        get(q_host)
        put(q_host)                                   # untimed: first-use set-up of the copy / NCCL paths
        gpu.time_integration(t_sim, 1, logs=True)
        put(q_host)
        ke = min(K, 5)
        barrier()
        e0 = time.perf_counter()
        for s in range(ke):
            put(q_host)
            gpu.time_integration(t_sim + s * dt, 1, logs=True)
            get(q_host)
        barrier()
        e_s = max_over_ranks(time.perf_counter() - e0)
        # amortised variant: the seam as the reference calls it (one call per save interval)
        barrier()
        a0 = time.perf_counter()
        put(q_host)
        gpu.time_integration(t_sim, K, logs=True)
        get(q_host)
        barrier()
        a_s = max_over_ranks(time.perf_counter() - a0)
        e2e = {"value": ncells * 4 * ke / e_s, "unit": "cell-stage updates/s", "h2d_bytes_per_step": ncells * 32,
               "d2h_bytes_per_step": ncells * 32 + (32 + 16 * 8) * world, "steps": ke,
               "note": "per step: fvs2d_gpu_set_state(pinned host cvar) + fvs2d_gpu_time_integration(1 step, logs) + fvs2d_gpu_get_state"
                       + ("" if world == 1 else " (per rank: its owned cells, *_state_local)"),
               "amortized_value": ncells * 4 * K / a_s,
               "amortized_note": f"one seam call as the reference makes it: set_state + time_integration({K} steps) + get_state"}

    gpu.close()
    other = {}
    if world == 1 and args.workload == "c4" and args.scale == 1.0 and not args.no_extra:
        # BASELINE configs[2] (C3, the single-GPU triangle case) and configs[1] (C2, NACA 65 k cells: latency-bound,
        # LSQ-nn + Venkatakrishnan, steady SSPRK) measured in the same run under the same rules
        pk, _ = load_peaks()
        for key, wname, kk in (("c3", "c3", K), ("c2", "naca", 100 * K)):
            meshx, runx, descx, (ax, bx) = make_workload(wname, 1)
            gx = solver.Fvs2dGpu(runx.to_config(1), device=local_rank)
            gx.set_option("fuse", -1)               # one kernel per stage where it fits three CTAs per SM (C3: yes; C2: limiter, no)
            gx.set_mesh(meshx)
            ncx = meshx.ncells
            del meshx
            gx.initialize_solution()
            time.sleep(2.0)                          # same power state as a fresh timed pass (see above)
            gx.time_integration(0.0, max(W, 3), logs=False)
            torch.cuda.synchronize()
            passes = []
            for rep in range(3):                     # median of three production passes of kk steps each,
                time.sleep(1.0)                      # each started from an idle GPU like the headline's timed pass
                gx.time_integration((W + rep * kk) * runx.dt, kk, logs=False)
                passes.append(gx.last_timing())
            tv = sorted(passes, key=lambda t: t["total_ms"])[1]
            gx.set_option("timing", 1)
            time.sleep(2.0)
            gx.time_integration((W + 3 * kk) * runx.dt, kk, logs=False)
            tt = gx.last_timing()
            gx.close()
            other[key] = {"workload": descx, "steps": kk, "passes_ms_per_step": [round(t["total_ms"] / kk, 4) for t in passes], "value": ncx * 4 * kk / (tv["total_ms"] * 1e-3),
                          "ms_per_step": tv["total_ms"] / kk,
                          "pass_b_avg_launch_ms": tt["flux_ms"] / (4 * kk), "pass_a_avg_launch_ms": tt["grad_ms"] / (4 * kk),
                          "pass_b_roofline_frac": bx * ncx / (tt["flux_ms"] / (4 * kk) * 1e-3) / 1e9 / pk,
                          "stage_roofline_frac": (ax + bx) * ncx * 4 * kk / (tv["total_ms"] * 1e-3) / 1e9 / pk}
            if tt["grad_ms"] == 0.0 and tt["flux_ms"] > 0.0:
                # k_stage_fused2 ran: no pass A; "pass_b_*" is the fused stage kernel, which moves neither the gradients
                # (64 B written + 64 B read per cell) nor the primitive state a second time (32 B)
                fb = ax + bx - 160.0
                other[key].update({"fused_stage_kernel": True, "fused_alg_bytes_per_cell": fb,
                                   "pass_b_roofline_frac": fb * ncx / (tt["flux_ms"] / (4 * kk) * 1e-3) / 1e9 / pk})
        other["c2"]["note"] = "65 536 cells fit in L2 and one step is ~10 dependent launches of a few us: launch/latency-bound, not HBM-bound"
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return

    peak, peak_src = load_peaks()
    # cells per rank for the per-kernel figures: the mean (the ranks' chunks have equal estimated cost, not equal counts,
    # on mixed meshes; the kernel time is the max over ranks)
    n_own = ncells / world
    roof = {"bound": "hbm", "kernel": "k_flux_pipe (pass B: face-flux gather + residual + RK update; persistent TMA/cp.async smem pipeline)",
            "achieved": bB * n_own / (flux_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "peak_source": peak_src,
            "alg_bytes_per_cell": bB, "avg_launch_ms": flux_ms, "traffic": None,
            "measured": "CUDA event pair around every launch on the library stream, second pass of the same K steps started after a 2 s idle",
            "clocks": clk2.summary()}
    fused_ran = grad_ms == 0.0 and flux_ms > 0.0
    if fused_ran:
        # the default schedule: k_stage_fused, no pass A; its algorithmic bytes are B_alg - 160 (no gradient write + read,
        # state read once).  avg_launch_ms is the kernel time per STAGE (a stage is one launch on triangle meshes, two on
        # mixed meshes: triangle tiles at three CTAs per SM, then the tiles with quadrilaterals at two)
        fb = bA + bB - 160.0
        roof.update({"kernel": "k_stage_fused (one kernel per RK stage: gradients rebuilt in shared memory inside the persistent TMA/cp.async "
                               "pass-B pipeline; on several ranks the halo exchange is stores to peer memory from the same kernel)",
                     "alg_bytes_per_cell": fb, "achieved": fb * n_own / (flux_ms * 1e-3) / 1e9})
    # DRAM traffic per stage of the same kernel(s) from the committed ncu --set full capture of this workload (ncu cannot
    # run inside a timed bench; the file names the capture and the commit it was taken at)
    tpath = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if world == 1 and args.scale == 1.0 and os.path.exists(tpath):
        tr = json.load(open(tpath)).get(("fused:" if fused_ran else "twopass:") + args.workload)
        if tr:
            roof["traffic"] = tr["dram_bytes_per_stage"] / 1e9
            roof["traffic_unit"] = "GB per stage (ncu dram__bytes_read.sum + dram__bytes_write.sum), " + tr["source"]
            roof["alg_GB_per_stage"] = roof["alg_bytes_per_cell"] * n_own / 1e9
    roof["frac"] = roof["achieved"] / peak
    stage = {"alg_bytes_per_cell_stage": bA + bB, "achieved_GBs": (bA + bB) * ncells * 4 * K / (dev_ms * 1e-3) / 1e9}
    stage["frac"] = stage["achieved_GBs"] / (peak * world)
    gradk = {"kernel": "k_gradient (pass A)", "alg_bytes_per_cell": bA, "avg_launch_ms": grad_ms,
             "achieved_GBs": bA * n_own / (grad_ms * 1e-3) / 1e9 if grad_ms > 0 else None}
    line = {"metric": metric, "value": value, "unit": "cell-stage updates/s", "n_gpus": ngpus, "steps": K, "warmup": W,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "strong" if args.mesh_for_gpus and args.mesh_for_gpus != ngpus else "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic" if args.workload in ("c3", "c4") or args.workload.startswith("x:") else "the reference's example mesh, freestream / vortex initial state",
            "config": {"workload": desc, "ncells": ncells, "ncells_per_gpu": n_own, "ncells_rank0": sizes["ncells_own"], "dt": dt, "l2": "inputs larger than L2 "
                       f"({scal['device_bytes'] / 1e9:.2f} GB resident per GPU vs 126 MB L2)", "parallelism": f"dd{ngpus}",
                       "cpu_binding": (f"rank 0 bound to {len(cpu_binding)} CPUs next to its GPU (NVML ideal-CPU mask)" if cpu_binding else "none"),
                       "setup_s": round(t_setup, 1),
                       "setup_breakdown_s": {"synthetic_mesh_generator": round(t_meshgen, 1), "set_mesh_and_initial_condition": round(t_lib, 1)}},
            "roofline": roof, "stage_roofline": stage, "gradient_kernel": gradk,
            "wall_ms_per_step": wall * 1e3 / K, "gpu_launches": launches, "clocks": clk.summary(), "e2e": e2e,
            "parity": parity, "state_check": state_check, "sustained": sustained}
    if other:
        line["other_configs"] = other
    if not args.no_cpu_baseline:
        # ~10 s of single-thread CPU work on the 1/16-size sibling (540 k cells x 24 RK4 steps at ~6.5 M cell-stages/s)
        cb, _ = cpu_baseline(args.workload, run, 24 if args.workload in ("c3", "c4") else 300)
        line["cpu_baseline"] = cb
    print(json.dumps(line))


if __name__ == "__main__":
    main()
