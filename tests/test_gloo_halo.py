"""CPU, world_size 2 (gloo): the N>1 host path -- partition + halo plan built by the library -- driven by real
multi-process communication.  Each rank owns a contiguous Hilbert chunk; owned values are exchanged with
send/recv per the plan (the same lists the NCCL path packs from / receives into) and every ghost must end up
holding its owner's value; a 'residual-like' stencil reduction over owned cells then equals the serial one."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fvs2d_b200 import capi, config, meshgen, solver
        mesh = meshgen.vortex_mixed_mesh(32)
        cfg = config.RunInput(grad_cellcntr_imethd=2, lvortex=True).to_config()
        solver.host_build(cfg, mesh, rank, world)
        A = capi.mesh_array
        s = np.zeros(10, dtype=np.int32); capi.lib().fvs2d_gpu_sizes(capi.ptr(s))
        n_own, n_loc = int(s[7]), int(s[8])
        orig = A("orig_id")
        peers, send_ptr, send_idx = A("peers"), A("send_ptr"), A("send_idx")
        recv_begin, recv_count = A("recv_begin"), A("recv_count")
        # local SoA array, 3 variables, pitch n_loc: owned = f(original id), ghosts poisoned
        val = lambda o: np.stack([np.sin(0.01 * o), o.astype(np.float64), 1.0 / (1.0 + o)], 0)
        a = np.full((3, n_loc), np.nan)
        a[:, :n_own] = val(orig[:n_own])
        reqs, bufs = [], []
        for k, peer in enumerate(peers):
            sl = send_idx[send_ptr[k]:send_ptr[k + 1]]
            sb = torch.from_numpy(np.ascontiguousarray(a[:, sl]))          # pack: [var][cells], like k_pack
            rb = torch.empty((3, int(recv_count[k])), dtype=torch.float64)
            bufs.append((k, rb))
            reqs.append(dist.isend(sb, int(peer)))
            reqs.append(dist.irecv(rb, int(peer)))
        for r in reqs:
            r.wait()
        for k, rb in bufs:
            a[:, recv_begin[k]:recv_begin[k] + recv_count[k]] = rb.numpy()  # ghosts of one peer are one contiguous run
        ok_ghosts = bool(np.array_equal(a, val(orig[:n_loc])))
        # stencil reduction over owned cells through local ids == the same through original ids
        g_off, g_idx, g_cx = A("g_off"), A("g_idx"), A("g_cx")
        acc = np.zeros(n_own)
        for i in range(n_own):
            sl, lane = i >> 5, i & 31
            for kk in range((g_off[sl + 1] - g_off[sl]) >> 5):
                e = g_off[sl] + 32 * kk + lane
                acc[i] += g_cx[e] * a[1, g_idx[e]]
        # the operator in the pre-processing's own numbering: a partition-local build numbers this rank's submesh
        ptr, idx, cx, sub = A("grad_ptr"), A("grad_idx"), A("grad_cx"), A("sub_orig")
        assert len(sub) > 0 and len(sub) < mesh.ncells, "several ranks: the pre-processing must be partition-local"
        mc = np.searchsorted(sub, orig[:n_own])
        assert np.array_equal(sub[mc], orig[:n_own])
        ref = np.array([sum(cx[t] * float(sub[idx[t]]) for t in range(ptr[o], ptr[o + 1])) for o in mc])
        ok_stencil = bool(np.allclose(acc, ref, rtol=1e-13, atol=1e-9))
        res = torch.tensor([float(ok_ghosts), float(ok_stencil), float(n_own)], dtype=torch.float64)
        gathered = [torch.zeros(3, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(gathered, res)
        if rank == 0:
            out.put([g.tolist() for g in gathered])
    finally:
        dist.destroy_process_group()


def test_halo_exchange_world_size_2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = out.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[0] == 1.0 and r[1] == 1.0 for r in res), res
    from fvs2d_b200 import meshgen
    assert sum(int(r[2]) for r in res) == meshgen.vortex_mixed_mesh(32).ncells
