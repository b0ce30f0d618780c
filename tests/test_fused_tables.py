"""CPU: the tables the one-kernel-per-stage variant (k_stage_fused, option "fuse") consumes are emulated tile by
tile in numpy -- ring-1 / ring-2 staging, stencil slots, coefficient rows, packed face words, edge slots -- following the
kernel's data flow (phase 1: gradients of tile + ring 1 from the staged slots; phase 2: both sides of every face from
the slots; -R/vol), and the result must be the oracle's residual.  No GPU, no product arithmetic: this checks the host
side of the fused path and documents what the kernel expects."""
import numpy as np
import pytest

from conftest import run_input  # noqa: F401

TILE = 128


def _emulate_residual(mesh, cfg, orc, time, rank=0, nranks=1):
    from fvs2d_b200 import capi, solver
    from oracle import oracle as om
    solver.host_build(cfg, mesh, rank, nranks)
    A = capi.mesh_array
    info = A("fz_info")
    assert info[0] == 1, "fused tables unusable for this mesh"
    W = int(info[1])
    F0 = 1 if cfg.grad_method in (1, 2) else 0
    orig = A("orig_id")                                # owned cells, then ghosts
    sz = np.zeros(10, dtype=np.int32)
    capi.lib().fvs2d_gpu_sizes(capi.ptr(sz))
    n_own, n_loc = int(sz[7]), int(sz[8])
    assert len(orig) == n_loc
    np_ = (n_loc + 31) // 32 * 32
    gc = A("fz_gc").reshape(W + F0, np_, 2)
    hdr = A("tile_hdr").reshape(-1, 8)
    fz = A("fz_hdr").reshape(-1, 8)[:, :4]
    hc_idx, he_idx, h2_idx = A("tile_hc_idx"), A("tile_he_idx"), A("fz_h2_idx")
    gslot, t_pack, t_bf = A("fz_gslot"), A("t_pack"), A("t_bf")
    bf_type, bf_edge = A("bf_type"), A("bf_edge")
    ex, ey, ea, enx, eny = A("lex"), A("ley"), A("lea"), A("lenx"), A("leny")
    xc, yc, vol = A("lxc"), A("lyc"), A("lvol")
    p = orc.array("pvar").reshape(-1, 4)[orig]          # primitive state in the library's cell order
    resid = np.zeros((n_own, 4))
    n2_seen = 0
    assert info[5] == 1, "tables of the second fused variant missing"
    pack2, hf_all = A("fz_pack2"), A("fz_hf")
    fz8 = A("fz_hdr").reshape(-1, 8)
    for t in range(len(hdr)):
        es, ne, hp, n1, ep, nhe, fbase, fw = (int(x) for x in hdr[t])
        h2p, n2, gsb, gw = (int(x) for x in fz[t])
        n2_seen += n2
        c0 = t * TILE
        ncell = min(TILE, n_own - c0)
        tw = (TILE + n1 + 7) & ~7
        # what the producer stages: slot -> cell id (own | ring 1 | ring 2), edge slot -> local edge id
        ids = np.full(TILE + n1 + n2, -1, dtype=np.int64)
        ids[:ncell] = np.arange(c0, c0 + ncell)
        ids[TILE:TILE + n1] = hc_idx[hp:hp + n1]
        ids[TILE + n1:] = h2_idx[h2p:h2p + n2]
        eids = np.concatenate([np.arange(es, es + ne), he_idx[ep:ep + nhe]])
        tab = gslot[gsb:gsb + gw * tw].reshape(gw, tw).astype(np.int64)
        # published face states: the face words point back at each other / at the list of tile/ring-1 faces
        hfp, nhf = int(fz8[t, 4]), int(fz8[t, 5])
        assert hfp % 4 == 0 and nhf <= int(info[6])
        seen = set()
        for j in range(ncell):
            for k in range(fw):
                w1, w2 = int(t_pack[fbase + k * TILE + j]), int(pack2[fbase + k * TILE + j])
                ns, code = w1 & 0xFFFF, w2 & 0xFFFF
                assert (w2 >> 31) == (w1 >> 31)
                if ns >= 0xFFFE:
                    assert code == ns
                    continue
                assert ((w2 >> 16) & 0xFFF) == ((w1 >> 16) & 0x7FFF)
                if ns < TILE:
                    kr = (w2 >> 28) & 3
                    back = int(pack2[fbase + kr * TILE + ns])
                    assert code == ns and (back & 0xFFFF) == j and ((back >> 16) & 0xFFF) == ((w2 >> 16) & 0xFFF)
                    assert ((back >> 28) & 3) == k and (back >> 31) != (w2 >> 31)
                else:
                    e = code - TILE
                    assert 0 <= e < nhf and e not in seen
                    seen.add(e)
                    assert int(hf_all[hfp + e]) == (ns - TILE) | (((w1 >> 16) & 0x7FFF) << 16)
        assert len(seen) == nhf
        # phase 1: gradients of the columns of tile + ring 1
        cols = np.concatenate([np.arange(ncell), np.arange(TILE, TILE + n1)])
        cid = ids[cols]
        pc = p[cid]
        if F0:
            gx, gy = gc[0, cid, 0][:, None] * pc, gc[0, cid, 1][:, None] * pc
        else:
            gx, gy = np.zeros_like(pc), np.zeros_like(pc)
        for k in range(gw):
            js = ids[tab[k, cols]]
            assert (js >= 0).all(), "stencil slot points at an unstaged cell"
            d = p[js] if F0 else p[js] - pc
            gx += gc[k + F0, cid, 0][:, None] * d
            gy += gc[k + F0, cid, 1][:, None] * d
        G = {int(c): (gx[i], gy[i]) for i, c in enumerate(cols)}
        # phase 2: faces of the own cells
        for j in range(ncell):
            acc = np.zeros(4)
            words = [int(t_pack[fbase + k * TILE + j]) for k in range(fw)]
            order = [k for k in range(fw) if (words[k] & 0xFFFF) < 0xFFFE] + [k for k in range(fw) if (words[k] & 0xFFFF) == 0xFFFF]
            for k in order:
                pk = words[k]
                ns, eslot, c2flag = pk & 0xFFFF, (pk >> 16) & 0x7FFF, pk >> 31
                le = int(eids[eslot])
                fx, fy, af, nx, ny = ex[le], ey[le], ea[le], enx[le], eny[le]

                def recon(slot):
                    i = int(ids[slot])
                    gxs, gys = G[slot]
                    return p[i] + (fx - xc[i]) * gxs + (fy - yc[i]) * gys
                if ns == 0xFFFF:
                    b = int(t_bf[fbase + k * TILE + j])
                    assert bf_edge[b] == le
                    sL = recon(j)
                    if bf_type[b] == 2:
                        un = sL[1] * nx + sL[2] * ny
                        sR = np.array([sL[0], sL[1] - 2 * un * nx, sL[2] - 2 * un * ny, sL[3]])
                    elif bf_type[b] == 1:
                        sR = np.array(cfg.pvar_inf[:])
                    else:
                        sR = om.vortex_point(cfg, time, fx, fy)
                    fl, _ = om.roe_flux(cfg.gamma, sL, sR, nx, ny)
                    acc += fl * af
                else:
                    assert ns < TILE + n1, "face neighbour outside tile + ring 1"
                    sl_, sr_ = (j, ns) if c2flag == 0 else (ns, j)
                    fl, _ = om.roe_flux(cfg.gamma, recon(sl_), recon(sr_), nx, ny)
                    acc += fl * af if c2flag == 0 else -fl * af
            resid[c0 + j] = -acc / vol[c0 + j]
    out = np.full((mesh.ncells, 4), np.nan)
    out[orig[:n_own]] = resid
    return out, n2_seen


@pytest.mark.parametrize("grad,stencil,mixed", [(1, "fn", True), (3, "fn", False), (3, "nn", True), (2, "fn", False)])
def test_fused_tables_reproduce_the_oracle_residual(grad, stencil, mixed):
    from fvs2d_b200 import config, meshgen
    from oracle.oracle import Oracle
    mesh = meshgen.vortex_mixed_mesh(32) if mixed else meshgen.vortex_tri_mesh(30)
    cfg = config.RunInput(grad_cellcntr_imethd=grad, grad_cellcntr_lsq_nghbr=stencil, lvortex=True, dt=0.01).to_config()
    orc = Oracle(mesh, cfg)
    orc.initialize_solution()
    r_o = orc.compute_residual(0.3).copy()
    r_e, n2 = _emulate_residual(mesh, cfg, orc, 0.3)
    assert n2 > 0
    scale = np.abs(r_o).max(axis=0)
    assert (np.abs(r_e - r_o) / scale).max() < 1e-11


def test_fused_tables_slip_wall_and_freestream(naca_mesh):
    """the NACA o-grid: quads, slip wall + freestream boundaries, GGCB."""
    from fvs2d_b200 import config
    from oracle.oracle import Oracle
    r = run_input("naca")
    r.grad_cellcntr_imethd = 1
    r.grad_limiter_imethd = 0
    cfg = r.to_config()
    # a quarter of the mesh is enough for the check and keeps the pure-Python face loop short
    orc = Oracle(naca_mesh, cfg)
    orc.initialize_solution()
    rng = np.random.default_rng(3)
    q = orc.cvar.copy()
    q *= 1.0 + 0.01 * rng.standard_normal(q.shape)       # a non-uniform state, so the residual is not round-off
    orc.set_state(q)
    r_o = orc.compute_residual(0.0).copy()
    r_e, _ = _emulate_residual(naca_mesh, cfg, orc, 0.0)
    scale = np.abs(r_o).max(axis=0)
    assert (np.abs(r_e - r_o) / scale).max() < 1e-11


@pytest.mark.parametrize("grad,stencil,nranks", [(1, "fn", 2), (3, "nn", 3)])
def test_fused_tables_with_deep_ghost_layers(grad, stencil, nranks, monkeypatch):
    """several ranks: with the extra ghost layer every tile's rings are local, ring-1 ghosts carry their gradient
    operator, and the ranks' emulated residuals tile the oracle's; the halo plans still match pairwise."""
    from fvs2d_b200 import capi, config, meshgen, solver
    from oracle.oracle import Oracle
    monkeypatch.setenv("FVS2D_DEEP_GHOSTS", "1")
    mesh = meshgen.vortex_mixed_mesh(32)
    cfg = config.RunInput(grad_cellcntr_imethd=grad, grad_cellcntr_lsq_nghbr=stencil, lvortex=True, dt=0.01).to_config()
    orc = Oracle(mesh, cfg)
    orc.initialize_solution()
    r_o = orc.compute_residual(0.3).copy()
    total = np.full_like(r_o, np.nan)
    plans = []
    for r in range(nranks):
        r_e, _ = _emulate_residual(mesh, cfg, orc, 0.3, r, nranks)
        own = ~np.isnan(r_e[:, 0])
        assert np.isnan(total[own]).all()
        total[own] = r_e[own]
        A = capi.mesh_array
        ti, tb = A("fz_tile_int"), A("fz_tile_bnd")
        assert len(ti) + len(tb) == len(A("tile_hdr")) // 8 and len(tb) > 0
        plans.append(dict(loc2new=A("loc2new").copy(), peers=A("peers").copy(), send_ptr=A("send_ptr").copy(), send_idx=A("send_idx").copy(),
                          recv_begin=A("recv_begin").copy(), recv_count=A("recv_count").copy()))
    assert not np.isnan(total).any()
    scale = np.abs(r_o).max(axis=0)
    assert (np.abs(total - r_o) / scale).max() < 1e-11
    for r, p in enumerate(plans):
        for k, peer in enumerate(p["peers"]):
            q = plans[peer]
            kk = list(q["peers"]).index(r)
            sent = p["loc2new"][p["send_idx"][p["send_ptr"][k]:p["send_ptr"][k + 1]]]
            recv = q["loc2new"][q["recv_begin"][kk]:q["recv_begin"][kk] + q["recv_count"][kk]]
            assert np.array_equal(sent, recv)
