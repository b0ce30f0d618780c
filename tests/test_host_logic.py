"""CPU: the C-ABI library loads and exports every declared symbol; the host pre-processing inside it
(connectivity, geometry, gradient operators, Hilbert layout, tiles, partition + halo plan) is checked
against the oracle's literal restatement of grid_data and against structural invariants.  No compute calls."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, run_input


def test_library_exports_every_declared_symbol():
    from fvs2d_b200 import capi
    hdr = open(os.path.join(ROOT, "include", "fvs2d_gpu.h")).read()
    declared = set(re.findall(r"\b(fvs2d_(?:gpu|host)_\w+)\s*\(", hdr))
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)
    L = ctypes.CDLL(capi.LIB_PATH)
    for s in declared:
        assert hasattr(L, s), s


def test_no_gpu_means_loud_failure():
    """without a CUDA device every computing entry point must fail (no CPU fallback)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from fvs2d_b200 import capi, config
    cfg = config.RunInput().to_config()
    L = capi.lib()
    assert L.fvs2d_gpu_init(ctypes.byref(cfg), 0) != 0
    assert b"no CUDA device" in L.fvs2d_gpu_last_error()
    assert L.fvs2d_gpu_time_integration(0.0, 1, None, None, None) != 0
    assert L.fvs2d_gpu_compute_residual(0.0, None, None) != 0


@pytest.mark.parametrize("name", ["vortex", "naca"])
def test_host_mesh_matches_oracle_on_reference_meshes(name, vortex_mesh, naca_mesh):
    from fvs2d_b200 import capi, solver
    from oracle.oracle import Oracle
    mesh = vortex_mesh if name == "vortex" else naca_mesh
    cfg = run_input(name).to_config()
    orc = Oracle(mesh, cfg)
    solver.host_build(cfg, mesh)
    for nm in ["xc", "yc", "vol", "ex", "ey", "ea", "enx", "eny", "en1", "en2", "ec1", "ec2", "cedge", "nghbre", "cell_intr", "b_edge"]:
        assert np.array_equal(capi.mesh_array(nm), orc.array(nm)), nm
    # LSQ operator: same stencil (incl. the 8-NN augmentation of boundary cells, fn) and coefficients
    assert np.array_equal(capi.mesh_array("grad_ptr"), orc.array("lsq_ptr"))
    assert np.array_equal(capi.mesh_array("grad_idx"), orc.array("lsq_cell"))
    oc, ow = orc.array("lsq_coef").reshape(-1, 2), orc.array("lsq_w")
    np.testing.assert_allclose(capi.mesh_array("grad_cx"), oc[:, 0] * ow, rtol=1e-13, atol=1e-300)
    np.testing.assert_allclose(capi.mesh_array("grad_cy"), oc[:, 1] * ow, rtol=1e-13, atol=1e-300)


@pytest.mark.parametrize("grad,stencil", [(1, "fn"), (2, "fn"), (3, "fn"), (3, "nn")])
def test_gradient_operator_applies_like_oracle(grad, stencil):
    """apply the library's generic sparse operator (numpy) to a random field and compare with the oracle's grad."""
    from fvs2d_b200 import capi, config, meshgen, solver
    from oracle.oracle import Oracle
    mesh = meshgen.vortex_mixed_mesh(24)
    cfg = config.RunInput(grad_cellcntr_imethd=grad, grad_cellcntr_lsq_nghbr=stencil, grad_cellcntr_lsq_pow=1.0 if grad == 3 else 0.0,
                          lvortex=True).to_config()
    orc = Oracle(mesh, cfg)
    orc.initialize_solution()
    orc.compute_residual(0.1)
    p = orc.array("pvar").reshape(-1, 4)
    g_o = orc.array("grad").reshape(2, -1, 4)
    solver.host_build(cfg, mesh)
    ptr, idx = capi.mesh_array("grad_ptr"), capi.mesh_array("grad_idx")
    cx, cy = capi.mesh_array("grad_cx"), capi.mesh_array("grad_cy")
    rows = np.repeat(np.arange(mesh.ncells), np.diff(ptr))
    if grad == 3:
        d = p[idx] - p[rows]
        gx = np.zeros_like(p); gy = np.zeros_like(p)
        np.add.at(gx, rows, cx[:, None] * d); np.add.at(gy, rows, cy[:, None] * d)
    else:
        gx = capi.mesh_array("grad_c0x")[:, None] * p; gy = capi.mesh_array("grad_c0y")[:, None] * p
        np.add.at(gx, rows, cx[:, None] * p[idx]); np.add.at(gy, rows, cy[:, None] * p[idx])
    scale = np.abs(g_o).max()
    assert np.abs(gx - g_o[0]).max() / scale < 1e-11 and np.abs(gy - g_o[1]).max() / scale < 1e-11


def _layout_checks(mesh, cfg, rank, nranks):
    from fvs2d_b200 import capi, solver
    solver.host_build(cfg, mesh, rank, nranks)
    A = capi.mesh_array
    perm, loc2new, orig = A("perm"), A("loc2new"), A("orig_id")
    nc = mesh.ncells
    assert np.array_equal(np.sort(perm), np.arange(nc))  # a permutation
    assert np.array_equal(perm[loc2new], orig)
    f_off, f_nbr, f_edge = A("f_off"), A("f_nbr"), A("f_edge")
    s = np.zeros(10, dtype=np.int32); capi.lib().fvs2d_gpu_sizes(capi.ptr(s))
    n_own, n_loc = int(s[7]), int(s[8])
    ec1, ec2, cedge, nghbre = A("ec1"), A("ec2"), A("cedge"), A("nghbre")
    cptr, _ = mesh.csr()
    lex, ex = A("lex"), A("ex")
    PAD = np.iinfo(np.int32).min
    # every owned cell lists exactly its faces, neighbours translate back to the original ids
    for i in np.random.default_rng(1).choice(n_own, size=min(n_own, 400), replace=False):
        sl, lane = i >> 5, i & 31
        w = (f_off[sl + 1] - f_off[sl]) >> 5
        ent = [f_off[sl] + 32 * k + lane for k in range(w)]
        nb = [f_nbr[e] for e in ent if f_nbr[e] != PAD]
        o = orig[i]
        true_nb = sorted(int(x) for x in nghbre[cptr[o]:cptr[o + 1]] if x >= 0)
        got_nb = sorted(int(orig[x]) for x in nb if x >= 0)
        assert true_nb == got_nb
        assert sum(1 for x in nb if x < 0) == sum(1 for x in nghbre[cptr[o]:cptr[o + 1]] if x < 0)
        # edge geometry + orientation flag
        for e in ent:
            if f_nbr[e] == PAD:
                continue
            le, flag = f_edge[e] >> 1, f_edge[e] & 1
            cands = [je for je in cedge[cptr[o]:cptr[o + 1]] if ex[je] == lex[le]]
            assert cands, "edge geometry not found"
            assert any((ec1[je] == o) == (flag == 0) for je in cands)
    return n_own, n_loc


def test_layout_single_rank(vortex_mesh):
    from fvs2d_b200 import capi
    cfg = run_input("vortex").to_config()
    n_own, n_loc = _layout_checks(vortex_mesh, cfg, 0, 1)
    assert n_own == n_loc == vortex_mesh.ncells
    # tiles: every face entry resolves to the same neighbour through the packed slots
    A = capi.mesh_array
    hdr = A("tile_hdr").reshape(-1, 8) if "tile_hdr" in capi._INT_ARRAYS else None
    f_off, f_nbr, f_pack = A("f_off"), A("f_nbr"), A("f_pack")
    hc_ptr, hc_idx = A("tile_hc_ptr"), A("tile_hc_idx")
    PAD = np.iinfo(np.int32).min
    for i in range(0, n_own, 7):
        t, sl, lane = i // 128, i >> 5, i & 31
        for k in range((f_off[sl + 1] - f_off[sl]) >> 5):
            e = f_off[sl] + 32 * k + lane
            ns = int(f_pack[e]) & 0xFFFF
            if f_nbr[e] == PAD:
                assert ns == 0xFFFE
            elif f_nbr[e] < 0:
                assert ns == 0xFFFF
            elif ns < 128:
                assert f_nbr[e] == 128 * t + ns
            else:
                assert f_nbr[e] == hc_idx[hc_ptr[t] + ns - 128]


@pytest.mark.parametrize("nranks", [2, 3])
def test_partition_and_halo_plan(nranks):
    """contiguous Hilbert chunks; ghosts = everything an owned cell (or, deep layout, a face-neighbour ghost's gradient)
    reads; the send list of rank a for rank b
    is exactly rank b's ghost run owned by a, in the same order."""
    from fvs2d_b200 import capi, config, meshgen, solver
    mesh = meshgen.vortex_mixed_mesh(32)
    cfg = config.RunInput(grad_cellcntr_imethd=3, grad_cellcntr_lsq_nghbr="nn", lvortex=True).to_config()
    plans = []
    for r in range(nranks):
        solver.host_build(cfg, mesh, r, nranks)
        A = capi.mesh_array
        sz = np.zeros(10, dtype=np.int32); capi.lib().fvs2d_gpu_sizes(capi.ptr(sz))
        n_own, n_loc = int(sz[7]), int(sz[8])
        assert len(A("sub_orig")) < mesh.ncells, "several ranks: partition-local pre-processing"
        plans.append(dict(n_own=n_own, n_loc=n_loc, loc2new=A("loc2new").copy(), peers=A("peers").copy(), send_ptr=A("send_ptr").copy(),
                          send_idx=A("send_idx").copy(), recv_begin=A("recv_begin").copy(), recv_count=A("recv_count").copy(),
                          g_idx=A("g_idx").copy(), f_nbr=A("f_nbr").copy(), gh_idx=A("gh_idx").copy()))
    assert sum(p["n_own"] for p in plans) == mesh.ncells
    for r, p in enumerate(plans):
        assert p["g_idx"].max() < p["n_loc"] and p["f_nbr"].max() < p["n_loc"]  # stencil closure
        used = np.zeros(p["n_loc"], bool)
        used[p["g_idx"]] = True
        used[p["f_nbr"][p["f_nbr"] >= 0]] = True
        # default on several ranks for this scheme: the deep layout of the fused stage kernel -- the face-neighbour ghosts
        # carry their own gradient stencils (gh_idx), whose members are ghosts too
        assert len(p["gh_idx"]) > 0
        used[p["gh_idx"]] = True
        assert used[p["n_own"]:].all()  # no useless ghost
        for k, peer in enumerate(p["peers"]):
            q = plans[peer]
            kk = list(q["peers"]).index(r)
            sent_new = p["loc2new"][p["send_idx"][p["send_ptr"][k]:p["send_ptr"][k + 1]]]
            recv_new = q["loc2new"][q["recv_begin"][kk]:q["recv_begin"][kk] + q["recv_count"][kk]]
            assert np.array_equal(sent_new, recv_new)


LAYOUT_ARRAYS = ["orig_id", "loc2new", "f_off", "f_nbr", "f_edge", "lex", "ley", "lea", "lenx", "leny", "lxc", "lyc", "lvol", "g_off", "g_idx",
                 "g_cx", "g_cy", "bf_type", "bf_edge", "peers", "send_ptr", "send_idx", "recv_begin", "recv_count", "tile_hdr", "t_pack", "t_bf",
                 "tile_hc_idx", "tile_he_idx", "is_intr", "gh_ptr", "gh_idx", "fz_tile_int", "fz_tile_bnd", "fz_hdr", "fz_h2_idx",
                 "fz_gslot", "fz_pack2", "fz_hf", "fz_gc"]


@pytest.mark.parametrize("grad,stencil,nranks", [(1, "fn", 2), (1, "fn", 5), (2, "fn", 3), (3, "nn", 4)])
def test_partition_local_build_equals_whole_mesh_build(grad, stencil, nranks, monkeypatch):
    """SURVEY 8 row f1 on several ranks: a rank pre-processes only its Hilbert chunk plus two rings of node-adjacent cells
    cut out of the caller's global arrays (no global connectivity).  Everything that reaches the device -- local
    numbering, faces, local edges and their geometry, gradient operator, boundary faces, tiles, halo plan, the fused
    kernel's tables -- must be bit-identical to what the whole-mesh pre-processing (the round-1 path, still used for the
    least-squares stencil over face neighbours) gives for the same rank; the whole-mesh path itself is checked against the
    oracle's grid_data restatement above."""
    from fvs2d_b200 import capi, config, meshgen, solver
    mesh = meshgen.vortex_mixed_mesh(48)
    cfg = config.RunInput(grad_cellcntr_imethd=grad, grad_cellcntr_lsq_nghbr=stencil, lvortex=True).to_config()
    A = capi.mesh_array
    for rank in range(nranks):
        monkeypatch.delenv("FVS2D_GLOBAL_BUILD", raising=False)
        solver.host_build(cfg, mesh, rank, nranks)
        assert 0 < len(A("sub_orig")) < mesh.ncells
        sz = np.zeros(10, dtype=np.int32); capi.lib().fvs2d_gpu_sizes(capi.ptr(sz))
        part = {k: A(k).copy() for k in LAYOUT_ARRAYS}
        monkeypatch.setenv("FVS2D_GLOBAL_BUILD", "1")
        solver.host_build(cfg, mesh, rank, nranks)
        assert len(A("sub_orig")) == 0
        sz2 = np.zeros(10, dtype=np.int32); capi.lib().fvs2d_gpu_sizes(capi.ptr(sz2))
        assert np.array_equal(sz[7:], sz2[7:]) and sz[0] == sz2[0] and sz[1] == sz2[1]     # (the other counts are per-rank shares until reduced)
        for k in LAYOUT_ARRAYS:
            assert np.array_equal(part[k], A(k)), f"rank {rank}/{nranks}: {k} differs between the partition-local and the whole-mesh build"


def test_input_and_mesh_files_round_trip(tmp_path, vortex_mesh):
    from fvs2d_b200 import config, meshio
    r = run_input("naca")
    r.grad_limiter_imethd = 1
    config.write_input(str(tmp_path / "fvs2d.input"), r)
    r2 = config.read_input(str(tmp_path / "fvs2d.input"))
    assert r2 == r
    v = run_input("vortex")
    config.write_input(str(tmp_path / "fvs2d.input"), v)
    v2 = config.read_input(str(tmp_path / "fvs2d.input"))
    assert v2 == v and v2.lvortex and v2.vortex_inf == (1.0, 0.2, 0.0, 1.0)
    c = v2.to_config()
    assert c.umuscl_cst == 0.0 and c.recon == 2  # src/input.f90:252-254: recon 2 forces kappa = 0
    assert v2.nsubsteps() == [80] * 50
    v.ntimes, v.nsaves = 10, 4
    assert v.nsubsteps() == [3, 3, 3, 1]  # src/input.f90:131-135
    base = str(tmp_path / "m")
    meshio.write_mesh(base, vortex_mesh)
    m2 = meshio.read_mesh(base)
    assert np.array_equal(m2.node_xy, vortex_mesh.node_xy) and np.array_equal(m2.tri, vortex_mesh.tri)
    assert m2.bndry_type == vortex_mesh.bndry_type and np.array_equal(m2.bndry_cell[0], vortex_mesh.bndry_cell[0])
    with pytest.raises(ValueError):
        bad = run_input("vortex"); bad.flux_inviscd_imethd = 2; bad.to_config()
    with pytest.raises(ValueError):
        bad = run_input("vortex"); bad.grad_cellcntr_imethd = 4; bad.to_config()


def test_bad_meshes_are_rejected(vortex_mesh):
    """grid_data's stop conditions (src/grid_procs.f90:722-728): boundary count mismatch -> error, not a crash."""
    import copy
    from fvs2d_b200 import capi, solver
    m = copy.deepcopy(vortex_mesh)
    m.bndry_cell = [m.bndry_cell[0][:-1]]
    with pytest.raises(capi.Fvs2dError, match="boundary cells"):
        solver.host_build(run_input("vortex").to_config(), m)
    from fvs2d_b200 import meshgen
    with pytest.raises(ValueError):
        meshgen.make_mesh(8, 8, quad_band=(0, 4))  # quad in a domain corner
