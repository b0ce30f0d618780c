// layout.cpp -- see layout.hpp.
#include "layout.hpp"

#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include <omp.h>

namespace fvs2d {

namespace {
struct FaceTmp {
  int key_b;   // 0 interior, 1+ib boundary: the reference sums interior edges first, then the boundaries in .bc order
  int edge;    // original global edge id (ascending = the reference's accumulation order, src/residual.f90:66,111)
  int nbr;     // original neighbour id or -1
};
}  // namespace

std::string build_layout(const HostMesh &m, const GradOp &g, const CellNumbering &num, int rank, int nranks, Layout &L,
                         bool deep) {
  // `m` holds the cells this call may touch: the whole mesh (one rank, or the global build) or this rank's submesh
  // (extract_submesh).  m-cell ids index m's arrays; num.new_id maps them to the global Hilbert order, num.order lists
  // them by ascending Hilbert id, num.orig (or identity) gives the original ids the caller's arrays use.
  const int nc = num.nc_global, ncm = m.ncells;
  const double t_begin = omp_get_wtime();
  auto lap = [&](const char *what) { if (getenv("FVS2D_DEBUG")) fprintf(stderr, "[fvs2d]   build_layout: %-22s %.2f s\n", what, omp_get_wtime() - t_begin); };
  L = Layout();
  L.deep = deep && nranks > 1;
  deep = L.deep != 0;
  L.rank = rank; L.nranks = nranks; L.nc_global = nc;
  if (!num.orig) L.perm = num.order;  // whole mesh: new id -> original id (output path, verification)
  const std::vector<int> &order = num.order, &new_id = num.new_id;
  // rank boundaries on tile boundaries of the global order (tiles are sorted internally by hilbert_order)
  if ((int)num.cuts.size() != nranks + 1) return "build_layout: partition cuts missing";
  auto range_begin = [&](int r) { return r >= nranks ? nc : num.cuts[r]; };
  auto pos_of = [&](int newid) {  // first entry of `order` whose Hilbert id is >= newid
    return (int)(std::lower_bound(order.begin(), order.end(), newid, [&](int mc, int v) { return new_id[mc] < v; }) - order.begin());
  };
  const int b0 = range_begin(rank), b1 = range_begin(rank + 1);
  const int k0 = pos_of(b0), k1 = pos_of(b1);
  if (k1 - k0 != b1 - b0) return "build_layout: the mesh handed in does not hold all of this rank's cells";
  L.own_begin = b0;
  L.n_own = b1 - b0;
  L.g_form = g.form;
  auto owned = [&](int mc) { return new_id[mc] >= b0 && new_id[mc] < b1; };

  // ---- ghosts: everything an owned cell reads that it does not own
  std::vector<int> ghosts;   // m-cells, ascending Hilbert id
  std::vector<int> loc_of;   // m-cell -> local id (owned first, then ghosts), -1: not stored by this rank
  if (nranks > 1) {
    std::vector<unsigned char> mark(ncm, 0);  // bit 0: ghost, bit 1: face-neighbour ghost (concurrent writes store the same bits)
#pragma omp parallel for schedule(static)
    for (int k = k0; k < k1; k++) {
      const int o = order[k];
      for (int s = m.cptr[o]; s < m.cptr[o + 1]; s++) {
        const int j = m.nghbre[s];
        if (j >= 0 && !owned(j)) {
#pragma omp atomic write
          mark[j] = 3;
        }
      }
    }
#pragma omp parallel for schedule(static)
    for (int k = k0; k < k1; k++) {
      const int o = order[k];
      for (int64_t t = g.ptr[o]; t < g.ptr[o + 1]; t++) {
        const int j = g.idx[t];
        if (owned(j)) continue;
        unsigned char seen;
#pragma omp atomic read
        seen = mark[j];
        if (!seen) {
#pragma omp atomic write
          mark[j] = 1;
        }
      }
    }
    if (deep) {  // the stencils of the face-neighbour ghosts must be local too
      std::vector<int> g1;
      for (int i = 0; i < ncm; i++) if (mark[i] & 2) g1.push_back(i);
      for (int j : g1)
        for (int64_t t = g.ptr[j]; t < g.ptr[j + 1]; t++) {
          const int kk = g.idx[t];
          if (!owned(kk) && !mark[kk]) mark[kk] = 1;
        }
    }
    for (int i = 0; i < ncm; i++) if (mark[i]) ghosts.push_back(i);
    std::sort(ghosts.begin(), ghosts.end(), [&](int x, int y) { return new_id[x] < new_id[y]; });
    loc_of.assign(ncm, -1);
    for (int k = k0; k < k1; k++) loc_of[order[k]] = new_id[order[k]] - b0;
    for (size_t k = 0; k < ghosts.size(); k++) loc_of[ghosts[k]] = L.n_own + (int)k;
    if (deep) {
      // gradient operator of the face-neighbour ghosts (local ids; the members are local by construction)
      const int ng = (int)ghosts.size();
      L.gh_ptr.assign(ng + 1, 0);
      for (int k = 0; k < ng; k++) L.gh_ptr[k + 1] = L.gh_ptr[k] + ((mark[ghosts[k]] & 2) ? (int)(g.ptr[ghosts[k] + 1] - g.ptr[ghosts[k]]) : 0);
      L.gh_idx.resize(L.gh_ptr[ng]); L.gh_cx.resize(L.gh_ptr[ng]); L.gh_cy.resize(L.gh_ptr[ng]);
      L.gh_c0x.assign(ng, 0.0); L.gh_c0y.assign(ng, 0.0);
    }
  }
  L.n_loc = L.n_own + (int)ghosts.size();
  L.loc2new.resize(L.n_loc);
  std::vector<int> loc2m(L.n_loc);  // local id -> m-cell
  for (int i = 0; i < L.n_own; i++) { loc2m[i] = order[k0 + i]; L.loc2new[i] = b0 + i; }
  for (size_t k = 0; k < ghosts.size(); k++) { loc2m[L.n_own + k] = ghosts[k]; L.loc2new[L.n_own + k] = new_id[ghosts[k]]; }
  auto to_local = [&](int mc) -> int { return nranks > 1 ? loc_of[mc] : new_id[mc] - b0; };
  if (deep) {
    const int ng = (int)ghosts.size();
    int bad = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : bad)
    for (int k = 0; k < ng; k++) {
      const int n = L.gh_ptr[k + 1] - L.gh_ptr[k];
      if (n == 0) continue;
      const int o = ghosts[k];
      double cx[kMaxStencil], cy[kMaxStencil], c0x = 0, c0y = 0;
      if (grad_cell_coeffs(m, g, o, cx, cy, c0x, c0y) < 0) bad++;
      for (int e = 0; e < n; e++) {
        const int l = to_local(g.idx[g.ptr[o] + e]);
        if (l < 0) bad++;
        L.gh_idx[L.gh_ptr[k] + e] = l;
        L.gh_cx[L.gh_ptr[k] + e] = cx[e];
        L.gh_cy[L.gh_ptr[k] + e] = cy[e];
      }
      L.gh_c0x[k] = c0x; L.gh_c0y[k] = c0y;
    }
    if (bad) return "gradient_lsq: singular least-squares system";
  }

  lap("ghosts");
  // ---- per-cell data
  L.orig_id.resize(L.n_loc); L.xc.resize(L.n_loc); L.yc.resize(L.n_loc);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < L.n_loc; i++) {
    const int o = loc2m[i];
    L.orig_id[i] = num.orig ? num.orig[o] : o; L.xc[i] = m.xc[o]; L.yc[i] = m.yc[o];
  }
  L.vol.resize(L.n_own); L.is_intr.resize(L.n_own);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < L.n_own; i++) {
    const int o = loc2m[i];
    L.vol[i] = m.vol[o];
    bool intr = true;
    for (int s = m.cptr[o]; s < m.cptr[o + 1]; s++) intr = intr && m.nghbre[s] >= 0;
    L.is_intr[i] = intr;
  }

  // ---- local edge numbering by first touch (sequential: order matters)
  std::vector<int> edge_loc(m.nedges, -1);
  int ne_loc = 0, nbf = 0;
  L.ntiles = (L.n_own + kTile - 1) / kTile;
  L.tile_es.assign(L.ntiles, 0);
  L.tile_ne.assign(L.ntiles, 0);
  for (int i = 0; i < L.n_own; i++) {
    if (i % kTile == 0) {  // a tile's own-edge range starts 16-byte aligned (bulk-copy requirement)
      if (i > 0) L.tile_ne[i / kTile - 1] = ((ne_loc + 1) & ~1) - L.tile_es[i / kTile - 1];
      ne_loc = (ne_loc + 1) & ~1;
      L.tile_es[i / kTile] = ne_loc;
    }
    const int o = loc2m[i];
    for (int s = m.cptr[o]; s < m.cptr[o + 1]; s++) {
      const int je = m.cedge[s];
      if (edge_loc[je] < 0) edge_loc[je] = ne_loc++;
      nbf += m.nghbre[s] < 0;
    }
  }
  ne_loc = (ne_loc + 1) & ~1;
  if (L.ntiles) L.tile_ne[L.ntiles - 1] = ne_loc - L.tile_es[L.ntiles - 1];
  L.nedges = ne_loc;
  L.ex.assign(ne_loc, 0.0); L.ey.assign(ne_loc, 0.0); L.ea.assign(ne_loc, 0.0); L.enx.assign(ne_loc, 0.0); L.eny.assign(ne_loc, 0.0);
#pragma omp parallel for schedule(static)
  for (int je = 0; je < m.nedges; je++) {
    const int l = edge_loc[je];
    if (l >= 0) { const EdgeGeom eg = edge_geom(m, je); L.ex[l] = eg.x; L.ey[l] = eg.y; L.ea[l] = eg.a; L.enx[l] = eg.nx; L.eny[l] = eg.ny; }
  }

  lap("local edges");
  // ---- faces as sliced ELL
  const int nsl = L.nslices = (L.n_own + 31) / 32;
  L.f_off.assign(nsl + 1, 0);
  L.g_off.assign(nsl + 1, 0);
  for (int s = 0; s < nsl; s++) {
    int wf = 0, wg = 0;
    for (int i = 32 * s; i < std::min(L.n_own, 32 * s + 32); i++) {
      const int o = loc2m[i];
      wf = std::max(wf, m.cptr[o + 1] - m.cptr[o]);
      wg = std::max(wg, (int)(g.ptr[o + 1] - g.ptr[o]));
    }
    const int64_t nf = (int64_t)L.f_off[s] + 32 * wf, ng = (int64_t)L.g_off[s] + 32 * wg;
    if (nf > INT32_MAX || ng > INT32_MAX) return "build_layout: per-GPU list exceeds 2^31 entries; use more GPUs";
    L.f_off[s + 1] = (int)nf;
    L.g_off[s + 1] = (int)ng;
  }
  L.f_nbr.assign(L.f_off[nsl], kFacePad);
  L.f_edge.assign(L.f_off[nsl], 0);
  L.nbf = nbf;
  L.bf_type.resize(nbf); L.bf_edge.resize(nbf);
  // boundary-face ids: sequential in cell order
  std::vector<int> bf_start(L.n_own + 1, 0);
  for (int i = 0; i < L.n_own; i++) {
    const int o = loc2m[i];
    int c = 0;
    for (int s = m.cptr[o]; s < m.cptr[o + 1]; s++) c += m.nghbre[s] < 0;
    bf_start[i + 1] = bf_start[i] + c;
  }
#pragma omp parallel for schedule(static)
  for (int i = 0; i < L.n_own; i++) {
    const int o = loc2m[i];
    const int nv = m.cptr[o + 1] - m.cptr[o];
    FaceTmp f[4];
    for (int k = 0; k < nv; k++) {
      const int s = m.cptr[o] + k, je = m.cedge[s];
      f[k].edge = je; f[k].nbr = m.nghbre[s];
      f[k].key_b = m.nghbre[s] >= 0 ? 0 : 1 + m.edge_bc[je];
    }
    std::sort(f, f + nv, [](const FaceTmp &a, const FaceTmp &b) { return a.key_b != b.key_b ? a.key_b < b.key_b : a.edge < b.edge; });
    const int sl = i >> 5, lane = i & 31;
    int bf = bf_start[i];
    for (int k = 0; k < nv; k++) {
      const int e = L.f_off[sl] + 32 * k + lane;
      const int le = edge_loc[f[k].edge];
      L.f_edge[e] = 2 * le + (m.ec1[f[k].edge] == o ? 0 : 1);
      if (f[k].nbr >= 0) {
        L.f_nbr[e] = to_local(f[k].nbr);
      } else {
        L.f_nbr[e] = -1 - bf;
        L.bf_type[bf] = m.b_type[m.edge_bc[f[k].edge]];
        L.bf_edge[bf] = le;
        bf++;
      }
    }
  }

  lap("faces");
  // ---- tile metadata for the shared-memory pass-B kernel
  {
    const int nt = L.ntiles;
    std::vector<std::vector<int>> hc(nt), he(nt);
#pragma omp parallel for schedule(dynamic, 64)
    for (int t = 0; t < nt; t++) {
      const int c0 = t * kTile, c1 = std::min(L.n_own, c0 + kTile);
      std::vector<int> &vc = hc[t], &ve = he[t];
      for (int i = c0; i < c1; i++) {
        const int sl = i >> 5, lane = i & 31;
        const int w = (L.f_off[sl + 1] - L.f_off[sl]) >> 5;
        for (int k = 0; k < w; k++) {
          const int e = L.f_off[sl] + 32 * k + lane;
          const int nb = L.f_nbr[e];
          if (nb == kFacePad) continue;
          if (nb >= 0 && (nb < c0 || nb >= c1)) vc.push_back(nb);
          const int le = L.f_edge[e] >> 1;
          if (le < L.tile_es[t]) ve.push_back(le);
        }
      }
      std::sort(vc.begin(), vc.end()); vc.erase(std::unique(vc.begin(), vc.end()), vc.end());
      std::sort(ve.begin(), ve.end()); ve.erase(std::unique(ve.begin(), ve.end()), ve.end());
    }
    L.tile_hc_ptr.assign(nt + 1, 0);
    L.tile_he_ptr.assign(nt + 1, 0);
    for (int t = 0; t < nt; t++) {
      L.tile_hc_ptr[t + 1] = L.tile_hc_ptr[t] + (int)hc[t].size();
      L.tile_he_ptr[t + 1] = L.tile_he_ptr[t] + (int)he[t].size();
      L.tile_hc_max = std::max(L.tile_hc_max, (int)hc[t].size());
      L.tile_e_max = std::max(L.tile_e_max, L.tile_ne[t] + (int)he[t].size());
    }
    L.tile_hc_idx.resize(L.tile_hc_ptr[nt]);
    L.tile_he_idx.resize(L.tile_he_ptr[nt]);
    L.f_pack.assign(L.f_nbr.size(), 0xFFFEu);
    L.f_bf.assign(L.f_nbr.size(), -1);
#pragma omp parallel for schedule(dynamic, 64)
    for (int t = 0; t < nt; t++) {
      std::copy(hc[t].begin(), hc[t].end(), L.tile_hc_idx.begin() + L.tile_hc_ptr[t]);
      std::copy(he[t].begin(), he[t].end(), L.tile_he_idx.begin() + L.tile_he_ptr[t]);
      const int c0 = t * kTile, c1 = std::min(L.n_own, c0 + kTile);
      for (int i = c0; i < c1; i++) {
        const int sl = i >> 5, lane = i & 31;
        const int w = (L.f_off[sl + 1] - L.f_off[sl]) >> 5;
        for (int k = 0; k < w; k++) {
          const int e = L.f_off[sl] + 32 * k + lane;
          const int nb = L.f_nbr[e];
          if (nb == kFacePad) continue;
          uint32_t ns;
          if (nb < 0) { ns = 0xFFFFu; L.f_bf[e] = -1 - nb; }
          else if (nb >= c0 && nb < c1) ns = (uint32_t)(nb - c0);
          else ns = (uint32_t)(kTile + (std::lower_bound(hc[t].begin(), hc[t].end(), nb) - hc[t].begin()));
          const int le = L.f_edge[e] >> 1;
          uint32_t es;
          if (le >= L.tile_es[t]) es = (uint32_t)(le - L.tile_es[t]);
          else es = (uint32_t)(L.tile_ne[t] + (std::lower_bound(he[t].begin(), he[t].end(), le) - he[t].begin()));
          L.f_pack[e] = ns | (es << 16) | ((uint32_t)(L.f_edge[e] & 1) << 31);
        }
      }
    }
    // tile headers + tile-sliced face table
    L.tile_hdr.assign(8 * (size_t)nt, 0);
    int64_t fbase = 0;
    for (int t = 0; t < nt; t++) {
      int fw = 0;
      for (int sl = (kTile / 32) * t; sl < std::min(L.nslices, (kTile / 32) * (t + 1)); sl++) fw = std::max(fw, (L.f_off[sl + 1] - L.f_off[sl]) >> 5);
      int *h = &L.tile_hdr[8 * (size_t)t];
      h[0] = L.tile_es[t]; h[1] = L.tile_ne[t];
      h[2] = L.tile_hc_ptr[t]; h[3] = L.tile_hc_ptr[t + 1] - L.tile_hc_ptr[t];
      h[4] = L.tile_he_ptr[t]; h[5] = L.tile_he_ptr[t + 1] - L.tile_he_ptr[t];
      h[6] = (int)fbase; h[7] = fw;
      fbase += (int64_t)fw * kTile;
      if (fbase > INT32_MAX) return "build_layout: face table exceeds 2^31 entries; use more GPUs";
    }
    L.t_pack.assign(fbase, 0xFFFEu);
    L.t_bf.assign(fbase, -1);
#pragma omp parallel for schedule(static)
    for (int t = 0; t < nt; t++) {
      const int *h = &L.tile_hdr[8 * (size_t)t];
      const int c0 = t * kTile, c1 = std::min(L.n_own, c0 + kTile);
      for (int i = c0; i < c1; i++) {
        const int sl = i >> 5, lane = i & 31;
        const int w = (L.f_off[sl + 1] - L.f_off[sl]) >> 5;
        for (int k = 0; k < w; k++) {
          const int e = L.f_off[sl] + 32 * k + lane;
          L.t_pack[h[6] + kTile * k + (i - c0)] = L.f_pack[e];
          L.t_bf[h[6] + kTile * k + (i - c0)] = L.f_bf[e];
        }
      }
    }
    if (L.tile_hc_max + kTile >= 0xFFFE || L.tile_e_max >= 0x7FFF) L.tile_hc_max = -1;  // no tile kernel for this mesh
  }

  lap("tiles");
  // ---- gradient stencil as sliced ELL (padding: the cell itself with zero weight)
  L.g_idx.resize(L.g_off[nsl]);
  L.g_cx.assign(L.g_off[nsl], 0.0);
  L.g_cy.assign(L.g_off[nsl], 0.0);
  if (g.form == 0) { L.c0x.resize(L.n_own); L.c0y.resize(L.n_own); }
  double verr = 0;
  int singular = 0;
#pragma omp parallel for schedule(static) reduction(max : verr) reduction(+ : singular)
  for (int s = 0; s < nsl; s++) {
    const int w = (L.g_off[s + 1] - L.g_off[s]) >> 5;
    double cx[kMaxStencil], cy[kMaxStencil];
    for (int lane = 0; lane < 32; lane++) {
      const int i = 32 * s + lane;
      const bool live = i < L.n_own;
      const int o = live ? loc2m[i] : 0;
      const int n = live ? (int)(g.ptr[o + 1] - g.ptr[o]) : 0;
      double c0x = 0, c0y = 0;
      if (live) {
        const double e = grad_cell_coeffs(m, g, o, cx, cy, c0x, c0y);
        if (e < 0) singular++; else verr = std::max(verr, e);
      }
      for (int k = 0; k < w; k++) {
        const int e = L.g_off[s] + 32 * k + lane;
        if (k < n) {
          L.g_idx[e] = to_local(g.idx[g.ptr[o] + k]);
          L.g_cx[e] = cx[k];
          L.g_cy[e] = cy[k];
        } else {
          L.g_idx[e] = live ? i : 0;
        }
      }
      if (live && g.form == 0) { L.c0x[i] = c0x; L.c0y[i] = c0y; }
    }
  }
  L.lsq_verify_err = verr;
  if (singular) return "gradient_lsq: singular least-squares system";
  if (g.form == 1 && !(verr <= 1.0e-10)) return " LSQ coefficients are not correct";

  lap("gradient operator");
  // ---- halo plan
  if (nranks > 1) {
    std::vector<std::vector<unsigned char>> need(nranks);
    for (int p = 0; p < nranks; p++) {
      if (p == rank) continue;
      const int p0 = range_begin(p), p1 = range_begin(p + 1);
      const int q0 = pos_of(p0), q1 = pos_of(p1);   // p's cells that `m` holds (all of them in the global build)
      std::vector<unsigned char> &mask = need[p];
      mask.assign(L.n_own, 0);
      bool any = false;
      auto of_p = [&](int mc) { return new_id[mc] >= p0 && new_id[mc] < p1; };
#pragma omp parallel for schedule(static) reduction(|| : any)
      for (int q = q0; q < q1; q++) {
        const int o = order[q];
        for (int s = m.cptr[o]; s < m.cptr[o + 1]; s++) {
          const int j = m.nghbre[s];
          if (j >= 0 && owned(j)) {
#pragma omp atomic write
            mask[new_id[j] - b0] = 1;
            any = true;
          }
        }
        for (int64_t t = g.ptr[o]; t < g.ptr[o + 1]; t++) {
          const int j = g.idx[t];
          if (owned(j)) {
#pragma omp atomic write
            mask[new_id[j] - b0] = 1;
            any = true;
          }
        }
        if (deep)  // p also stores the stencil members of its face-neighbour ghosts
          for (int s = m.cptr[o]; s < m.cptr[o + 1]; s++) {
            const int j = m.nghbre[s];
            if (j < 0 || of_p(j)) continue;
            for (int64_t t = g.ptr[j]; t < g.ptr[j + 1]; t++) {
              const int kk = g.idx[t];
              if (owned(kk)) {
#pragma omp atomic write
            mask[new_id[kk] - b0] = 1;
            any = true;
          }
            }
          }
      }
      if (!any) mask.clear();
    }
    L.send_ptr.push_back(0);
    auto ghost_pos = [&](int newid) {
      return (int)(std::lower_bound(ghosts.begin(), ghosts.end(), newid, [&](int mc, int v) { return new_id[mc] < v; }) - ghosts.begin());
    };
    for (int p = 0; p < nranks; p++) {
      if (p == rank) continue;
      const int p0 = range_begin(p), p1 = range_begin(p + 1);
      // ghosts owned by p: contiguous run of the sorted ghost list
      const int gb = ghost_pos(p0), ge = ghost_pos(p1);
      const bool sends = !need[p].empty();
      if (ge == gb && !sends) continue;
      L.peers.push_back(p);
      L.recv_begin.push_back(L.n_own + gb);
      L.recv_count.push_back(ge - gb);
      if (sends)
        for (int i = 0; i < L.n_own; i++) if (need[p][i]) L.send_idx.push_back(i);
      L.send_ptr.push_back((int)L.send_idx.size());
    }
    // interior / boundary tiles
    std::vector<unsigned char> bnd(L.ntiles, 0);
    for (int i : L.send_idx) bnd[i / kTile] = 1;
#pragma omp parallel for schedule(static)
    for (int t = 0; t < L.ntiles; t++) {
      bool b = bnd[t];
      for (int i = t * kTile; i < std::min(L.n_own, (t + 1) * kTile) && !b; i++) {
        const int sl = i >> 5, lane = i & 31;
        for (int k = 0; k < ((L.f_off[sl + 1] - L.f_off[sl]) >> 5) && !b; k++) b = L.f_nbr[L.f_off[sl] + 32 * k + lane] >= L.n_own;
        for (int k = 0; k < ((L.g_off[sl + 1] - L.g_off[sl]) >> 5) && !b; k++) b = L.g_idx[L.g_off[sl] + 32 * k + lane] >= L.n_own;
      }
      bnd[t] = b;
    }
    for (int t = 0; t < L.ntiles; t++) (bnd[t] ? L.tile_bnd : L.tile_int).push_back(t);
  }
  return "";
}

std::string build_fused_tables(Layout &L) {
  if (L.fz_built) return "";
  L.fz_built = -1;
  if (L.nranks != 1 && !L.deep) return "";  // several ranks need the deep ghost layers (build_layout(..., deep))
  if (L.tile_hc_max < 0) return "";  // no tile kernel for this mesh
  const int nt = L.ntiles, nsl = L.nslices, n_own = L.n_own;
  // stencil entries of a local cell: owned cells -- the live prefix of the sliced-ELL column (padding entries carry a
  // zero coefficient and the cell's own id; an interior zero coefficient is kept, it costs nothing); face-neighbour
  // ghosts -- their CSR list
  auto width = [&](int i) { return i < n_own ? (L.g_off[(i >> 5) + 1] - L.g_off[i >> 5]) >> 5 : L.gh_ptr[i - n_own + 1] - L.gh_ptr[i - n_own]; };
  auto member = [&](int i, int k) { return i < n_own ? L.g_idx[L.g_off[i >> 5] + 32 * k + (i & 31)] : L.gh_idx[L.gh_ptr[i - n_own] + k]; };
  int wmax = 0;
  for (int s = 0; s < nsl; s++) wmax = std::max(wmax, (L.g_off[s + 1] - L.g_off[s]) >> 5);
  for (size_t k = 0; k + 1 < L.gh_ptr.size(); k++) wmax = std::max(wmax, L.gh_ptr[k + 1] - L.gh_ptr[k]);
  L.fz_w = wmax;
  std::vector<std::vector<int>> h2(nt);
  std::vector<int> gw(nt, 0);
#pragma omp parallel for schedule(dynamic, 64)
  for (int t = 0; t < nt; t++) {
    const int c0 = t * kTile, c1 = std::min(L.n_own, c0 + kTile);
    const int *h1 = L.tile_hc_idx.data() + L.tile_hc_ptr[t];
    const int n1 = L.tile_hc_ptr[t + 1] - L.tile_hc_ptr[t];
    std::vector<int> &v = h2[t];
    int w = 0;
    auto visit = [&](int i) {
      const int wi = width(i);
      w = std::max(w, wi);
      for (int k = 0; k < wi; k++) {
        const int j = member(i, k);
        if (j >= c0 && j < c1) continue;
        if (std::binary_search(h1, h1 + n1, j)) continue;
        v.push_back(j);
      }
    };
    for (int i = c0; i < c1; i++) visit(i);
    for (int h = 0; h < n1; h++) visit(h1[h]);
    std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end());
    gw[t] = w;
  }
  L.fz_hdr.assign(8 * (size_t)nt, 0);
  int64_t hp = 0, gs = 0;
  for (int t = 0; t < nt; t++) {
    const int n1 = L.tile_hc_ptr[t + 1] - L.tile_hc_ptr[t], n2 = (int)h2[t].size();
    const int tw = (kTile + n1 + 7) & ~7;
    int *h = &L.fz_hdr[8 * (size_t)t];
    h[0] = (int)hp; h[1] = n2; h[2] = (int)gs; h[3] = gw[t];
    hp += n2; gs += (int64_t)gw[t] * tw;
    L.fz_h2_max = std::max(L.fz_h2_max, n2);
    L.fz_s2_max = std::max(L.fz_s2_max, kTile + n1 + n2);
    L.fz_tw_max = std::max(L.fz_tw_max, tw);
    if (hp > INT32_MAX || gs > INT32_MAX) return "";  // tables too large for 32-bit offsets: stay on the two-pass path
  }
  if (L.fz_s2_max >= 0xFFFF) return "";
  L.fz_h2_idx.resize(hp);
  L.fz_gslot.assign(gs, 0);
#pragma omp parallel for schedule(dynamic, 64)
  for (int t = 0; t < nt; t++) {
    const int c0 = t * kTile, c1 = std::min(L.n_own, c0 + kTile);
    const int *h1 = L.tile_hc_idx.data() + L.tile_hc_ptr[t];
    const int n1 = L.tile_hc_ptr[t + 1] - L.tile_hc_ptr[t];
    const int tw = (kTile + n1 + 7) & ~7;
    const std::vector<int> &v = h2[t];
    const int *h = &L.fz_hdr[8 * (size_t)t];
    std::copy(v.begin(), v.end(), L.fz_h2_idx.begin() + h[0]);
    uint16_t *tab = L.fz_gslot.data() + h[2];
    for (int k = 0; k < h[3]; k++)
      for (int c = 0; c < tw; c++) tab[(size_t)k * tw + c] = (uint16_t)std::min(c, kTile + n1 - 1);  // default: itself
    auto slot_of = [&](int j) -> int {
      if (j >= c0 && j < c1) return j - c0;
      const int *p1 = std::lower_bound(h1, h1 + n1, j);
      if (p1 != h1 + n1 && *p1 == j) return kTile + (int)(p1 - h1);
      return kTile + n1 + (int)(std::lower_bound(v.begin(), v.end(), j) - v.begin());
    };
    auto fill = [&](int i, int c) {
      const int wi = width(i);
      for (int k = 0; k < wi; k++) tab[(size_t)k * tw + c] = (uint16_t)slot_of(member(i, k));
    };
    for (int i = c0; i < c1; i++) fill(i, i - c0);
    for (int hh = 0; hh < n1; hh++) fill(h1[hh], kTile + hh);
  }
  L.fz_built = 1;
  if (L.nranks > 1) {  // overlap of the state exchange: tiles that neither read a ghost (rings 1, 2) nor hold a sent cell
    std::vector<unsigned char> bnd(nt, 0);
    for (int i : L.send_idx) bnd[i / kTile] = 1;
    for (int t = 0; t < nt; t++) {
      const int *h1 = L.tile_hc_idx.data() + L.tile_hc_ptr[t];
      const int n1 = L.tile_hc_ptr[t + 1] - L.tile_hc_ptr[t];
      if (n1 > 0 && h1[n1 - 1] >= n_own) bnd[t] = 1;               // lists are ascending: the last entry decides
      if (!h2[t].empty() && h2[t].back() >= n_own) bnd[t] = 1;
      (bnd[t] ? L.fz_tile_bnd : L.fz_tile_int).push_back(t);
    }
  }

  // ---- second variant: face table with reverse face indices + the list of tile/ring-1 faces
  if (L.tile_e_max >= 0x1000) return "";
  std::vector<std::vector<uint32_t>> hf(nt);
  L.fz_pack2.assign(L.t_pack.size(), 0xFFFEu);
#pragma omp parallel for schedule(dynamic, 64)
  for (int t = 0; t < nt; t++) {
    const int c0 = t * kTile, c1 = std::min(L.n_own, c0 + kTile);
    const int *th = &L.tile_hdr[8 * (size_t)t];
    const int fbase = th[6];
    for (int k = 0; k < th[7]; k++)
      for (int j = 0; j < c1 - c0; j++) {
        const uint32_t pk = L.t_pack[fbase + kTile * k + j];
        uint32_t code = pk & 0xFFFFu, rev = 0;
        const uint32_t es = (pk >> 16) & 0x7FFFu;
        if (code < (uint32_t)kTile) {
          // the same local edge in the neighbour's face list
          const int i = c0 + j, nb = c0 + (int)code;
          const int le = L.f_edge[L.f_off[i >> 5] + 32 * k + (i & 31)] >> 1;
          const int wn = (L.f_off[(nb >> 5) + 1] - L.f_off[nb >> 5]) >> 5;
          for (int kk = 0; kk < wn; kk++) {
            const int e = L.f_off[nb >> 5] + 32 * kk + (nb & 31);
            if (L.f_nbr[e] != kFacePad && (L.f_edge[e] >> 1) == le) rev = (uint32_t)kk;
          }
        } else if (code < 0xFFFEu) {
          hf[t].push_back((code - kTile) | (es << 16));
          code = (uint32_t)(kTile + hf[t].size() - 1);
        }
        L.fz_pack2[fbase + kTile * k + j] = code | (es << 16) | (rev << 28) | (pk & 0x80000000u);
      }
  }
  int64_t fp = 0;
  for (int t = 0; t < nt; t++) {
    int *h = &L.fz_hdr[8 * (size_t)t];
    h[4] = (int)fp; h[5] = (int)hf[t].size();
    fp += ((int64_t)hf[t].size() + 3) & ~3;  // 16-byte granules (bulk copy)
    L.fz_hf_max = std::max(L.fz_hf_max, (int)hf[t].size());
    if (fp > INT32_MAX) return "";
  }
  L.fz_hf.assign(fp, 0);
  for (int t = 0; t < nt; t++) std::copy(hf[t].begin(), hf[t].end(), L.fz_hf.begin() + L.fz_hdr[8 * (size_t)t + 4]);
  L.fz_v2 = 1;
  return "";
}

void fused_coeff_rows(const Layout &L, size_t np, std::vector<double> &rows) {
  const int F0 = L.g_form == 0 ? 1 : 0;
  rows.assign((size_t)(L.fz_w + F0) * np * 2, 0.0);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < L.n_own; i++) {
    const int sl = i >> 5, lane = i & 31, w = (L.g_off[sl + 1] - L.g_off[sl]) >> 5;
    if (F0) { rows[2 * (size_t)i] = L.c0x[i]; rows[2 * (size_t)i + 1] = L.c0y[i]; }
    for (int k = 0; k < w; k++) {
      const int e = L.g_off[sl] + 32 * k + lane;
      const size_t o = 2 * ((size_t)(k + F0) * np + i);
      rows[o] = L.g_cx[e]; rows[o + 1] = L.g_cy[e];
    }
  }
  for (size_t k = 0; k + 1 < L.gh_ptr.size(); k++) {  // face-neighbour ghosts (several ranks, deep ghost layers)
    const size_t i = (size_t)L.n_own + k;
    if (L.gh_ptr[k + 1] == L.gh_ptr[k]) continue;
    if (F0) { rows[2 * i] = L.gh_c0x[k]; rows[2 * i + 1] = L.gh_c0y[k]; }
    for (int e = L.gh_ptr[k]; e < L.gh_ptr[k + 1]; e++) {
      const size_t o = 2 * ((size_t)(e - L.gh_ptr[k] + F0) * np + i);
      rows[o] = L.gh_cx[e]; rows[o + 1] = L.gh_cy[e];
    }
  }
}

}  // namespace fvs2d
