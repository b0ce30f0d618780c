// host_mesh.cpp -- host-side (CPU, runs once) mesh pre-processing for the GPU hot path.
//
// What: the products of the reference's grid_data (src/grid_procs.f90:170-794) and of the three
// gradient set-ups (src/gradient_ggcb.f90:48-110, src/gradient_ggnb.f90:49-177,
// src/gradient_lsq.f90:70-365) with the reference's numbering (edge ids, c1<c2, normal c1->c2,
// boundary-edge list order), because the device accumulates face fluxes in the reference's order.
// How: not the reference's algorithms -- counting sorts, prefix sums and independent per-cell loops
// (OpenMP), so a 69 M-cell mesh is processed in seconds rather than minutes.
#include "host_mesh.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <numeric>

#include <omp.h>

namespace fvs2d {

static inline double tri_area(double x1, double x2, double x3, double y1, double y2, double y3) {
  return 0.5 * (x1 * (y2 - y3) + x2 * (y3 - y1) + x3 * (y1 - y2));  // src/grid_procs.f90:800-808
}

std::string build_mesh(HostMesh &m) {
  const int nc = m.ncells = m.ntri + m.nquad;
  const int nn = m.nnodes;
  if ((int)m.cptr.size() != nc + 1) return "build_mesh: cell_ptr has the wrong length";
  const int nslots = m.cptr[nc];
  for (int ic = 0; ic < nc; ic++) {
    int nv = m.cptr[ic + 1] - m.cptr[ic];
    if (nv != (ic < m.ntri ? 3 : 4)) return "build_mesh: cells must be listed triangles first, then quads";
  }
  for (int s = 0; s < nslots; s++)
    if (m.cnode[s] < 0 || m.cnode[s] >= nn) return "build_mesh: node id out of range";

  // -- centroids and volumes (src/grid_procs.f90:185-227)
  m.xc.resize(nc); m.yc.resize(nc); m.vol.resize(nc);
#pragma omp parallel for schedule(static)
  for (int ic = 0; ic < nc; ic++) {
    const int *nd = &m.cnode[m.cptr[ic]];
    const int nv = m.nvrt(ic);
    double xc = 0, yc = 0;
    for (int iv = 0; iv < nv; iv++) { xc = xc + m.xn[nd[iv]]; yc = yc + m.yn[nd[iv]]; }
    m.xc[ic] = xc / (double)nv;
    m.yc[ic] = yc / (double)nv;
    double x1 = m.xn[nd[0]], y1 = m.yn[nd[0]], x2 = m.xn[nd[1]], y2 = m.yn[nd[1]], x3 = m.xn[nd[2]], y3 = m.yn[nd[2]];
    double v = tri_area(x1, x2, x3, y1, y2, y3);
    if (nv == 4) { double x4 = m.xn[nd[3]], y4 = m.yn[nd[3]]; v = v + tri_area(x1, x3, x4, y1, y3, y4); }
    m.vol[ic] = v;
  }
  {  // effective lengths, sequential sums as in src/grid_procs.f90:233-243
    double at = 0, as = 0;
    for (int ic = 0; ic < nc; ic++) { at += m.vol[ic]; as += std::sqrt(m.vol[ic]); }
    m.vol_sum = at;
    m.heff1 = std::sqrt(at / (double)nc);
    m.heff2 = as / (double)nc;
  }

  // -- node -> cell by counting sort; ascending cell id per node like src/grid_procs.f90:291-299
  m.n2c_ptr.assign(nn + 1, 0);
  for (int s = 0; s < nslots; s++) m.n2c_ptr[m.cnode[s] + 1]++;
  for (int i = 0; i < nn; i++) m.n2c_ptr[i + 1] += m.n2c_ptr[i];
  m.n2c.resize(nslots);
  {
    std::vector<int> fill(m.n2c_ptr.begin(), m.n2c_ptr.end() - 1);
    for (int ic = 0; ic < nc; ic++)
      for (int s = m.cptr[ic]; s < m.cptr[ic + 1]; s++) m.n2c[fill[m.cnode[s]]++] = ic;
  }

  // -- face neighbours: the cell around v_k that holds the reversed edge (v_k+1 -> v_k)
  m.nghbre.assign(nslots, -1);
#pragma omp parallel for schedule(static)
  for (int ic = 0; ic < nc; ic++) {
    const int nv = m.nvrt(ic);
    const int *nd = &m.cnode[m.cptr[ic]];
    for (int e = 0; e < nv; e++) {
      const int vR = nd[e], vL = nd[(e + 1) % nv];
      int found = -1;
      for (int j = m.n2c_ptr[vR]; j < m.n2c_ptr[vR + 1] && found < 0; j++) {
        const int jc = m.n2c[j];
        const int nvj = m.nvrt(jc);
        const int *ndj = &m.cnode[m.cptr[jc]];
        for (int ii = 0; ii < nvj; ii++)
          if (ndj[ii] == vR && ndj[(ii + nvj - 1) % nvj] == vL) { found = jc; break; }
      }
      m.nghbre[m.cptr[ic] + e] = found;
    }
  }

  // -- global edges: cell order, local-edge order, owner = lower cell id (src/grid_procs.f90:403-624)
  std::vector<int> estart(nc + 1, 0);
#pragma omp parallel for schedule(static)
  for (int ic = 0; ic < nc; ic++) {
    int cnt = 0;
    for (int s = m.cptr[ic]; s < m.cptr[ic + 1]; s++)
      if (m.nghbre[s] > ic || m.nghbre[s] < 0) cnt++;
    estart[ic + 1] = cnt;
  }
  for (int ic = 0; ic < nc; ic++) estart[ic + 1] += estart[ic];
  const int ne = m.nedges = estart[nc];
  m.en1.resize(ne); m.en2.resize(ne); m.ec1.resize(ne); m.ec2.resize(ne);
  m.cedge.assign(nslots, -1);
  int bad = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad)
  for (int ic = 0; ic < nc; ic++) {
    const int nv = m.nvrt(ic), b = m.cptr[ic];
    const int *nd = &m.cnode[b];
    int own = 0;
    for (int e = 0; e < nv; e++) {
      const int jc = m.nghbre[b + e];
      if (jc > ic || jc < 0) {
        const int id = estart[ic] + own++;
        m.en1[id] = nd[e]; m.en2[id] = nd[(e + 1) % nv];
        m.ec1[id] = ic; m.ec2[id] = jc;
        m.cedge[b + e] = id;
      } else {
        // the reference scans jc's nghbr slots k=1..nvrt for ic and maps slot -> local edge (k-2)
        const int nvj = m.nvrt(jc), bj = m.cptr[jc];
        int ej = -1;
        for (int k = 0; k < nvj && ej < 0; k++) {
          const int e2 = (k + nvj - 2) % nvj;
          if (m.nghbre[bj + e2] == ic) ej = e2;
        }
        if (ej < 0) { bad++; continue; }
        int rank = 0;
        for (int e2 = 0; e2 < ej; e2++)
          if (m.nghbre[bj + e2] > jc || m.nghbre[bj + e2] < 0) rank++;
        m.cedge[b + e] = estart[jc] + rank;
      }
    }
  }
  if (bad) return "build_mesh: inconsistent face neighbours (non-conforming mesh?)";

  // -- interior / boundary cells and edges (src/grid_procs.f90:697-763)
  m.cell_intr.clear();
  m.cell_intr.reserve(nc);
  for (int ic = 0; ic < nc; ic++) {
    bool intr = true;
    for (int s = m.cptr[ic]; s < m.cptr[ic + 1]; s++) intr = intr && m.nghbre[s] >= 0;
    if (intr) m.cell_intr.push_back(ic);
  }
  m.ncells_intr = (int)m.cell_intr.size();
  m.ncells_bndr = nc - m.ncells_intr;
  m.b_cell_ptr.assign(m.nb + 1, 0);
  for (int ib = 0; ib < m.nb; ib++) m.b_cell_ptr[ib + 1] = m.b_cell_ptr[ib] + m.b_ncells[ib];
  if (!m.partial && m.ncells_bndr != m.b_cell_ptr[m.nb]) {
    char buf[160];
    snprintf(buf, sizeof buf, "#s of boundary cells does not match (determined %d, .bc file %d)", m.ncells_bndr, m.b_cell_ptr[m.nb]);
    return buf;
  }
  m.nedges_bndr = 0;
  for (int ie = 0; ie < ne; ie++) m.nedges_bndr += m.ec2[ie] < 0;
  m.nedges_intr = ne - m.nedges_bndr;

  // -- boundary edge lists (src/grid_procs.f90:765-791)
  m.edge_bc.assign(ne, -1);
  m.b_edge_ptr.assign(m.nb + 1, 0);
  m.b_edge.clear();
  m.b_edge_src.clear();
  for (int ib = 0; ib < m.nb; ib++) {
    for (int i = m.b_cell_ptr[ib]; i < m.b_cell_ptr[ib + 1]; i++) {
      const int ic = m.b_cell[i];
      if (ic < 0 || ic >= nc) return "build_mesh: boundary cell id out of range";
      for (int s = m.cptr[ic]; s < m.cptr[ic + 1]; s++) {
        const int je = m.cedge[s];
        if (m.ec1[je] == ic && m.ec2[je] < 0) {
          m.b_edge.push_back(je);
          m.b_edge_src.push_back(i);
          m.edge_bc[je] = ib;
        }
      }
    }
    m.b_edge_ptr[ib + 1] = (int)m.b_edge.size();
  }
  if (!m.partial && (int)m.b_edge.size() != m.nedges_bndr) {
    char buf[160];
    snprintf(buf, sizeof buf, "#s of boundary edges/faces does not match (determined %d, read %d)", m.nedges_bndr, (int)m.b_edge.size());
    return buf;
  }
  if (!m.partial)
    for (int ie = 0; ie < ne; ie++)
      if (m.ec2[ie] < 0 && m.edge_bc[ie] < 0) return "build_mesh: a boundary edge belongs to no boundary of the .bc file";

  // -- Green-theorem volume (src/grid_procs.f90:831-840), logged only
  double vg = 0;
#pragma omp parallel for schedule(static) reduction(+ : vg)
  for (int ic = 0; ic < nc; ic++) {
    double v = 0;
    for (int s = m.cptr[ic]; s < m.cptr[ic + 1]; s++) {
      const int je = m.cedge[s];
      const double sgn = m.ec1[je] == ic ? 1.0 : -1.0;
      const EdgeGeom eg = edge_geom(m, je);
      v += eg.nx * sgn * eg.x * eg.a;
    }
    vg += v;
  }
  m.vol_green = vg;
  return "";
}

// ------------------------------------------------------------------------------------------------
// exact k-nearest centroids on a uniform bucket grid (replaces kdtree2 for the LSQ-fn boundary cells,
// src/gradient_lsq.f90:85,108; ties broken by lower cell id)
// ------------------------------------------------------------------------------------------------
namespace {
struct BucketGrid {
  double x0, y0, h;
  int nx, ny;
  std::vector<int> ptr, item;
  void build(const std::vector<double> &x, const std::vector<double> &y) {
    const int n = (int)x.size();
    double x1 = -1e300, y1 = -1e300;
    x0 = y0 = 1e300;
    for (int i = 0; i < n; i++) { x0 = std::min(x0, x[i]); x1 = std::max(x1, x[i]); y0 = std::min(y0, y[i]); y1 = std::max(y1, y[i]); }
    // ~2 centroids per bucket; a degenerate bounding box (one row or column of cells) becomes a 1-D grid along the
    // longer extent.  The bucket counts are formed and clamped in double before the conversion to int.
    const double ex = x1 - x0, ey = y1 - y0, emax = std::max(ex, ey);
    const double area = ex * ey;
    h = area > 1e-12 * emax * emax ? std::sqrt(area / std::max(1.0, n / 2.0)) : emax / std::max(1.0, n / 2.0);
    if (!(h > 0.0)) h = 1.0;  // every centroid at one point
    nx = (int)std::max(1.0, std::min(16384.0, std::floor(ex / h) + 1.0));
    ny = (int)std::max(1.0, std::min(16384.0, std::floor(ey / h) + 1.0));
    h = std::max(ex / nx, ey / ny) * (1.0 + 1e-12) + 1e-300;
    ptr.assign((size_t)nx * ny + 1, 0);
    std::vector<int> b(n);
    for (int i = 0; i < n; i++) { b[i] = bucket(x[i], y[i]); ptr[b[i] + 1]++; }
    for (size_t k = 0; k + 1 < ptr.size(); k++) ptr[k + 1] += ptr[k];
    item.resize(n);
    std::vector<int> fill(ptr.begin(), ptr.end() - 1);
    for (int i = 0; i < n; i++) item[fill[b[i]]++] = i;
  }
  int bx(double x) const { return (int)std::max(0.0, std::min((double)(nx - 1), std::floor((x - x0) / h))); }
  int by(double y) const { return (int)std::max(0.0, std::min((double)(ny - 1), std::floor((y - y0) / h))); }
  int bucket(double x, double y) const { return by(y) * nx + bx(x); }
};
struct Near { double d2; int idx; };
inline bool nearer(const Near &a, const Near &b) { return a.d2 < b.d2 || (a.d2 == b.d2 && a.idx < b.idx); }

void knn(const BucketGrid &g, const std::vector<double> &x, const std::vector<double> &y, int q, int k, Near *out) {
  for (int i = 0; i < k; i++) out[i] = {HUGE_VAL, 0x7fffffff};
  const int cx = g.bx(x[q]), cy = g.by(y[q]);
  const int rmax = std::max(g.nx, g.ny);
  for (int r = 0; r <= rmax; r++) {
    // every bucket at Chebyshev ring r is at least (r-1)*h away from the query
    if (r >= 2) { double dmin = (r - 1) * g.h; if (dmin * dmin > out[k - 1].d2) break; }
    for (int j = cy - r; j <= cy + r; j++) {
      if (j < 0 || j >= g.ny) continue;
      const bool edge_row = (j == cy - r || j == cy + r);
      for (int i = cx - r; i <= cx + r; i += (edge_row ? 1 : 2 * r > 0 ? 2 * r : 1)) {
        if (i < 0 || i >= g.nx) continue;
        const int b = j * g.nx + i;
        for (int t = g.ptr[b]; t < g.ptr[b + 1]; t++) {
          const int c = g.item[t];
          const double dx = x[c] - x[q], dy = y[c] - y[q];
          Near cand{dx * dx + dy * dy, c};
          if (nearer(cand, out[k - 1])) {
            int p = k - 1;
            while (p > 0 && nearer(cand, out[p - 1])) { out[p] = out[p - 1]; p--; }
            out[p] = cand;
          }
        }
      }
    }
  }
}
}  // namespace

std::string build_gradient(const HostMesh &m, int grad_method, int lsq_stencil, double lsq_pow, GradOp &g) {
  const int nc = m.ncells;
  g = GradOp();
  g.method = grad_method; g.lsq_pow = lsq_pow;
  g.form = grad_method == 3 ? 1 : 0;
  g.ptr.assign(nc + 1, 0);
  if (grad_method == 1) {
    // ---- Green-Gauss cell-based: one entry per face; a boundary face points back at the cell itself
    for (int ic = 0; ic < nc; ic++) g.ptr[ic + 1] = m.cptr[ic + 1];
    g.idx.resize(g.ptr[nc]);
#pragma omp parallel for schedule(static)
    for (int ic = 0; ic < nc; ic++)
      for (int s = m.cptr[ic]; s < m.cptr[ic + 1]; s++) g.idx[s] = m.nghbre[s] >= 0 ? m.nghbre[s] : ic;
    return "";
  }
  if (grad_method == 2) {
    // ---- Green-Gauss node-based: stencil = sorted unique cells sharing a vertex (the cell itself is
    // dropped: coefnb = 0, src/gradient_ggnb.f90:145; a neighbour reached through two vertices is merged)
    const int nn = m.nnodes;
    g.idw.resize(nn);
#pragma omp parallel for schedule(static)
    for (int in = 0; in < nn; in++) {
      double idt = 0;
      for (int j = m.n2c_ptr[in]; j < m.n2c_ptr[in + 1]; j++) {
        const int ic = m.n2c[j];
        const double dx = m.xc[ic] - m.xn[in], dy = m.yc[ic] - m.yn[in];
        idt = idt + 1.0 / std::sqrt(dx * dx + dy * dy);
      }
      g.idw[in] = 1.0 / idt;
    }
  } else if (grad_method != 3) {
    return "check cell-center gradient scheme in input file";
  }
  if (grad_method == 3 && lsq_stencil == 0) {
    // ---- LSQ fn: face neighbours; each boundary slot (in nghbr slot order) takes the next nearest centroid
    // that is neither the cell nor a face neighbour (src/gradient_lsq.f90:88-131)
    for (int ic = 0; ic < nc; ic++) g.ptr[ic + 1] = m.cptr[ic + 1];
    std::vector<int> &sten = g.idx;
    sten.assign(g.ptr[nc], -1);
    std::vector<int> bcells;
    int too_many = 0;
    for (int ic = 0; ic < nc; ic++) {
      const int nv = m.nvrt(ic);
      int izb = 0;
      for (int k = 0; k < nv; k++) {
        const int jc = m.nghbr(ic, k);
        sten[m.cptr[ic] + k] = jc;
        izb += jc < 0;
      }
      if (izb > 2) too_many++;
      else if (izb > 0) bcells.push_back(ic);
    }
    if (too_many) return "error in gradient_lsq, sub: setup_fn: #s of edges on the boundary>2!";
    if (!bcells.empty()) {
      BucketGrid bg;
      bg.build(m.xc, m.yc);
      int fail = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : fail)
      for (int t = 0; t < (int)bcells.size(); t++) {
        const int ic = bcells[t], nv = m.nvrt(ic), b = m.cptr[ic];
        Near near[8];
        knn(bg, m.xc, m.yc, ic, 8, near);
        int slot = 0;
        for (int i = 0; i < 8; i++) {
          while (slot < nv && sten[b + slot] >= 0) slot++;
          if (slot == nv) break;
          const int jc = near[i].idx;
          if (jc == ic || jc == 0x7fffffff) continue;
          bool isnb = false;
          for (int k = 0; k < nv; k++) isnb = isnb || m.nghbr(ic, k) == jc;
          if (!isnb) sten[b + slot] = jc;
        }
        for (int k = 0; k < nv; k++) fail += sten[b + k] < 0;
      }
      if (fail) return "gradient_lsq setup_fn: could not complete a boundary-cell stencil from the 8 nearest cells";
    }
    return "";
  }
  // ---- vertex-neighbour stencils (LSQ nn: src/gradient_lsq.f90:227-271; GGNB)
  std::vector<int> cnt(nc);
  auto collect = [&](int ic, int *tmp) {
    int nt = 0;
    for (int s = m.cptr[ic]; s < m.cptr[ic + 1]; s++) {
      const int iv = m.cnode[s];
      for (int j = m.n2c_ptr[iv]; j < m.n2c_ptr[iv + 1]; j++)
        if (m.n2c[j] != ic && nt < kMaxStencil) tmp[nt++] = m.n2c[j];
    }
    std::sort(tmp, tmp + nt);
    return (int)(std::unique(tmp, tmp + nt) - tmp);
  };
  int overflow = 0;
#pragma omp parallel for schedule(static) reduction(+ : overflow)
  for (int ic = 0; ic < nc; ic++) {
    int tot = 0;
    for (int s = m.cptr[ic]; s < m.cptr[ic + 1]; s++) tot += m.n2c_ptr[m.cnode[s] + 1] - m.n2c_ptr[m.cnode[s]];
    if (tot > kMaxStencil) { overflow++; cnt[ic] = 0; continue; }
    int tmp[kMaxStencil];
    cnt[ic] = collect(ic, tmp);
  }
  if (overflow) return "build_gradient: node valence too high for the stencil buffer";
  for (int ic = 0; ic < nc; ic++) g.ptr[ic + 1] = g.ptr[ic] + cnt[ic];
  g.idx.resize(g.ptr[nc]);
#pragma omp parallel for schedule(static)
  for (int ic = 0; ic < nc; ic++) { int tmp[kMaxStencil]; const int n = collect(ic, tmp); std::copy(tmp, tmp + n, g.idx.begin() + g.ptr[ic]); }
  return "";
}

// Coefficients of one cell for its stencil g.idx[g.ptr[ic] ...): cx/cy (n entries) and, for the Green-Gauss
// forms, c0x/c0y.  Returns the LSQ linear-exactness error of the cell (0 for Green-Gauss), or -1 when the
// least-squares system is singular.
double grad_cell_coeffs(const HostMesh &m, const GradOp &g, int ic, double *cx, double *cy, double &c0x, double &c0y) {
  const int64_t b = g.ptr[ic];
  const int n = (int)(g.ptr[ic + 1] - b);
  const double xc = m.xc[ic], yc = m.yc[ic], vol = m.vol[ic];
  c0x = c0y = 0;
  if (g.method == 1) {  // src/gradient_ggcb.f90:55-107, 1/vol folded in
    for (int k = 0; k < n; k++) {
      const int s = m.cptr[ic] + k, je = m.cedge[s];
      const double sgn = m.ec1[je] == ic ? 1.0 : -1.0;
      const EdgeGeom eg = edge_geom(m, je);
      const double af = eg.a, nxf = eg.nx * sgn, nyf = eg.ny * sgn;
      double dx = eg.x - xc, dy = eg.y - yc;
      const double d0 = std::sqrt(dx * dx + dy * dy);
      const int jc = g.idx[b + k];
      dx = eg.x - m.xc[jc]; dy = eg.y - m.yc[jc];
      const double d1 = std::sqrt(dx * dx + dy * dy);
      c0x = c0x + d1 / (d0 + d1) * nxf * af;
      c0y = c0y + d1 / (d0 + d1) * nyf * af;
      cx[k] = d0 / (d0 + d1) * nxf * af / vol;
      cy[k] = d0 / (d0 + d1) * nyf * af / vol;
    }
    c0x /= vol; c0y /= vol;
    return 0.0;
  }
  if (g.method == 2) {  // src/gradient_ggnb.f90:85-174: coefedg(v) * coefnb(v, j) summed per neighbour j, 1/vol folded in
    const int nv = m.nvrt(ic), cb = m.cptr[ic];
    double ex[4] = {0, 0, 0, 0}, ey[4] = {0, 0, 0, 0};
    for (int e = 0; e < nv; e++) {
      const int je = m.cedge[cb + e];
      const double sgn = m.ec1[je] == ic ? 1.0 : -1.0;
      const EdgeGeom eg = edge_geom(m, je);
      const double af = eg.a, nxf = eg.nx * sgn, nyf = eg.ny * sgn;
      const int iv1 = m.en1[je], iv2 = m.en2[je];
      double dx = xc - m.xn[iv1], dy = yc - m.yn[iv1];
      const double w1 = 1.0 / std::sqrt(dx * dx + dy * dy);
      dx = xc - m.xn[iv2]; dy = yc - m.yn[iv2];
      const double w2 = 1.0 / std::sqrt(dx * dx + dy * dy);
      c0x = c0x + af * nxf / 2.0 * (w1 * g.idw[iv1] + w2 * g.idw[iv2]);
      c0y = c0y + af * nyf / 2.0 * (w1 * g.idw[iv1] + w2 * g.idw[iv2]);
      for (int v = 0; v < nv; v++) {
        const int iv = m.cnode[cb + v];
        if (iv == iv1) { ex[v] += af * nxf / 2.0; ey[v] += af * nyf / 2.0; }
        if (iv == iv2) { ex[v] += af * nxf / 2.0; ey[v] += af * nyf / 2.0; }
      }
    }
    for (int k = 0; k < n; k++) { cx[k] = 0; cy[k] = 0; }
    for (int v = 0; v < nv; v++) {
      const int iv = m.cnode[cb + v];
      for (int j = m.n2c_ptr[iv]; j < m.n2c_ptr[iv + 1]; j++) {
        const int jc = m.n2c[j];
        if (jc == ic) continue;
        const double dx = m.xc[jc] - m.xn[iv], dy = m.yc[jc] - m.yn[iv];
        const double cnb = g.idw[iv] / std::sqrt(dx * dx + dy * dy);
        const int k = (int)(std::lower_bound(g.idx.begin() + b, g.idx.begin() + b + n, jc) - (g.idx.begin() + b));
        cx[k] += ex[v] * cnb;
        cy[k] += ey[v] * cnb;
      }
    }
    for (int k = 0; k < n; k++) { cx[k] /= vol; cy[k] /= vol; }
    c0x /= vol; c0y /= vol;
    return 0.0;
  }
  if (!g.user_cx.empty()) {  // the caller's own least-squares table; the linear-exactness check is the reference's
    double dfx = 0, dfy = 0;
    for (int i = 0; i < n; i++) {
      cx[i] = g.user_cx[b + i]; cy[i] = g.user_cy[b + i];
      const int jc = g.idx[b + i];
      const double diff = 1.0 * m.yc[jc] + 2.0 * m.xc[jc] - (1.0 * yc + 2.0 * xc);
      dfx += cx[i] * diff; dfy += cy[i] * diff;
    }
    return std::max(std::fabs(dfx - 2.0), std::fabs(dfy - 1.0));
  }
  // least squares: normal equations per cell (src/gradient_lsq.f90:137-203 / 281-347), w folded in
  double g11 = 0, g12 = 0, g21 = 0, g22 = 0;
  double dxw[kMaxStencil], dyw[kMaxStencil], w[kMaxStencil];
  for (int i = 0; i < n; i++) {
    const int jc = g.idx[b + i];
    const double ddx = m.xc[jc] - xc, ddy = m.yc[jc] - yc;
    const double dis = std::sqrt(ddx * ddx + ddy * ddy);
    w[i] = dis > 0.0 ? 1.0 / std::pow(dis, g.lsq_pow) : 0.0;
    dxw[i] = w[i] * ddx; dyw[i] = w[i] * ddy;
  }
  for (int i = 0; i < n; i++) { g11 += dxw[i] * dxw[i]; g12 += dxw[i] * dyw[i]; g21 += dyw[i] * dxw[i]; g22 += dyw[i] * dyw[i]; }
  const double det = g11 * g22 - g12 * g21;
  if (!(std::fabs(det) > 0)) return -1.0;
  const double i11 = 1.0 / det * g22, i22 = 1.0 / det * g11, i12 = -1.0 / det * g12, i21 = -1.0 / det * g21;
  double dfx = 0, dfy = 0;
  for (int i = 0; i < n; i++) {
    const double c1 = i11 * dxw[i] + i12 * dyw[i], c2 = i21 * dxw[i] + i22 * dyw[i];
    cx[i] = c1 * w[i];  // grad = sum coef*w*(p_j - p_i)
    cy[i] = c2 * w[i];
    // linear-exactness self check with f = 2x + y (src/gradient_lsq.f90:490-529)
    const int jc = g.idx[b + i];
    const double diff = 1.0 * m.yc[jc] + 2.0 * m.xc[jc] - (1.0 * yc + 2.0 * xc);
    dfx += c1 * diff * w[i];
    dfy += c2 * diff * w[i];
  }
  return std::max(std::fabs(dfx - 2.0), std::fabs(dfy - 1.0));
}

// ------------------------------------------------------------------------------------------------
// Hilbert ordering
// ------------------------------------------------------------------------------------------------
// LSD radix sort of (key, value) by the low `bits` bits of key, 14 bits per pass, stable (ties keep the input order), with
// per-thread histograms so that the passes run on all the cores the rank has
static void radix_sort_pairs(std::vector<uint64_t> &key, std::vector<int> &val, int bits) {
  const size_t n = key.size();
  std::vector<uint64_t> key2(n);
  std::vector<int> val2(n);
  constexpr int RB = 14, NB = 1 << RB;
  int nthr = 1;
#pragma omp parallel
  {
#pragma omp single
    nthr = omp_get_num_threads();
  }
  std::vector<size_t> hist((size_t)nthr * NB);
  for (int sh = 0; sh < bits; sh += RB) {
    std::fill(hist.begin(), hist.end(), (size_t)0);
#pragma omp parallel num_threads(nthr)
    {
      const int t = omp_get_thread_num();
      const size_t lo = n * (size_t)t / nthr, hi = n * (size_t)(t + 1) / nthr;
      size_t *h = &hist[(size_t)t * NB];
      for (size_t i = lo; i < hi; i++) h[(key[i] >> sh) & (NB - 1)]++;
#pragma omp barrier
#pragma omp single
      {
        size_t run = 0;
        for (int b = 0; b < NB; b++)
          for (int tt = 0; tt < nthr; tt++) { const size_t c = hist[(size_t)tt * NB + b]; hist[(size_t)tt * NB + b] = run; run += c; }
      }
      for (size_t i = lo; i < hi; i++) {
        const size_t pos = h[(key[i] >> sh) & (NB - 1)]++;
        key2[pos] = key[i]; val2[pos] = val[i];
      }
    }
    key.swap(key2); val.swap(val2);
  }
}

// Hilbert order from the raw arrays (no HostMesh needed: the partition-local pre-processing orders the GLOBAL cells
// without building any global connectivity).  xn / yn with stride `xs` (1: separate arrays, 2: interleaved node_xy).
// Step 1 (host, one light pass): the key transformation.  Step 2: keys + stable sort -- on the device when a sorter is
// supplied (fvs2d_gpu_set_mesh: 69 M keys sort in milliseconds there, seconds on two host threads), else on the host;
// both give the same permutation (same key arithmetic, both sorts stable in the original id).
HilbertFrame hilbert_frame(int nc, const int *cptr, const int *cnode, const double *xn, const double *yn, int xs) {
  double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300;
  // Anisotropy: the curve should be compact in CELL COUNTS, not in physical distance (a 128-cell tile of a
  // mesh with 8:1 cells would otherwise be a 2-row strip with a huge halo).  Each axis is measured in units of
  // the mean cell extent along it.
  // (the extent sums are formed per fixed chunk of cells and added up in chunk order: every rank must get the same
  // bits whatever its thread count, because every rank derives the same global order from them)
  constexpr int kChunk = 1 << 16;
  const int nchunk = (nc + kChunk - 1) / kChunk;
  std::vector<double> cex(nchunk, 0.0), cey(nchunk, 0.0);
#pragma omp parallel for schedule(dynamic, 1) reduction(min : x0, y0) reduction(max : x1, y1)
  for (int c = 0; c < nchunk; c++) {
    double sex = 0, sey = 0;
    for (int i = c * kChunk; i < std::min(nc, (c + 1) * kChunk); i++) {
      double xa = 1e300, xb = -1e300, ya = 1e300, yb = -1e300, sx = 0, sy = 0;
      for (int s = cptr[i]; s < cptr[i + 1]; s++) {
        const size_t v = (size_t)cnode[s] * xs;
        xa = std::min(xa, xn[v]); xb = std::max(xb, xn[v]); ya = std::min(ya, yn[v]); yb = std::max(yb, yn[v]);
        sx = sx + xn[v]; sy = sy + yn[v];
      }
      const double nv = (double)(cptr[i + 1] - cptr[i]);
      sx /= nv; sy /= nv;
      x0 = std::min(x0, sx); x1 = std::max(x1, sx); y0 = std::min(y0, sy); y1 = std::max(y1, sy);
      sex += xb - xa; sey += yb - ya;
    }
    cex[c] = sex; cey[c] = sey;
  }
  double ex = 0, ey = 0;
  for (int c = 0; c < nchunk; c++) { ex += cex[c]; ey += cey[c]; }
  ex = std::max(ex / nc, 1e-300); ey = std::max(ey / nc, 1e-300);
  HilbertFrame f;
  f.bits = 20;
  const double span = std::max(std::max((x1 - x0) / ex, (y1 - y0) / ey), 1e-300);
  const double scale = ((double)(1u << f.bits) - 1.0) / span;
  f.x0 = x0; f.y0 = y0; f.scale_x = scale / ex; f.scale_y = scale / ey;
  return f;
}

void hilbert_order_raw(int nc, const int *cptr, const int *cnode, const double *xn, const double *yn, int xs, std::vector<int> &perm,
                       const HilbertSorter *device_sort) {
  const HilbertFrame f = hilbert_frame(nc, cptr, cnode, xn, yn, xs);
  if (device_sort && *device_sort && (*device_sort)(f, nc, cptr, cnode, xn, yn, xs, perm)) return;
  std::vector<uint64_t> key(nc);
  perm.resize(nc);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < nc; i++) {
    double sx = 0, sy = 0;
    for (int s = cptr[i]; s < cptr[i + 1]; s++) { const size_t v = (size_t)cnode[s] * xs; sx = sx + xn[v]; sy = sy + yn[v]; }
    const double nv = (double)(cptr[i + 1] - cptr[i]);
    sx /= nv; sy /= nv;
    const uint32_t ix = (uint32_t)((sx - f.x0) * f.scale_x), iy = (uint32_t)((sy - f.y0) * f.scale_y);
    key[i] = hilbert_d(ix, iy, f.bits);
    perm[i] = i;
  }
  radix_sort_pairs(key, perm, 2 * f.bits);
}

// Cut the Hilbert order into `nranks` contiguous chunks of equal estimated COST, on 128-cell tile boundaries.  A
// quadrilateral costs about 1.4 triangles in the stage kernel (four faces instead of three, and its tiles run at two CTAs
// per SM instead of three; measured 0.108 vs 0.0765 ns per cell-stage on B200), so equal cell counts leave the ranks that
// hold the quadrilateral band of a mixed mesh 25-30 % slower than the all-triangle ones -- and a time step takes as long as
// the slowest rank.  perm[new] = original id; original ids below ntri are triangles.
std::vector<int> partition_cuts(const std::vector<int> &perm, int ntri, int nranks) {
  const int nc = (int)perm.size();
  std::vector<int> cuts(nranks + 1, nc);
  cuts[0] = 0;
  if (nranks == 1) return cuts;
  constexpr long long kTri = 5, kQuad = 7;
  long long total = 0;
#pragma omp parallel for schedule(static) reduction(+ : total)
  for (int i = 0; i < nc; i++) total += perm[i] < ntri ? kTri : kQuad;
  long long run = 0;
  int r = 1;
  for (int i = 0; i < nc && r < nranks; i++) {
    if ((i & 127) == 0 && run * nranks >= total * r) cuts[r++] = i;   // first tile boundary at or past r / nranks of the cost
    run += perm[i] < ntri ? kTri : kQuad;
  }
  for (int k = 1; k <= nranks; k++) cuts[k] = std::max(cuts[k], cuts[k - 1]);
  return cuts;
}

void hilbert_order(const HostMesh &m, std::vector<int> &perm, const HilbertSorter *device_sort) {
  hilbert_order_raw(m.ncells, m.cptr.data(), m.cnode.data(), m.xn.data(), m.yn.data(), 1, perm, device_sort);
}

// ------------------------------------------------------------------------------------------------
// Partition-local pre-processing (several ranks): this rank's cells plus `rings` layers of node-adjacent cells, cut out
// of the caller's global arrays.  Only O(global) work: the Hilbert keys + their sort, and one pass over the global
// cell -> node list per ring against a node bitmap; no global connectivity is ever built.  The submesh keeps the global
// relative order of cells and nodes (ascending original ids), so build_mesh numbers its edges in the reference's relative
// order, and the boundary lists keep the .bc order.
// ------------------------------------------------------------------------------------------------
std::string extract_submesh(int nnodes, int ntri, int nquad, const double *node_xy, const int *cptr, const int *cnode, int nb,
                            const int *b_ncells, const int *b_type, const int *b_cell, int rank, int nranks, int rings, SubMesh &out,
                            const HilbertSorter *device_sort) {
  const int nc = ntri + nquad;
  out = SubMesh();
  out.nc_global = nc; out.nn_global = nnodes;
  for (int ic = 0; ic < nc; ic++) {
    const int nv = cptr[ic + 1] - cptr[ic];
    if (nv != (ic < ntri ? 3 : 4)) return "build_mesh: cells must be listed triangles first, then quads";
  }
  {
    int bad = 0;
    const int nslots = cptr[nc];
#pragma omp parallel for schedule(static) reduction(+ : bad)
    for (int s = 0; s < nslots; s++) bad += cnode[s] < 0 || cnode[s] >= nnodes;
    if (bad) return "build_mesh: node id out of range";
  }
  const double t0 = omp_get_wtime();
  auto lap = [&](const char *what) { if (getenv("FVS2D_DEBUG")) fprintf(stderr, "[fvs2d]   extract_submesh: %-24s %.2f s\n", what, omp_get_wtime() - t0); };
  std::vector<int> perm;
  hilbert_order_raw(nc, cptr, cnode, node_xy, node_xy + 1, 2, perm, device_sort);
  lap("hilbert order");
  out.cuts = partition_cuts(perm, ntri, nranks);
  const int b0 = out.b0 = out.cuts[rank], b1 = out.b1 = out.cuts[rank + 1];
  // ring 0 = owned cells; level[c] = ring + 1, 0 = outside
  std::vector<unsigned char> level(nc, 0), nmark(nnodes, 0);
#pragma omp parallel for schedule(static)
  for (int i = b0; i < b1; i++) level[perm[i]] = 1;
  for (int r = 0; r <= rings; r++) {
    // mark the nodes of ring r, then (r < rings) every unmarked cell that touches a marked node becomes ring r + 1
#pragma omp parallel for schedule(static)
    for (int ic = 0; ic < nc; ic++)
      if (level[ic] == r + 1)
        for (int s = cptr[ic]; s < cptr[ic + 1]; s++) {
#pragma omp atomic write
          nmark[cnode[s]] = 1;
        }
    if (r == rings) break;
#pragma omp parallel for schedule(static)
    for (int ic = 0; ic < nc; ic++) {
      if (level[ic]) continue;
      bool hit = false;
      for (int s = cptr[ic]; s < cptr[ic + 1] && !hit; s++) hit = nmark[cnode[s]] != 0;
      if (hit) level[ic] = (unsigned char)(r + 2);
    }
  }
  lap("rings");
  // Hilbert id of every original cell is needed for the submesh cells only: scatter it through a transient global array
  std::vector<int> new_of(nc);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < nc; i++) new_of[perm[i]] = i;
  perm.clear(); perm.shrink_to_fit();
  // submesh cells / nodes in ascending original id
  std::vector<int> &orig = out.orig;
  for (int ic = 0; ic < nc; ic++) if (level[ic]) orig.push_back(ic);
  const int ns = (int)orig.size();
  std::vector<int> node_loc(nnodes, -1);
  int nn = 0;
  for (int v = 0; v < nnodes; v++) if (nmark[v]) node_loc[v] = nn++;
  HostMesh &m = out.m;
  m.partial = true;
  m.nnodes = nn;
  m.xn.resize(nn); m.yn.resize(nn);
  out.node_orig.resize(nn);
#pragma omp parallel for schedule(static)
  for (int v = 0; v < nnodes; v++)
    if (node_loc[v] >= 0) { m.xn[node_loc[v]] = node_xy[2 * (size_t)v]; m.yn[node_loc[v]] = node_xy[2 * (size_t)v + 1]; out.node_orig[node_loc[v]] = v; }
  m.ntri = (int)(std::lower_bound(orig.begin(), orig.end(), ntri) - orig.begin());
  m.nquad = ns - m.ntri;
  m.ncells = ns;
  m.cptr.resize(ns + 1);
  m.cptr[0] = 0;
  for (int k = 0; k < ns; k++) m.cptr[k + 1] = m.cptr[k] + (cptr[orig[k] + 1] - cptr[orig[k]]);
  m.cnode.resize(m.cptr[ns]);
  out.new_id.resize(ns);
  out.ring.resize(ns);
#pragma omp parallel for schedule(static)
  for (int k = 0; k < ns; k++) {
    const int o = orig[k];
    for (int s = 0; s < cptr[o + 1] - cptr[o]; s++) m.cnode[m.cptr[k] + s] = node_loc[cnode[cptr[o] + s]];
    out.new_id[k] = new_of[o];
    out.ring[k] = (unsigned char)(level[o] - 1);
  }
  lap("submesh arrays");
  // boundary lists restricted to the submesh, .bc order kept (level doubles as the "in submesh" test; ids via binary search)
  m.nb = nb;
  m.b_type.assign(b_type, b_type + nb);
  m.b_ncells.assign(nb, 0);
  size_t off = 0;
  long long nbc_global = 0;
  for (int ib = 0; ib < nb; ib++) {
    for (int i = 0; i < b_ncells[ib]; i++) {
      const int c = b_cell[off + i];
      if (c < 0 || c >= nc) return "build_mesh: boundary cell id out of range";
      if (level[c]) {
        m.b_cell.push_back((int)(std::lower_bound(orig.begin(), orig.end(), c) - orig.begin()));
        out.b_pos.push_back(i);
        m.b_ncells[ib]++;
      }
    }
    off += b_ncells[ib];
    nbc_global += b_ncells[ib];
  }
  out.nbcells_global = nbc_global;
  {  // centroid of original cell 0 (maxloc of an all-zero error field points at it, src/mms.f90:363)
    double sx = 0, sy = 0;
    for (int s = cptr[0]; s < cptr[1]; s++) { sx = sx + node_xy[2 * (size_t)cnode[s]]; sy = sy + node_xy[2 * (size_t)cnode[s] + 1]; }
    out.xy_cell0[0] = sx / (double)(cptr[1] - cptr[0]); out.xy_cell0[1] = sy / (double)(cptr[1] - cptr[0]);
  }
  return "";
}

}  // namespace fvs2d
