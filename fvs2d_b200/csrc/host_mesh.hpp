// host_mesh.hpp -- host-side mesh products of the reference's grid_data (src/grid_procs.f90:170-794),
// rebuilt with the reference's entity numbering by O(n), OpenMP-parallel algorithms.
// All ids are 0-based; "boundary / none" is -1 where the Fortran stores 0.
#pragma once
#include <cmath>
#include <cstdint>
#include <functional>
#include <string>
#include <vector>

namespace fvs2d {

struct HostMesh {
  // raw input (what grid_read / grid_bc_read deliver)
  int nnodes = 0, ntri = 0, nquad = 0, ncells = 0;
  std::vector<double> xn, yn;
  std::vector<int> cptr, cnode;  // CSR cell -> node, triangles first
  int nb = 0;
  std::vector<int> b_ncells, b_type, b_cell_ptr, b_cell;
  // derived
  std::vector<double> xc, yc, vol;
  std::vector<int> n2c_ptr, n2c;      // node -> cells, ascending cell id (src/grid_procs.f90:275-300)
  std::vector<int> nghbre;            // per cell-slot: neighbour across local edge k (v_k -> v_k+1), -1 boundary
  std::vector<int> cedge;             // per cell-slot: global edge id of local edge k
  int nedges = 0, nedges_intr = 0, nedges_bndr = 0;
  std::vector<int> en1, en2, ec1, ec2;  // c1 < c2, c2 == -1 on the boundary, normal points c1 -> c2
  // edge geometry is not stored (4.6 GB at 69 M cells): edge_geom() recomputes it from the end nodes
  int ncells_intr = 0, ncells_bndr = 0;
  std::vector<int> cell_intr;
  std::vector<int> b_edge_ptr, b_edge;  // boundary edge lists in .bc order then local-edge order
  std::vector<int> b_edge_src;          // per b_edge entry: the index into b_cell of the list entry it comes from
  std::vector<int> edge_bc;             // per edge: -1 interior, else boundary index ib
  double heff1 = 0, heff2 = 0, vol_sum = 0, vol_green = 0;
  // partial == true: the mesh is one rank's submesh (extract_submesh); cells of its outermost ring miss neighbours that
  // lie outside ("cut" edges look like boundary edges but belong to no boundary), so the .bc consistency checks of
  // grid_data are not applied and global sums are formed over the owned cells by the caller
  bool partial = false;

  int nvrt(int ic) const { return cptr[ic + 1] - cptr[ic]; }
  // nghbr(slot) of the reference: slot k holds the neighbour across local edge (k-2) (src/grid_procs.f90:330-366)
  int nghbr(int ic, int k) const {
    int nv = nvrt(ic);
    return nghbre[cptr[ic] + (k + nv - 2) % nv];
  }
};

// Builds everything in `m` from the raw fields.  Returns "" or the reference's stop message.
std::string build_mesh(HostMesh &m);

// Edge centre, length and unit normal (c1 -> c2), src/grid_procs.f90:630-647.
struct EdgeGeom { double x, y, a, nx, ny; };
inline EdgeGeom edge_geom(const HostMesh &m, int je) {
  const int v1 = m.en1[je], v2 = m.en2[je];
  const double dx = m.xn[v2] - m.xn[v1], dy = m.yn[v2] - m.yn[v1];
  const double a = std::sqrt(dx * dx + dy * dy);
  return {0.5 * (m.xn[v1] + m.xn[v2]), 0.5 * (m.yn[v1] + m.yn[v2]), a, dy / a, -dx / a};
}

// Gradient operators in one generic sparse form (original numbering):
//   grad_i = c0_i * p_i + sum_k coef_k * p_{idx_k}            (form == 0, Green-Gauss; 1/vol folded in)
//   grad_i =              sum_k coef_k * (p_{idx_k} - p_i)    (form == 1, least squares; w folded in)
// Only the stencil STRUCTURE is kept for all cells (the partitioner needs it); coefficients are produced
// per cell on demand (grad_cell_coeffs), so a rank only ever computes and stores those of its own cells.
constexpr int kMaxStencil = 256;
struct GradOp {
  int form = 0, method = 1;
  double lsq_pow = 0;
  std::vector<int64_t> ptr;  // ncells+1
  std::vector<int> idx;      // the LSQ stencil is also the limiter's min/max set (src/gradient_limiter.f90:54-58)
  std::vector<double> idw;   // GGNB: node weights 1 / sum_c 1/|x_c - x_v| (src/gradient_ggnb.f90:59-80)
  std::vector<double> user_cx, user_cy;  // least squares supplied by the caller (fvs2d_gpu_set_lsq): coef * w per entry
};

// grad_method 1 GGCB (src/gradient_ggcb.f90:48-110), 2 GGNB (src/gradient_ggnb.f90:49-177),
// 3 LSQ fn/nn (src/gradient_lsq.f90:70-365).  Returns "" or an error message.
std::string build_gradient(const HostMesh &m, int grad_method, int lsq_stencil, double lsq_pow, GradOp &g);
double grad_cell_coeffs(const HostMesh &m, const GradOp &g, int ic, double *cx, double *cy, double &c0x, double &c0y);

// Hilbert-curve ordering of the cell centroids (perm[new] = old), measured in cell counts per axis so that
// anisotropic meshes still give compact tiles.  (Tried and rejected on B200: scanline order inside each
// 128-cell tile -- 4-10 % slower pass B than the plain curve.)
std::vector<int> partition_cuts(const std::vector<int> &perm, int ntri, int nranks);
struct HilbertFrame { double x0, y0, scale_x, scale_y; int bits; };  // centroid -> integer lattice of the curve
// Optional device implementation of "keys + stable sort" (api.cu: CUB radix sort); returns false to fall back to the host.
typedef std::function<bool(const HilbertFrame &, int nc, const int *cptr, const int *cnode, const double *xn, const double *yn, int xs,
                           std::vector<int> &perm)> HilbertSorter;
void hilbert_order(const HostMesh &m, std::vector<int> &perm, const HilbertSorter *device_sort = nullptr);
void hilbert_order_raw(int nc, const int *cptr, const int *cnode, const double *xn, const double *yn, int xs, std::vector<int> &perm,
                       const HilbertSorter *device_sort = nullptr);
#ifdef __CUDACC__
__host__ __device__
#endif
inline uint64_t hilbert_d(uint32_t x, uint32_t y, int bits) {
  uint64_t d = 0;
  const uint32_t n1 = (1u << bits) - 1;
  for (uint32_t s = 1u << (bits - 1); s > 0; s >>= 1) {
    const uint32_t rx = (x & s) ? 1 : 0, ry = (y & s) ? 1 : 0;
    d += (uint64_t)s * s * ((3 * rx) ^ ry);
    if (ry == 0) {  // rotate the quadrant; only the bits below s are looked at afterwards
      if (rx == 1) { x = n1 - x; y = n1 - y; }
      const uint32_t t = x; x = y; y = t;
    }
  }
  return d;
}


// One rank's part of a mesh: its cells (a contiguous range [b0, b1) of the global Hilbert order) plus rings of
// node-adjacent cells, as a HostMesh with local ids that keep the original relative order.
struct SubMesh {
  HostMesh m;
  std::vector<int> orig;        // submesh cell -> original cell id (ascending)
  std::vector<int> new_id;      // submesh cell -> Hilbert id in the global order
  std::vector<unsigned char> ring;  // 0 owned, 1 / 2 ... node-adjacency ring around the owned cells
  std::vector<int> node_orig;   // submesh node -> original node id
  std::vector<int> b_pos;       // per entry of m.b_cell: its position inside its boundary's cell list in the caller's .bc order
  int nc_global = 0, nn_global = 0, b0 = 0, b1 = 0;
  std::vector<int> cuts;        // Hilbert ids where the ranks' chunks begin (nranks + 1 entries)
  long long nbcells_global = 0;
  double xy_cell0[2] = {0, 0};
};
std::string extract_submesh(int nnodes, int ntri, int nquad, const double *node_xy, const int *cptr, const int *cnode, int nb,
                            const int *b_ncells, const int *b_type, const int *b_cell, int rank, int nranks, int rings, SubMesh &out,
                            const HilbertSorter *device_sort = nullptr);

}  // namespace fvs2d
