// layout.hpp -- the device-side data layout, built on the host from HostMesh + GradOp.
//
// Cells are renumbered along a Hilbert curve ("new" ids).  A rank owns a contiguous range of new ids
// (the whole mesh on one GPU) and additionally stores ghost copies of every cell its owned cells read
// (face neighbours and gradient-stencil members).  Local ids: owned cells first in new-id order, then
// ghosts sorted by new id -- so the ghosts received from one peer are one contiguous run.
//
// Per-cell lists (faces, gradient stencil) are stored as sliced ELL with slices of 32 cells (one warp):
// entry (k, lane) of slice s sits at off[s] + 32*k + lane, so every list access of a warp is coalesced.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "host_mesh.hpp"

#ifndef FVS2D_TILE
#define FVS2D_TILE 128
#endif

namespace fvs2d {

constexpr int kFacePad = INT32_MIN;  // f_nbr value of a padding entry (triangle in a width-4 slice)
constexpr int kTile = FVS2D_TILE;           // cells per tile (== threads per CTA of the cell-parallel kernels)

struct Layout {
  int rank = 0, nranks = 1;
  int nc_global = 0;
  int n_own = 0, n_loc = 0;        // owned cells, owned + ghost cells
  int own_begin = 0;               // first owned new id
  std::vector<int> perm;           // new id -> original id (whole-mesh builds only; empty after a partition-local build)
  std::vector<int> loc2new;        // local id -> new id (owned: own_begin + i)
  std::vector<int> orig_id;        // local id -> original id
  // per local cell
  std::vector<double> xc, yc;
  // per owned cell
  std::vector<double> vol;
  std::vector<unsigned char> is_intr;  // cell_intr membership (no boundary face), src/grid_procs.f90:704-720
  // faces, sliced ELL over owned cells
  int nslices = 0;
  std::vector<int> f_off;    // nslices+1 entry offsets
  std::vector<int> f_nbr;    // local id of the neighbour | -1-bf (bf = local boundary-face id) | kFacePad
  std::vector<int> f_edge;   // 2*local_edge + (0 if this cell is the edge's c1, else 1)
  // local edges (those touched by owned cells), numbered by first touch
  int nedges = 0;
  std::vector<double> ex, ey, ea, enx, eny;
  // local boundary faces
  int nbf = 0;
  std::vector<int> bf_type, bf_edge;
  // gradient stencil, sliced ELL over owned cells
  int g_form = 0;
  std::vector<int> g_off, g_idx;
  std::vector<double> g_cx, g_cy, c0x, c0y;
  // ---- tiles: kTile consecutive owned cells = one CTA of the shared-memory pass-B kernel.
  // A tile stages in smem: its own cells [t*kTile, ...), its halo cells (face neighbours outside the
  // tile, tile_hc_idx), its own edges (first touched by this tile: the contiguous, even-aligned range
  // [tile_es, tile_es+tile_ne)) and its halo edges (first touched by an earlier tile, tile_he_idx).
  // f_pack = per face entry: bits 0-15 neighbour slot (own: id - tile start; halo: kTile + pos;
  // 0xFFFF boundary; 0xFFFE padding), bits 16-30 edge slot (own: id - tile_es; halo: tile_ne + pos),
  // bit 31 set when this cell is the edge's c2.  f_bf = boundary-face id of a boundary entry.
  int ntiles = 0, tile_hc_max = 0, tile_e_max = 0;
  std::vector<int> tile_es, tile_ne, tile_hc_ptr, tile_he_ptr, tile_hc_idx, tile_he_idx;
  std::vector<uint32_t> f_pack;
  std::vector<int> f_bf;
  // the same per tile in the form the persistent pipeline kernel consumes: one 32-byte header per tile
  // {es, ne, hc_ptr, n_hc, he_ptr, n_he, fbase, fw} and the face table sliced per TILE: entry (k, j) of
  // tile t at fbase + kTile*k + j, k < fw (so a tile's face table is one contiguous, 512-byte aligned run)
  std::vector<int> tile_hdr;       // 8 ints per tile
  std::vector<uint32_t> t_pack;
  std::vector<int> t_bf;
  // halo plan: peers in ascending rank; send_idx = owned local ids the peer needs (ascending new id),
  // recv = contiguous run [recv_begin, recv_begin+recv_count) of local ghost ids
  std::vector<int> peers, send_ptr, send_idx, recv_begin, recv_count;
  // overlap of the halo exchange with interior work: a tile is "boundary" when one of its cells is sent to a
  // peer or reads a ghost (through a face or the gradient stencil); all other tiles are "interior"
  std::vector<int> tile_int, tile_bnd;
  double lsq_verify_err = 0;  // max linear-exactness error over the owned cells (src/gradient_lsq.f90:490-529)
  // ---- one-kernel-per-stage variant (k_stage_fused, single rank): the gradients of a tile's cells AND of its halo
  // cells ("ring 1", tile_hc_idx) are rebuilt in shared memory, so a tile also stages "ring 2" = the gradient-stencil
  // members of tile + ring 1 that are in neither (primitive state only) and the gradient operator of tile + ring 1.
  // Slots of a tile: own cells [0, kTile), ring 1 [kTile, kTile+n1), ring 2 [kTile+n1, kTile+n1+n2).
  // fz_gslot = per tile gw rows of pitch TW = roundup8(kTile+n1): entry (k, c) = slot of the k-th stencil member of
  // the cell in slot c (same k order as g_idx; cells with fewer entries point at themselves, their coefficient is 0).
  // Several ranks: the fused kernel rebuilds the gradients of ring-1 cells that are GHOSTS, so with `deep` ghost layers
  // (build_layout(..., deep = true)) a rank also stores the gradient-stencil members of its face-neighbour ghosts (one
  // state exchange per stage instead of state + gradients: SURVEY 8e variant 2a) and their gradient operator:
  // gh_ptr/gh_idx/gh_cx/gh_cy = CSR over the ghosts (only face-neighbour ghosts have entries; local ids), gh_c0x/y.
  int deep = 0;
  std::vector<int> gh_ptr, gh_idx;
  std::vector<double> gh_cx, gh_cy, gh_c0x, gh_c0y;
  std::vector<int> fz_tile_int, fz_tile_bnd;  // tiles whose rings hold no ghost and none of whose cells is sent / the rest
  int fz_built = 0;            // 0 not built, 1 usable, -1 built but unusable for this mesh
  int fz_w = 0;                // stencil entries per cell, maximum over the mesh (coefficient rows)
  int fz_s2_max = 0, fz_tw_max = 0, fz_h2_max = 0;
  std::vector<int> fz_hdr;     // 8 ints per tile {h2_ptr, n_h2, gs_base, gw, hf_ptr, n_hf, 0, 0}
  std::vector<int> fz_h2_idx;  // ring-2 cells (local ids, ascending) of all tiles
  std::vector<uint16_t> fz_gslot;
  // every reconstructed face state is evaluated once, by the thread that holds the cell's gradient, and published in
  // shared memory.  fz_pack2 = t_pack with the neighbour code of bits 0-15
  // redefined -- < kTile: neighbour's slot in the tile, and bits 28-29 = index of this face in the NEIGHBOUR's face
  // list; >= kTile: kTile + index into the tile's list of tile/ring-1 faces (fz_hf: ring-1 index | edge slot << 16);
  // 0xFFFF boundary, 0xFFFE padding -- and the edge slot in bits 16-27.
  int fz_v2 = 0;               // 1 when the tables below exist (edge slots fit 12 bits)
  int fz_hf_max = 0;
  std::vector<uint32_t> fz_pack2, fz_hf;
};

// How the cells of a HostMesh relate to the global mesh: the whole mesh (orig == null: m-cell id = original id, order = the
// Hilbert permutation) or one rank's submesh (extract_submesh).
struct CellNumbering {
  int nc_global = 0;
  std::vector<int> order;      // m-cells by ascending Hilbert id
  std::vector<int> new_id;     // m-cell -> Hilbert id
  const int *orig = nullptr;   // m-cell -> original id (null: identity)
  std::vector<int> cuts;       // Hilbert ids where the ranks' chunks begin (partition_cuts; nranks + 1 entries)
};
// Builds the layout of `rank` out of `nranks` (contiguous chunks of the Hilbert order of equal estimated cost).
std::string build_layout(const HostMesh &m, const GradOp &g, const CellNumbering &num, int rank, int nranks, Layout &L,
                         bool deep = false);

// Adds the fz_* tables of the fused stage kernel to a single-rank layout (on demand: they are only needed when the
// "fuse" option is on).  Returns "" (then L.fz_built is 1 or -1) or an error message.
std::string build_fused_tables(Layout &L);
// Gradient coefficients in the fused kernel's form: rows of `np` (cx, cy) pairs -- row 0 = c0 for the Green-Gauss form,
// then one row per stencil entry k (the sliced-ELL entry k of every cell; missing entries are zero).
void fused_coeff_rows(const Layout &L, size_t np, std::vector<double> &rows);

}  // namespace fvs2d
