// kernels_fused.cuh -- ONE kernel per Runge-Kutta stage (option "fuse", single GPU, second-order upwind
// reconstruction without limiter): pass A (gradients) folded into the persistent pass-B pipeline of kernels.cuh.
//
// Why: the two-pass schedule writes the gradients (64 B/cell) in pass A and reads them back in pass B, and reads the
// primitive state twice -- 160 of the 512-528 algorithmic bytes of a stage ("B_alg - 128 B" is the
// single-pass lower bound).  Here a tile stages, besides what k_flux_pipe stages,
//   * the primitive state of "ring 2" (gradient-stencil members of the tile's cells and of its halo cells),
//   * the gradient operator (coefficient rows + a per-tile table of 16-bit stencil slots) of tile + halo cells,
// rebuilds the gradients of tile + halo in shared memory (phase 1; ~35 % redundant gradient evaluations, a few
// per cent of the flux arithmetic), and after a consumer-side barrier runs the face loop of k_flux_pipe on them
// (phase 2).  The gradient of a cell is evaluated with explicit fma() in one fixed order wherever it is rebuilt, so
// every copy is bit-identical and both evaluations of an interior face flux still agree bitwise (discrete
// conservation), and the result equals the two-pass path's bitwise as long as nvcc contracts k_gradient's
// `a += c*d` into the same fma (it does).
//
// Reference: src/gradient_ggcb.f90:116-138, src/gradient_ggnb.f90:183-210, src/gradient_lsq.f90:393-401 (phase 1);
// src/residual.f90:66-166, src/flux_invscid.f90:37-136, src/runge_kutta.f90:156-162,225-226,299-313,383-387 (phase 2).
//
// The face code is a copy of k_flux_pipe's RC_K0 branch on purpose: the production kernel stays untouched while this
// variant is being measured.
#pragma once
#include "kernels.cuh"

namespace fvs2d {

struct FusedMeta {
  const int4 *hdr;  // 3 x int4 per tile: {es, ne, hc_ptr, n1}, {he_ptr, n_he, fbase, fw}, {h2_ptr, n2, gs_base, gw}
  const int *hc_idx, *he_idx, *h2_idx;
  const uint32_t *t_pack;
  const int *t_bf;
  const uint16_t *gslot;  // per tile gw rows of pitch TW = roundup8(kBlock + n1)
  const double2 *gc2;     // gradient coefficients (cx, cy), rows of pitch np: [c0 (FORM 0)], entry 0, entry 1, ...
  int S1, S2, E, TW, W;   // smem pitches in elements: cells with gradient (tile + ring 1), all cells, edges; max TW; max entries
  int CG;                 // rows of the coefficient / gradient block = max(4, W + (FORM == 0))
  int ntiles;
};

// bytes of one stage:  p [2][S2] | coefficients -> gradients [CG][S1] | xy [S1] | exy, enxy [2][E] | ea [E] |
//                      face table [4][kBlock] | {fw, fbase, gw, n1} | stencil slots [W][TW] (16 bit)
__host__ __device__ inline size_t fused_stage_bytes(int S1, int S2, int E, int TW, int W, int CG) {
  return (size_t)2 * S2 * 16 + (size_t)CG * S1 * 16 + (size_t)S1 * 16 + (size_t)E * 40 + 4 * kBlock * sizeof(uint32_t) + 16 +
         (((size_t)W * TW * 2 + 15) & ~(size_t)15);
}

template <int UM, bool STEADY, int FORM, int CTAS>
__global__ void __launch_bounds__(kPipeThreads, CTAS) k_stage_fused(const DevMesh m, const FusedMeta fm, const Phys P, const StageParams S,
                                                                    const double *__restrict__ p, const double *__restrict__ bc,
                                                                    double *__restrict__ q, double *__restrict__ f,
                                                                    double *__restrict__ pout, double *__restrict__ dtl,
                                                                    double *__restrict__ partial) {
  constexpr int F0 = FORM == 0 ? 1 : 0;  // coefficient row 0 is c0 for the Green-Gauss form
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int S1 = fm.S1, S2 = fm.S2, EE = fm.E, CG = fm.CG, np = m.np;
  const size_t stage_bytes = fused_stage_bytes(S1, S2, EE, fm.TW, fm.W, CG);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + kStages * stage_bytes);
  uint64_t *empty = full + kStages;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  auto st_p = [&](int s) { return reinterpret_cast<double2 *>(smem_raw + s * stage_bytes); };
  auto st_cg = [&](int s) { return st_p(s) + 2 * S2; };
  auto st_xy = [&](int s) { return st_cg(s) + CG * S1; };
  auto st_e2 = [&](int s) { return st_xy(s) + S1; };
  auto st_ea = [&](int s) { return reinterpret_cast<double *>(st_e2(s) + 2 * EE); };
  auto st_f = [&](int s) { return reinterpret_cast<uint32_t *>(st_ea(s) + EE); };
  auto st_misc = [&](int s) { return reinterpret_cast<int *>(st_f(s) + 4 * kBlock); };
  auto st_gs = [&](int s) { return reinterpret_cast<uint16_t *>(st_misc(s) + 4); };
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStages; s++) { mbar_init(&full[s], 33); mbar_init(&empty[s], kBlock); }
  }
  __syncthreads();

  if (warp == kBlock / 32) {
    // ================================ producer warp ================================
    const double2 *p2 = reinterpret_cast<const double2 *>(p);
    int4 h0 = make_int4(0, 0, 0, 0), h1 = make_int4(0, 0, 0, 0), h2 = make_int4(0, 0, 0, 0);
    int jc[3] = {0, 0, 0}, je[3] = {0, 0, 0}, j2[3] = {0, 0, 0};
    auto fetch_meta = [&](int t) {
      h0 = __ldg(&fm.hdr[3 * t]);
      h1 = __ldg(&fm.hdr[3 * t + 1]);
      h2 = __ldg(&fm.hdr[3 * t + 2]);
#pragma unroll
      for (int r = 0; r < 3; r++) {
        jc[r] = (lane + 32 * r < h0.w) ? __ldg(&fm.hc_idx[h0.z + lane + 32 * r]) : 0;
        je[r] = (lane + 32 * r < h1.y) ? __ldg(&fm.he_idx[h1.x + lane + 32 * r]) : 0;
        j2[r] = (lane + 32 * r < h2.y) ? __ldg(&fm.h2_idx[h2.x + lane + 32 * r]) : 0;
      }
    };
    if ((int)blockIdx.x < fm.ntiles) fetch_meta(blockIdx.x);
    int it = 0;
    for (int t = blockIdx.x; t < fm.ntiles; t += gridDim.x, it++) {
      const int s = it % kStages;
      const uint32_t ph = (it / kStages) & 1;
      const int es = h0.x, ne = h0.y, hp = h0.z, n1 = h0.w, ep = h1.x, nhe = h1.y, fbase = h1.z, fw = h1.w;
      const int h2p = h2.x, n2 = h2.y, gsb = h2.z, gw = h2.w;
      const int rows = gw + F0;                  // coefficient rows of this tile
      const int tw = (kBlock + n1 + 7) & ~7;     // pitch of its stencil-slot table
      const int jcc[3] = {jc[0], jc[1], jc[2]}, jee[3] = {je[0], je[1], je[2]}, j22[3] = {j2[0], j2[1], j2[2]};
      mbar_wait(&empty[s], ph ^ 1);
      const int c0 = t * kBlock;
      const int ncell = min(kBlock, m.n_own - c0);
      double2 *sp = st_p(s), *scg = st_cg(s), *sxy = st_xy(s), *e2 = st_e2(s);
      if (lane == 0) {
        uint32_t *sf = st_f(s);
        int *sh = st_misc(s);
        sh[0] = fw; sh[1] = fbase; sh[2] = gw; sh[3] = n1;  // published to the consumers by the arrive below (release)
        const uint32_t bytes_c = (uint32_t)ncell * 16u;
        const uint32_t bytes_e = (uint32_t)ne * 16u, bytes_ea = (uint32_t)ne * 8u, bytes_f = (uint32_t)fw * kBlock * 4u;
        const uint32_t bytes_gs = (uint32_t)gw * (uint32_t)tw * 2u;
        mbar_expect_tx(&full[s], (3u + (uint32_t)rows) * bytes_c + 2u * bytes_e + bytes_ea + bytes_f + bytes_gs);
        bulk_g2s(sp, p2 + c0, bytes_c, &full[s]);
        bulk_g2s(sp + S2, p2 + (size_t)np + c0, bytes_c, &full[s]);
        bulk_g2s(sxy, m.xy + c0, bytes_c, &full[s]);
        for (int r = 0; r < rows; r++) bulk_g2s(scg + r * S1, fm.gc2 + (size_t)r * np + c0, bytes_c, &full[s]);
        if (ne > 0) {
          bulk_g2s(e2, m.exy + es, bytes_e, &full[s]);
          bulk_g2s(e2 + EE, m.enxy + es, bytes_e, &full[s]);
          bulk_g2s(st_ea(s), m.ea + es, bytes_ea, &full[s]);
        }
        if (fw > 0) bulk_g2s(sf, fm.t_pack + fbase, bytes_f, &full[s]);
        if (gw > 0) bulk_g2s(st_gs(s), fm.gslot + gsb, bytes_gs, &full[s]);
      }
      auto gather_h1 = [&](int h, int j) {  // ring 1: state, centroid, gradient operator
        cp_async16(sp + kBlock + h, p2 + j);
        cp_async16(sp + S2 + kBlock + h, p2 + (size_t)np + j);
        cp_async16(sxy + kBlock + h, m.xy + j);
        for (int r = 0; r < rows; r++) cp_async16(scg + r * S1 + kBlock + h, fm.gc2 + (size_t)r * np + j);
      };
      auto gather_h2 = [&](int h, int j) {  // ring 2: state only
        cp_async16(sp + kBlock + n1 + h, p2 + j);
        cp_async16(sp + S2 + kBlock + n1 + h, p2 + (size_t)np + j);
      };
      auto gather_edge = [&](int h, int j) {
        cp_async16(e2 + ne + h, m.exy + j);
        cp_async16(e2 + EE + ne + h, m.enxy + j);
        cp_async8(st_ea(s) + ne + h, m.ea + j);
      };
#pragma unroll
      for (int r = 0; r < 3; r++) {
        const int h = lane + 32 * r;
        if (h < n1) gather_h1(h, jcc[r]);
        if (h < n2) gather_h2(h, j22[r]);
        if (h < nhe) gather_edge(h, jee[r]);
      }
      for (int h = lane + 96; h < n1; h += 32) gather_h1(h, __ldg(&fm.hc_idx[hp + h]));  // rare: more than 96 entries
      for (int h = lane + 96; h < n2; h += 32) gather_h2(h, __ldg(&fm.h2_idx[h2p + h]));
      for (int h = lane + 96; h < nhe; h += 32) gather_edge(h, __ldg(&fm.he_idx[ep + h]));
      cp_async_mbar_arrive_noinc(&full[s]);
      if (t + (int)gridDim.x < fm.ntiles) fetch_meta(t + gridDim.x);
    }
    return;
  }

  // ================================== consumer warps ==================================
  double dq2[4] = {0.0, 0.0, 0.0, 0.0};
  int it = 0;
  for (int t = blockIdx.x; t < fm.ntiles; t += gridDim.x, it++) {
    const int s = it % kStages;
    const uint32_t ph = (it / kStages) & 1;
    const int c0 = t * kBlock;
    const int ncell = min(kBlock, m.n_own - c0);
    const int i = c0 + tid;
    const bool live = tid < ncell;
    double q0[4], fo[4], dl = 0.0, ivol = 1.0;
    if (live) {  // RK data of this cell: in flight while the gradients and the faces are computed
      stage_load<UM, STEADY>(S, i, np, q, f, dtl, q0, fo, dl);
      ivol = m.ivol[i];
    }
    const double2 *sp = st_p(s), *sxy = st_xy(s), *e2 = st_e2(s);
    double2 *scg = st_cg(s);
    const double *sea = st_ea(s);
    const uint32_t *sf = st_f(s);
    const uint16_t *sgs = st_gs(s);
    mbar_wait(&full[s], ph);
    const int fw = st_misc(s)[0], fbase = st_misc(s)[1], gw = st_misc(s)[2], n1 = st_misc(s)[3];
    const int tw = (kBlock + n1 + 7) & ~7;

    // ---- phase 1: gradients of the tile's cells and of ring 1, into the block that held their coefficients.
    // Column c of that block is read and then overwritten by this thread only; every coefficient of the column has
    // been consumed when the gradient is stored.
    for (int c = tid; c < kBlock + n1; c += kBlock) {
      if (c < kBlock && c >= ncell) continue;
      double p0[4], ax[4], ay[4];
      {
        const double2 a = sp[c], b = sp[S2 + c];
        p0[0] = a.x; p0[1] = a.y; p0[2] = b.x; p0[3] = b.y;
      }
      if (FORM == 0) {
        const double2 cc = scg[c];
#pragma unroll
        for (int v = 0; v < 4; v++) { ax[v] = cc.x * p0[v]; ay[v] = cc.y * p0[v]; }
      } else {
#pragma unroll
        for (int v = 0; v < 4; v++) { ax[v] = 0.0; ay[v] = 0.0; }
      }
      for (int k = 0; k < gw; k++) {
        const int js = sgs[k * tw + c];
        const double2 cf = scg[(k + F0) * S1 + c];
        const double2 a = sp[js], b = sp[S2 + js];
        const double pj[4] = {a.x, a.y, b.x, b.y};
#pragma unroll
        for (int v = 0; v < 4; v++) {
          const double d = FORM == 0 ? pj[v] : pj[v] - p0[v];
          ax[v] = fma(cf.x, d, ax[v]);
          ay[v] = fma(cf.y, d, ay[v]);
        }
      }
      scg[c] = make_double2(ax[0], ax[1]);
      scg[S1 + c] = make_double2(ax[2], ax[3]);
      scg[2 * S1 + c] = make_double2(ay[0], ay[1]);
      scg[3 * S1 + c] = make_double2(ay[2], ay[3]);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kBlock) : "memory");  // consumer warps only: all gradients are in place

    // ---- phase 2: faces (k_flux_pipe's RC_K0 path on the rebuilt gradients)
    double acc[4] = {0.0, 0.0, 0.0, 0.0}, wsacc = 0.0;
    if (live) {
      auto face = [&](const uint32_t pk, const int k, const auto bnd_tag) {
        constexpr bool BND = decltype(bnd_tag)::value;
        const int ns = pk & 0xFFFFu, eslot = (pk >> 16) & 0x7FFF;
        const bool self_c1 = BND || (pk >> 31) == 0;
        const double2 fc = e2[eslot], fn = e2[EE + eslot];
        const double af = sea[eslot], nx = fn.x, ny = fn.y;
        const int sl_ = self_c1 ? tid : ns, sr_ = (self_c1 && !BND) ? ns : tid;
        double sL[4], sR[4];
        {
          const double2 a = sp[sl_], b = sp[S2 + sl_], xl = sxy[sl_];
          const double2 ga = scg[sl_], gb = scg[S1 + sl_], gc = scg[2 * S1 + sl_], gd = scg[3 * S1 + sl_];
          const double dx = fc.x - xl.x, dy = fc.y - xl.y;
          sL[0] = recon_k0(a.x, ga.x, gc.x, dx, dy); sL[1] = recon_k0(a.y, ga.y, gc.y, dx, dy);
          sL[2] = recon_k0(b.x, gb.x, gd.x, dx, dy); sL[3] = recon_k0(b.y, gb.y, gd.y, dx, dy);
        }
        if (!BND) {
          const double2 a = sp[sr_], b = sp[S2 + sr_], xr = sxy[sr_];
          const double2 ga = scg[sr_], gb = scg[S1 + sr_], gc = scg[2 * S1 + sr_], gd = scg[3 * S1 + sr_];
          const double dx = fc.x - xr.x, dy = fc.y - xr.y;
          sR[0] = recon_k0(a.x, ga.x, gc.x, dx, dy); sR[1] = recon_k0(a.y, ga.y, gc.y, dx, dy);
          sR[2] = recon_k0(b.x, gb.x, gd.x, dx, dy); sR[3] = recon_k0(b.y, gb.y, gd.y, dx, dy);
        } else {
          const int b = __ldg(&fm.t_bf[fbase + k * kBlock + tid]);
          const int type = __ldg(&m.bf_type[b]);
          if (type == 2) {  // slip wall: mirror the normal velocity (src/residual.f90:200-204)
            const double un = sL[1] * nx + sL[2] * ny;
            sR[0] = sL[0]; sR[3] = sL[3];
            sR[1] = sL[1] - 2.0 * un * nx;
            sR[2] = sL[2] - 2.0 * un * ny;
          } else {
#pragma unroll
            for (int v = 0; v < 4; v++) sR[v] = __ldg(&bc[v * m.nbf + b]);
          }
        }
        double flux[4], ws;
        roe_flux2(P, sL, sR, nx, ny, flux, ws);
        const double ha = 0.5 * af, sa = self_c1 ? ha : -ha;
#pragma unroll
        for (int v = 0; v < 4; v++) acc[v] += flux[v] * sa;
        wsacc += ws * ha;
      };
      bool has_bnd = false;
      {
        // two interior faces per iteration in one basic block, so the two independent flux evaluations interleave
        int k = 0;
#pragma unroll 1
        for (; k + 1 < fw; k += 2) {
          const uint32_t pk0 = sf[k * kBlock + tid], pk1 = sf[(k + 1) * kBlock + tid];
          const uint32_t n0 = pk0 & 0xFFFFu, n1_ = pk1 & 0xFFFFu;
          has_bnd = has_bnd || n0 == 0xFFFFu || n1_ == 0xFFFFu;
          if (n0 < 0xFFFEu && n1_ < 0xFFFEu) {
            face(pk0, k, std::false_type{});
            face(pk1, k + 1, std::false_type{});
          } else {
#pragma unroll 1
            for (int h = 0; h < 2; h++) {
              const uint32_t pk = h ? pk1 : pk0;
              if ((pk & 0xFFFFu) < 0xFFFEu) face(pk, k + h, std::false_type{});
            }
          }
        }
        if (k < fw) {
          const uint32_t pk = sf[k * kBlock + tid], ns = pk & 0xFFFFu;
          if (ns == 0xFFFFu) has_bnd = true;
          else if (ns != 0xFFFEu) face(pk, k, std::false_type{});
        }
      }
      if (has_bnd) {
#pragma unroll 1
        for (int k = 0; k < fw; k++) {
          const uint32_t pk = sf[k * kBlock + tid];
          if ((pk & 0xFFFFu) == 0xFFFFu) face(pk, k, std::true_type{});
        }
      }
    }
    // this thread's gradient stores (generic proxy) are ordered before the bulk copies (async proxy) that refill the stage
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_arrive(&empty[s]);  // this thread is done with stage s
    if (live) stage_update_pre<UM, STEADY>(P, S, i, np, ivol, m.vol, q0, fo, dl, acc, wsacc, q, f, pout, dtl, nullptr, nullptr, dq2);
  }
  if (S.last) {
    // sum of (q - q0)^2 over this CTA's cells: warp shuffles, then the consumer warps through smem
    __shared__ double red[4][kBlock / 32];
#pragma unroll
    for (int v = 0; v < 4; v++) {
      double x = dq2[v];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
      if (lane == 0) red[v][warp] = x;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kBlock) : "memory");
    if (tid < 4) {
      double ssum = 0.0;
#pragma unroll
      for (int w = 0; w < kBlock / 32; w++) ssum += red[tid][w];
      partial[blockIdx.x * 4 + tid] = ssum;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Second variant ("fuse" = 2): every reconstructed face state is evaluated ONCE, by the thread that holds the cell's
// gradient in registers, and published in shared memory; the flux phase then reads two 32-byte states per face (its
// own: conflict-free; the neighbour's: one scattered slot) instead of gathering state, gradient and centroid of both
// cells (14 scattered 16-byte loads per face in k_flux_pipe / k_stage_fused, the busiest pipe of those kernels).
//   phase 1a  gradients: own cell -> registers; ring-1 cells -> in place over their coefficients (as k_stage_fused)
//   barrier   (the tile's own-cell region X -- state, centroid, coefficient rows -- is dead from here on)
//   phase 1b  publish: own[k][tid] = state of this cell at its face k (into X); hst[i] = state of the ring-1 cell at
//             the i-th tile/ring-1 face (list fz_hf, one entry per thread, into X behind the own states)
//   barrier
//   phase 2   per face: own[k][tid], and own[k'][ns] (k' = index of the face in the neighbour's list, from the face
//             word) or hst[i]; unit normal and length of the edge; Roe flux; accumulate.  Stage update as before.
// Stage layout: X [XR][kBlock] double2 (rows 0,1 state; 2 centroid; 3.. coefficient rows) | state of rings 1+2
// [2][HP] | coefficients -> gradients of ring 1 [CG][H1] | centroids of ring 1 [H1] | exy, enxy [2][E] | ea [E] |
// face words [4][kBlock] | 8 ints | stencil slots [W][TW] u16 | tile/ring-1 faces [HF] u32
struct Fused2Meta {
  const int4 *hdr;  // 4 x int4 per tile: k_stage_fused's three + {hf_ptr, n_hf, uf_ptr, n_uf}
  const int *hc_idx, *he_idx, *h2_idx;
  const uint32_t *pack2, *hf;
  const uint2 *uf;  // unique faces of all tiles (VAR 3), see layout.hpp
  const int *t_bf;
  const uint16_t *gslot;
  const double2 *gc2;
  int H1, HP, E, TW, W, CG, XR, FW, HF;  // pitches: ring 1, rings 1+2, edges, slot table; max stencil entries; rows of the
                                         // ring-1 block and of X; max faces per cell; max tile/ring-1 faces (multiple of 4)
  int ntiles;
  const int *tile_list;  // null: tiles 0..ntiles-1; else the ntiles tile ids to process (a launch over a subset of the tiles)
};
__host__ __device__ inline size_t fused2_stage_bytes(const Fused2Meta &f) {
  return (size_t)f.XR * kBlock * 16 + (size_t)2 * f.HP * 16 + (size_t)f.CG * f.H1 * 16 + (size_t)f.H1 * 16 + (size_t)f.E * 40 +
         4 * kBlock * sizeof(uint32_t) + 32 + (((size_t)f.W * f.TW * 2 + 15) & ~(size_t)15) + (size_t)f.HF * 4;
}

// VAR 3 ("fuse" = 3) additionally evaluates every face flux of the tile ONCE: after the states are published, the
// tile's unique faces (list fz_uf: tile/tile faces once, tile/ring-1 faces, boundary faces) are dealt out evenly to the
// threads, two per thread and round -- a triangle tile has ~215 unique faces against 384 cell-faces, so the flux
// arithmetic (the dominant cost: ~190 dependent fp64 instructions per face) drops by ~40 % and a thread's faces fit one
// interleaved pair instead of a pair plus a single.  The flux is written IN PLACE over the two states it consumed (each
// published state belongs to exactly one unique face, so no other thread reads them), in the edge's own orientation;
// after a third barrier every thread sums the fluxes of its cell in face order with the sign of its side -- the same
// values in the same order as the other kernels, so the result is still bitwise theirs.  The wave speeds (needed by
// the local time step only) go to the dead ring-state block.
template <int UM, bool STEADY, int FORM, int CTAS, int VAR>
__global__ void __launch_bounds__(kPipeThreads, CTAS) k_stage_fused2(const DevMesh m, const Fused2Meta fm, const Phys P, const StageParams S,
                                                                     const double *__restrict__ p, const double *__restrict__ bc,
                                                                     double *__restrict__ q, double *__restrict__ f,
                                                                     double *__restrict__ pout, double *__restrict__ dtl,
                                                                     double *__restrict__ partial) {
  constexpr int F0 = FORM == 0 ? 1 : 0;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int H1 = fm.H1, HP = fm.HP, EE = fm.E, CG = fm.CG, np = m.np;
  const size_t stage_bytes = fused2_stage_bytes(fm);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + kStages * stage_bytes);
  uint64_t *empty = full + kStages;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  auto st_x = [&](int s) { return reinterpret_cast<double2 *>(smem_raw + s * stage_bytes); };
  auto st_ph = [&](int s) { return st_x(s) + fm.XR * kBlock; };
  auto st_cg = [&](int s) { return st_ph(s) + 2 * HP; };
  auto st_xy = [&](int s) { return st_cg(s) + CG * H1; };
  auto st_e2 = [&](int s) { return st_xy(s) + H1; };
  auto st_ea = [&](int s) { return reinterpret_cast<double *>(st_e2(s) + 2 * EE); };
  auto st_f = [&](int s) { return reinterpret_cast<uint32_t *>(st_ea(s) + EE); };
  auto st_misc = [&](int s) { return reinterpret_cast<int *>(st_f(s) + 4 * kBlock); };
  auto st_gs = [&](int s) { return reinterpret_cast<uint16_t *>(st_misc(s) + 8); };
  auto st_hf = [&](int s) { return reinterpret_cast<uint32_t *>(reinterpret_cast<unsigned char *>(st_gs(s)) + (((size_t)fm.W * fm.TW * 2 + 15) & ~(size_t)15)); };
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStages; s++) { mbar_init(&full[s], 33); mbar_init(&empty[s], kBlock); }
  }
  __syncthreads();

  if (warp == kBlock / 32) {
    // ================================ producer warp ================================
    const double2 *p2 = reinterpret_cast<const double2 *>(p);
    int4 h0 = make_int4(0, 0, 0, 0), h1 = make_int4(0, 0, 0, 0), h2 = make_int4(0, 0, 0, 0), h3 = make_int4(0, 0, 0, 0);
    int jc[3] = {0, 0, 0}, je[3] = {0, 0, 0}, j2[3] = {0, 0, 0};
    auto fetch_meta = [&](int t) {
      h0 = __ldg(&fm.hdr[4 * t]);
      h1 = __ldg(&fm.hdr[4 * t + 1]);
      h2 = __ldg(&fm.hdr[4 * t + 2]);
      h3 = __ldg(&fm.hdr[4 * t + 3]);
#pragma unroll
      for (int r = 0; r < 3; r++) {
        jc[r] = (lane + 32 * r < h0.w) ? __ldg(&fm.hc_idx[h0.z + lane + 32 * r]) : 0;
        je[r] = (lane + 32 * r < h1.y) ? __ldg(&fm.he_idx[h1.x + lane + 32 * r]) : 0;
        j2[r] = (lane + 32 * r < h2.y) ? __ldg(&fm.h2_idx[h2.x + lane + 32 * r]) : 0;
      }
    };
    auto tile_id = [&](int j) { return fm.tile_list ? __ldg(&fm.tile_list[j]) : j; };
    if ((int)blockIdx.x < fm.ntiles) fetch_meta(tile_id(blockIdx.x));
    int it = 0;
    for (int j = blockIdx.x; j < fm.ntiles; j += gridDim.x, it++) {
      const int t = tile_id(j);
      const int s = it % kStages;
      const uint32_t ph = (it / kStages) & 1;
      const int es = h0.x, ne = h0.y, hp = h0.z, n1 = h0.w, ep = h1.x, nhe = h1.y, fbase = h1.z, fw = h1.w;
      const int h2p = h2.x, n2 = h2.y, gsb = h2.z, gw = h2.w, hfp = h3.x, nhf = h3.y;
      const int rows = gw + F0;
      const int tw = (kBlock + n1 + 7) & ~7;
      const int jcc[3] = {jc[0], jc[1], jc[2]}, jee[3] = {je[0], je[1], je[2]}, j22[3] = {j2[0], j2[1], j2[2]};
      mbar_wait(&empty[s], ph ^ 1);
      const int c0 = t * kBlock;
      const int ncell = min(kBlock, m.n_own - c0);
      double2 *sx = st_x(s), *sph = st_ph(s), *scg = st_cg(s), *sxy = st_xy(s), *e2 = st_e2(s);
      if (lane == 0) {
        int *sh = st_misc(s);
        sh[0] = fw; sh[1] = fbase; sh[2] = gw; sh[3] = n1; sh[4] = nhf;  // published by the arrive below (release)
        const uint32_t bytes_c = (uint32_t)ncell * 16u;
        const uint32_t bytes_e = (uint32_t)ne * 16u, bytes_ea = (uint32_t)ne * 8u, bytes_f = (uint32_t)fw * kBlock * 4u;
        const uint32_t bytes_gs = (uint32_t)gw * (uint32_t)tw * 2u, bytes_hf = (uint32_t)((nhf + 3) & ~3) * 4u;
        mbar_expect_tx(&full[s], (3u + (uint32_t)rows) * bytes_c + 2u * bytes_e + bytes_ea + bytes_f + bytes_gs + bytes_hf);
        bulk_g2s(sx, p2 + c0, bytes_c, &full[s]);
        bulk_g2s(sx + kBlock, p2 + (size_t)np + c0, bytes_c, &full[s]);
        bulk_g2s(sx + 2 * kBlock, m.xy + c0, bytes_c, &full[s]);
        for (int r = 0; r < rows; r++) bulk_g2s(sx + (3 + r) * kBlock, fm.gc2 + (size_t)r * np + c0, bytes_c, &full[s]);
        if (ne > 0) {
          bulk_g2s(e2, m.exy + es, bytes_e, &full[s]);
          bulk_g2s(e2 + EE, m.enxy + es, bytes_e, &full[s]);
          bulk_g2s(st_ea(s), m.ea + es, bytes_ea, &full[s]);
        }
        if (fw > 0) bulk_g2s(st_f(s), fm.pack2 + fbase, bytes_f, &full[s]);
        if (gw > 0) bulk_g2s(st_gs(s), fm.gslot + gsb, bytes_gs, &full[s]);
        if (nhf > 0) bulk_g2s(st_hf(s), fm.hf + hfp, bytes_hf, &full[s]);
      }
      auto gather_h1 = [&](int h, int j) {  // ring 1: state, centroid, gradient operator
        cp_async16(sph + h, p2 + j);
        cp_async16(sph + HP + h, p2 + (size_t)np + j);
        cp_async16(sxy + h, m.xy + j);
        for (int r = 0; r < rows; r++) cp_async16(scg + r * H1 + h, fm.gc2 + (size_t)r * np + j);
      };
      auto gather_h2 = [&](int h, int j) {  // ring 2: state only
        cp_async16(sph + n1 + h, p2 + j);
        cp_async16(sph + HP + n1 + h, p2 + (size_t)np + j);
      };
      auto gather_edge = [&](int h, int j) {
        cp_async16(e2 + ne + h, m.exy + j);
        cp_async16(e2 + EE + ne + h, m.enxy + j);
        cp_async8(st_ea(s) + ne + h, m.ea + j);
      };
#pragma unroll
      for (int r = 0; r < 3; r++) {
        const int h = lane + 32 * r;
        if (h < n1) gather_h1(h, jcc[r]);
        if (h < n2) gather_h2(h, j22[r]);
        if (h < nhe) gather_edge(h, jee[r]);
      }
      for (int h = lane + 96; h < n1; h += 32) gather_h1(h, __ldg(&fm.hc_idx[hp + h]));
      for (int h = lane + 96; h < n2; h += 32) gather_h2(h, __ldg(&fm.h2_idx[h2p + h]));
      for (int h = lane + 96; h < nhe; h += 32) gather_edge(h, __ldg(&fm.he_idx[ep + h]));
      cp_async_mbar_arrive_noinc(&full[s]);
      if (j + (int)gridDim.x < fm.ntiles) fetch_meta(tile_id(j + gridDim.x));
    }
    return;
  }

  // ================================== consumer warps ==================================
  double dq2[4] = {0.0, 0.0, 0.0, 0.0};
  const int hst0 = 2 * fm.FW * kBlock;  // first double2 of the ring-1 face states inside X
  int it = 0;
  for (int j = blockIdx.x; j < fm.ntiles; j += gridDim.x, it++) {
    const int t = fm.tile_list ? __ldg(&fm.tile_list[j]) : j;
    const int s = it % kStages;
    const uint32_t ph = (it / kStages) & 1;
    const int c0 = t * kBlock;
    const int ncell = min(kBlock, m.n_own - c0);
    const int i = c0 + tid;
    const bool live = tid < ncell;
    double q0[4], fo[4], dl = 0.0, ivol = 1.0;
    if (live) {
      stage_load<UM, STEADY>(S, i, np, q, f, dtl, q0, fo, dl);
      ivol = m.ivol[i];
    }
    uint2 u0 = make_uint2(0u, 0u), u1 = make_uint2(0u, 0u);
    int ufb = 0, nuf = 0;
    if (VAR == 3) {  // this thread's first two unique faces: in flight like the RK data
      const int4 h3 = __ldg(&fm.hdr[4 * t + 3]);
      ufb = h3.z; nuf = h3.w;
      if (tid < nuf) u0 = __ldg(&fm.uf[ufb + tid]);
      if (tid + kBlock < nuf) u1 = __ldg(&fm.uf[ufb + tid + kBlock]);
    }
    double2 *sx = st_x(s), *scg = st_cg(s);
    const double2 *sph = st_ph(s), *sxy = st_xy(s), *e2 = st_e2(s);
    const double *sea = st_ea(s);
    const uint32_t *sf = st_f(s), *shf = st_hf(s);
    const uint16_t *sgs = st_gs(s);
    mbar_wait(&full[s], ph);
    const int fw = st_misc(s)[0], fbase = st_misc(s)[1], gw = st_misc(s)[2], n1 = st_misc(s)[3], nhf = st_misc(s)[4];
    const int tw = (kBlock + n1 + 7) & ~7;
    // state of the cell in slot js (own cells in X, rings 1 and 2 behind it)
    auto ld_p = [&](int js, double pj[4]) {
      const double2 *a = js < kBlock ? sx + js : sph + (js - kBlock);
      const int pitch = js < kBlock ? kBlock : HP;
      const double2 u = a[0], w = a[pitch];
      pj[0] = u.x; pj[1] = u.y; pj[2] = w.x; pj[3] = w.y;
    };

    // ---- phase 1a: gradients
    double p0[4] = {0.0, 0.0, 0.0, 0.0}, gx[4] = {0.0, 0.0, 0.0, 0.0}, gy[4] = {0.0, 0.0, 0.0, 0.0};
    double2 xc = make_double2(0.0, 0.0);
    if (live) {  // own cell: result stays in registers
      ld_p(tid, p0);
      xc = sx[2 * kBlock + tid];
      if (FORM == 0) {
        const double2 cc = sx[3 * kBlock + tid];
#pragma unroll
        for (int v = 0; v < 4; v++) { gx[v] = cc.x * p0[v]; gy[v] = cc.y * p0[v]; }
      }
      for (int k = 0; k < gw; k++) {
        const int js = sgs[k * tw + tid];
        const double2 cf = sx[(3 + k + F0) * kBlock + tid];
        double pj[4];
        ld_p(js, pj);
#pragma unroll
        for (int v = 0; v < 4; v++) {
          const double d = FORM == 0 ? pj[v] : pj[v] - p0[v];
          gx[v] = fma(cf.x, d, gx[v]);
          gy[v] = fma(cf.y, d, gy[v]);
        }
      }
    }
    for (int h = tid; h < n1; h += kBlock) {  // ring 1: in place over the column's coefficients (column-private)
      double ph0[4], ax[4], ay[4];
      ld_p(kBlock + h, ph0);
      if (FORM == 0) {
        const double2 cc = scg[h];
#pragma unroll
        for (int v = 0; v < 4; v++) { ax[v] = cc.x * ph0[v]; ay[v] = cc.y * ph0[v]; }
      } else {
#pragma unroll
        for (int v = 0; v < 4; v++) { ax[v] = 0.0; ay[v] = 0.0; }
      }
      for (int k = 0; k < gw; k++) {
        const int js = sgs[k * tw + kBlock + h];
        const double2 cf = scg[(k + F0) * H1 + h];
        double pj[4];
        ld_p(js, pj);
#pragma unroll
        for (int v = 0; v < 4; v++) {
          const double d = FORM == 0 ? pj[v] : pj[v] - ph0[v];
          ax[v] = fma(cf.x, d, ax[v]);
          ay[v] = fma(cf.y, d, ay[v]);
        }
      }
      scg[h] = make_double2(ax[0], ax[1]);
      scg[H1 + h] = make_double2(ax[2], ax[3]);
      scg[2 * H1 + h] = make_double2(ay[0], ay[1]);
      scg[3 * H1 + h] = make_double2(ay[2], ay[3]);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kBlock) : "memory");  // X is dead, ring-1 gradients are in place

    // ---- phase 1b: publish the reconstructed face states
    if (live) {
      for (int k = 0; k < fw; k++) {
        const uint32_t pk = sf[k * kBlock + tid];
        if ((pk & 0xFFFFu) == 0xFFFEu) continue;
        const double2 fc = e2[(pk >> 16) & 0xFFFu];
        const double dx = fc.x - xc.x, dy = fc.y - xc.y;
        sx[(2 * k) * kBlock + tid] = make_double2(recon_k0(p0[0], gx[0], gy[0], dx, dy), recon_k0(p0[1], gx[1], gy[1], dx, dy));
        sx[(2 * k + 1) * kBlock + tid] = make_double2(recon_k0(p0[2], gx[2], gy[2], dx, dy), recon_k0(p0[3], gx[3], gy[3], dx, dy));
      }
    }
    for (int e = tid; e < nhf; e += kBlock) {
      const uint32_t w = shf[e];
      const int h = w & 0xFFFFu;
      const double2 fc = e2[w >> 16], xh = sxy[h];
      const double2 a = sph[h], b = sph[HP + h];
      const double2 ga = scg[h], gb = scg[H1 + h], gc = scg[2 * H1 + h], gd = scg[3 * H1 + h];
      const double dx = fc.x - xh.x, dy = fc.y - xh.y;
      sx[hst0 + 2 * e] = make_double2(recon_k0(a.x, ga.x, gc.x, dx, dy), recon_k0(a.y, ga.y, gc.y, dx, dy));
      sx[hst0 + 2 * e + 1] = make_double2(recon_k0(b.x, gb.x, gd.x, dx, dy), recon_k0(b.y, gb.y, gd.y, dx, dy));
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kBlock) : "memory");

    // ---- phase 2: faces
    double acc[4] = {0.0, 0.0, 0.0, 0.0}, wsacc = 0.0;
    if (VAR == 2 && live) {
      auto face = [&](const uint32_t pk, const int k, const auto bnd_tag) {
        constexpr bool BND = decltype(bnd_tag)::value;
        const int code = pk & 0xFFFFu, eslot = (pk >> 16) & 0xFFFu, kr = (pk >> 28) & 3;
        const bool self_c1 = BND || (pk >> 31) == 0;
        const double2 fn = e2[EE + eslot];
        const double af = sea[eslot], nx = fn.x, ny = fn.y;
        const int own0 = (2 * k) * kBlock + tid, own1 = own0 + kBlock;
        double sL[4], sR[4];
        if (!BND) {
          const int nb0 = code < kBlock ? (2 * kr) * kBlock + code : hst0 + 2 * (code - kBlock);
          const int nb1 = code < kBlock ? nb0 + kBlock : nb0 + 1;
          const double2 a = sx[self_c1 ? own0 : nb0], b = sx[self_c1 ? own1 : nb1];
          const double2 c = sx[self_c1 ? nb0 : own0], d = sx[self_c1 ? nb1 : own1];
          sL[0] = a.x; sL[1] = a.y; sL[2] = b.x; sL[3] = b.y;
          sR[0] = c.x; sR[1] = c.y; sR[2] = d.x; sR[3] = d.y;
        } else {
          const double2 a = sx[own0], b2 = sx[own1];
          sL[0] = a.x; sL[1] = a.y; sL[2] = b2.x; sL[3] = b2.y;
          const int b = __ldg(&fm.t_bf[fbase + k * kBlock + tid]);
          const int type = __ldg(&m.bf_type[b]);
          if (type == 2) {  // slip wall: mirror the normal velocity (src/residual.f90:200-204)
            const double un = sL[1] * nx + sL[2] * ny;
            sR[0] = sL[0]; sR[3] = sL[3];
            sR[1] = sL[1] - 2.0 * un * nx;
            sR[2] = sL[2] - 2.0 * un * ny;
          } else {
#pragma unroll
            for (int v = 0; v < 4; v++) sR[v] = __ldg(&bc[v * m.nbf + b]);
          }
        }
        double flux[4], ws;
        roe_flux2(P, sL, sR, nx, ny, flux, ws);
        const double ha = 0.5 * af, sa = self_c1 ? ha : -ha;
#pragma unroll
        for (int v = 0; v < 4; v++) acc[v] += flux[v] * sa;
        wsacc += ws * ha;
      };
      bool has_bnd = false;
      {
        int k = 0;
#pragma unroll 1
        for (; k + 1 < fw; k += 2) {
          const uint32_t pk0 = sf[k * kBlock + tid], pk1 = sf[(k + 1) * kBlock + tid];
          const uint32_t n0 = pk0 & 0xFFFFu, n1_ = pk1 & 0xFFFFu;
          has_bnd = has_bnd || n0 == 0xFFFFu || n1_ == 0xFFFFu;
          if (n0 < 0xFFFEu && n1_ < 0xFFFEu) {
            face(pk0, k, std::false_type{});
            face(pk1, k + 1, std::false_type{});
          } else {
#pragma unroll 1
            for (int h = 0; h < 2; h++) {
              const uint32_t pk = h ? pk1 : pk0;
              if ((pk & 0xFFFFu) < 0xFFFEu) face(pk, k + h, std::false_type{});
            }
          }
        }
        if (k < fw) {
          const uint32_t pk = sf[k * kBlock + tid], ns = pk & 0xFFFFu;
          if (ns == 0xFFFFu) has_bnd = true;
          else if (ns != 0xFFFEu) face(pk, k, std::false_type{});
        }
      }
      if (has_bnd) {
#pragma unroll 1
        for (int k = 0; k < fw; k++) {
          const uint32_t pk = sf[k * kBlock + tid];
          if ((pk & 0xFFFFu) == 0xFFFFu) face(pk, k, std::true_type{});
        }
      }
    }
    if (VAR == 3) {
      // ---- phase 2a: the tile's unique faces, two per thread and round
      double *sws = reinterpret_cast<double *>(st_ph(s));  // wave speeds [fw][kBlock] over the dead ring-state block
      auto st_addr = [&](const int loc) { return loc < 0x400 ? ((loc / kBlock) * 2 * kBlock + (loc % kBlock)) : hst0 + 2 * (loc - 0x400); };
      auto uface = [&](const uint2 u, const auto bnd_tag) {
        constexpr bool BND = decltype(bnd_tag)::value;
        const int locL = u.x & 0xFFFFu, locR = u.x >> 16, eslot = u.y & 0xFFFu, outL = (u.y >> 12) & 0x3FFu, outR = u.y >> 22;
        const double2 fn = e2[EE + eslot];
        const double nx = fn.x, ny = fn.y;
        double sL[4], sR[4];
        {
          const int a0 = st_addr(locL), a1 = a0 + (locL < 0x400 ? kBlock : 1);
          const double2 a = sx[a0], b = sx[a1];
          sL[0] = a.x; sL[1] = a.y; sL[2] = b.x; sL[3] = b.y;
        }
        if (!BND) {
          const int a0 = st_addr(locR), a1 = a0 + (locR < 0x400 ? kBlock : 1);
          const double2 c = sx[a0], d = sx[a1];
          sR[0] = c.x; sR[1] = c.y; sR[2] = d.x; sR[3] = d.y;
        } else {
          const int b = __ldg(&fm.t_bf[fbase + outL]);
          const int type = __ldg(&m.bf_type[b]);
          if (type == 2) {  // slip wall: mirror the normal velocity (src/residual.f90:200-204)
            const double un = sL[1] * nx + sL[2] * ny;
            sR[0] = sL[0]; sR[3] = sL[3];
            sR[1] = sL[1] - 2.0 * un * nx;
            sR[2] = sL[2] - 2.0 * un * ny;
          } else {
#pragma unroll
            for (int v = 0; v < 4; v++) sR[v] = __ldg(&bc[v * m.nbf + b]);
          }
        }
        double flux[4], ws;
        roe_flux2(P, sL, sR, nx, ny, flux, ws);
        if (outL != 0x3FF) {
          const int o = (outL / kBlock) * 2 * kBlock + (outL % kBlock);
          sx[o] = make_double2(flux[0], flux[1]);
          sx[o + kBlock] = make_double2(flux[2], flux[3]);
          if (STEADY) sws[outL] = ws;
        }
        if (!BND && outR != 0x3FF) {
          const int o = (outR / kBlock) * 2 * kBlock + (outR % kBlock);
          sx[o] = make_double2(flux[0], flux[1]);
          sx[o + kBlock] = make_double2(flux[2], flux[3]);
          if (STEADY) sws[outR] = ws;
        }
      };
#pragma unroll 1
      for (int e0 = tid; e0 < nuf; e0 += 2 * kBlock) {
        const bool first = e0 == tid, two = e0 + kBlock < nuf;
        const uint2 ua = first ? u0 : __ldg(&fm.uf[ufb + e0]);
        const uint2 ub = !two ? ua : first ? u1 : __ldg(&fm.uf[ufb + e0 + kBlock]);
        const bool ba = (ua.x >> 16) == 0xFFFFu, bb = (ub.x >> 16) == 0xFFFFu;
        if (two && !ba && !bb) {  // one basic block: the two independent flux evaluations interleave
          uface(ua, std::false_type{});
          uface(ub, std::false_type{});
        } else {
          if (ba) uface(ua, std::true_type{}); else uface(ua, std::false_type{});
          if (two) { if (bb) uface(ub, std::true_type{}); else uface(ub, std::false_type{}); }
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kBlock) : "memory");  // every flux of the tile is in place
      // ---- phase 2b: this cell's fluxes, in face order, with the sign of its side
      if (live) {
        for (int k = 0; k < fw; k++) {
          const uint32_t pk = sf[k * kBlock + tid];
          if ((pk & 0xFFFFu) == 0xFFFEu) continue;
          const double2 fa = sx[(2 * k) * kBlock + tid], fb = sx[(2 * k + 1) * kBlock + tid];
          const double ha = 0.5 * sea[(pk >> 16) & 0xFFFu];
          const double sa = (pk >> 31) == 0 ? ha : -ha;
          acc[0] = fma(fa.x, sa, acc[0]); acc[1] = fma(fa.y, sa, acc[1]);
          acc[2] = fma(fb.x, sa, acc[2]); acc[3] = fma(fb.y, sa, acc[3]);
          if (STEADY) wsacc = fma(sws[k * kBlock + tid], ha, wsacc);
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores before the refilling bulk copies
    mbar_arrive(&empty[s]);
    if (live) stage_update_pre<UM, STEADY>(P, S, i, np, ivol, m.vol, q0, fo, dl, acc, wsacc, q, f, pout, dtl, nullptr, nullptr, dq2);
  }
  if (S.last) {
    __shared__ double red[4][kBlock / 32];
#pragma unroll
    for (int v = 0; v < 4; v++) {
      double x = dq2[v];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
      if (lane == 0) red[v][warp] = x;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kBlock) : "memory");
    if (tid < 4) {
      double ssum = 0.0;
#pragma unroll
      for (int w = 0; w < kBlock / 32; w++) ssum += red[tid][w];
      partial[blockIdx.x * 4 + tid] = ssum;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// "fuse" = 5: k_stage_fused2<2> on a shared-memory diet, so that tiles of QUADRILATERALS also fit three CTAs per SM
// (C4: 97 KB -> 68 KB per CTA).  Everything only the owning thread reads comes straight from global memory into its
// registers, issued with the cell's RK data at the top of a tile and consumed after the stage's mbarrier wait:
// the own cell's coefficient rows (<= 5: Green-Gauss / least squares over face neighbours), its stencil slots, its face
// words, and its face displacements x_f - x_c (a new array fdxy[k][np], bitwise the difference the other kernels form).
// The published states live in ONE buffer X per CTA that the producer never touches (the barrier after phase 1a of the
// next tile already orders its reuse), so only the gathered data is double-buffered:
//   stage:  state of own cells [2][kBlock] | state of rings 1+2 [2][HP] | coefficients -> gradients of ring 1 [CG][H1] |
//           enxy [E] | ea [E] | 8 ints | stencil slots of ring 1 [W][TW-kBlock] u16 | tile/ring-1 faces [HF] u32 |
//           their displacements [HF] double2
//   X:      own[k][kBlock] states (2 double2 each) | ring-1 face states [HF]
struct Fused2cMeta {
  const int4 *hdr;  // as Fused2Meta
  const int *hc_idx, *he_idx, *h2_idx;
  const uint32_t *pack2, *hf;
  const double2 *hfd;   // per tile/ring-1 face: x_f - x_c of the ring-1 cell (same offsets as hf)
  const double2 *fdxy;  // [k][np]: x_f - x_c of face k of a cell
  const int *t_bf;
  const uint16_t *gslot;
  const double2 *gc2;
  int H1, HP, E, TW, W, CG, FW, HF;
  int ntiles;
  const int *tile_list;
};
constexpr int kF2cRows = 5;  // own coefficient rows held in registers (W + (FORM == 0) <= 5)
__host__ __device__ inline size_t fused2c_stage_bytes(const Fused2cMeta &f) {
  return (size_t)2 * kBlock * 16 + (size_t)2 * f.HP * 16 + (size_t)f.CG * f.H1 * 16 + (size_t)f.E * 24 + 32 +
         (((size_t)f.W * (f.TW - kBlock) * 2 + 15) & ~(size_t)15) + (size_t)f.HF * 4 + (size_t)f.HF * 16;
}
__host__ __device__ inline size_t fused2c_x_bytes(const Fused2cMeta &f) { return (size_t)2 * f.FW * kBlock * 16 + (size_t)f.HF * 32; }

template <int UM, bool STEADY, int FORM, int CTAS>
__global__ void __launch_bounds__(kPipeThreads, CTAS) k_stage_fused2c(const DevMesh m, const Fused2cMeta fm, const Phys P, const StageParams S,
                                                                      const double *__restrict__ p, const double *__restrict__ bc,
                                                                      double *__restrict__ q, double *__restrict__ f,
                                                                      double *__restrict__ pout, double *__restrict__ dtl,
                                                                      double *__restrict__ partial) {
  constexpr int F0 = FORM == 0 ? 1 : 0;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int H1 = fm.H1, HP = fm.HP, EE = fm.E, CG = fm.CG, np = m.np;
  const size_t stage_bytes = fused2c_stage_bytes(fm);
  double2 *sx = reinterpret_cast<double2 *>(smem_raw);  // X: consumer-only scratch, one per CTA
  unsigned char *stages = smem_raw + fused2c_x_bytes(fm);
  uint64_t *full = reinterpret_cast<uint64_t *>(stages + kStages * stage_bytes);
  uint64_t *empty = full + kStages;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  auto st_pa = [&](int s) { return reinterpret_cast<double2 *>(stages + s * stage_bytes); };
  auto st_ph = [&](int s) { return st_pa(s) + 2 * kBlock; };
  auto st_cg = [&](int s) { return st_ph(s) + 2 * HP; };
  auto st_en = [&](int s) { return st_cg(s) + CG * H1; };
  auto st_ea = [&](int s) { return reinterpret_cast<double *>(st_en(s) + EE); };
  auto st_misc = [&](int s) { return reinterpret_cast<int *>(st_ea(s) + EE); };
  auto st_gs = [&](int s) { return reinterpret_cast<uint16_t *>(st_misc(s) + 8); };
  auto st_hf = [&](int s) { return reinterpret_cast<uint32_t *>(reinterpret_cast<unsigned char *>(st_gs(s)) + (((size_t)fm.W * (fm.TW - kBlock) * 2 + 15) & ~(size_t)15)); };
  auto st_hfd = [&](int s) { return reinterpret_cast<double2 *>(st_hf(s) + fm.HF); };
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStages; s++) { mbar_init(&full[s], 33); mbar_init(&empty[s], kBlock); }
  }
  __syncthreads();

  if (warp == kBlock / 32) {
    // ================================ producer warp ================================
    const double2 *p2 = reinterpret_cast<const double2 *>(p);
    int4 h0 = make_int4(0, 0, 0, 0), h1 = make_int4(0, 0, 0, 0), h2 = make_int4(0, 0, 0, 0), h3 = make_int4(0, 0, 0, 0);
    int jc[3] = {0, 0, 0}, je[3] = {0, 0, 0}, j2[3] = {0, 0, 0};
    auto fetch_meta = [&](int t) {
      h0 = __ldg(&fm.hdr[4 * t]);
      h1 = __ldg(&fm.hdr[4 * t + 1]);
      h2 = __ldg(&fm.hdr[4 * t + 2]);
      h3 = __ldg(&fm.hdr[4 * t + 3]);
#pragma unroll
      for (int r = 0; r < 3; r++) {
        jc[r] = (lane + 32 * r < h0.w) ? __ldg(&fm.hc_idx[h0.z + lane + 32 * r]) : 0;
        je[r] = (lane + 32 * r < h1.y) ? __ldg(&fm.he_idx[h1.x + lane + 32 * r]) : 0;
        j2[r] = (lane + 32 * r < h2.y) ? __ldg(&fm.h2_idx[h2.x + lane + 32 * r]) : 0;
      }
    };
    auto tile_id = [&](int j) { return fm.tile_list ? __ldg(&fm.tile_list[j]) : j; };
    if ((int)blockIdx.x < fm.ntiles) fetch_meta(tile_id(blockIdx.x));
    int it = 0;
    for (int j = blockIdx.x; j < fm.ntiles; j += gridDim.x, it++) {
      const int t = tile_id(j);
      const int s = it % kStages;
      const uint32_t ph = (it / kStages) & 1;
      const int es = h0.x, ne = h0.y, hp = h0.z, n1 = h0.w, ep = h1.x, nhe = h1.y;
      const int h2p = h2.x, n2 = h2.y, gsb = h2.z, gw = h2.w, hfp = h3.x, nhf = h3.y;
      const int rows = gw + F0;
      const int tw = (kBlock + n1 + 7) & ~7, tw1 = tw - kBlock;  // pitch of the tile's slot table / of its ring-1 part
      const int jcc[3] = {jc[0], jc[1], jc[2]}, jee[3] = {je[0], je[1], je[2]}, j22[3] = {j2[0], j2[1], j2[2]};
      mbar_wait(&empty[s], ph ^ 1);
      const int c0 = t * kBlock;
      const int ncell = min(kBlock, m.n_own - c0);
      double2 *spa = st_pa(s), *sph = st_ph(s), *scg = st_cg(s), *sen = st_en(s);
      if (lane == 0) {
        int *sh = st_misc(s);
        sh[0] = gw; sh[1] = n1; sh[2] = nhf;  // published by the arrive below (release)
        const uint32_t bytes_c = (uint32_t)ncell * 16u;
        const uint32_t bytes_e = (uint32_t)ne * 16u, bytes_ea = (uint32_t)ne * 8u;
        const uint32_t bytes_gs = (uint32_t)tw1 * 2u, nhf4 = (uint32_t)((nhf + 3) & ~3);
        mbar_expect_tx(&full[s], 2u * bytes_c + bytes_e + bytes_ea + (uint32_t)gw * bytes_gs + nhf4 * 20u);
        bulk_g2s(spa, p2 + c0, bytes_c, &full[s]);
        bulk_g2s(spa + kBlock, p2 + (size_t)np + c0, bytes_c, &full[s]);
        if (ne > 0) {
          bulk_g2s(sen, m.enxy + es, bytes_e, &full[s]);
          bulk_g2s(st_ea(s), m.ea + es, bytes_ea, &full[s]);
        }
        if (tw1 > 0)
          for (int k = 0; k < gw; k++) bulk_g2s(st_gs(s) + k * tw1, fm.gslot + gsb + k * tw + kBlock, bytes_gs, &full[s]);
        if (nhf > 0) {
          bulk_g2s(st_hf(s), fm.hf + hfp, nhf4 * 4u, &full[s]);
          bulk_g2s(st_hfd(s), fm.hfd + hfp, nhf4 * 16u, &full[s]);
        }
      }
      auto gather_h1 = [&](int h, int jg) {  // ring 1: state, gradient operator
        cp_async16(sph + h, p2 + jg);
        cp_async16(sph + HP + h, p2 + (size_t)np + jg);
        for (int r = 0; r < rows; r++) cp_async16(scg + r * H1 + h, fm.gc2 + (size_t)r * np + jg);
      };
      auto gather_h2 = [&](int h, int jg) {  // ring 2: state only
        cp_async16(sph + n1 + h, p2 + jg);
        cp_async16(sph + HP + n1 + h, p2 + (size_t)np + jg);
      };
      auto gather_edge = [&](int h, int jg) {
        cp_async16(sen + ne + h, m.enxy + jg);
        cp_async8(st_ea(s) + ne + h, m.ea + jg);
      };
#pragma unroll
      for (int r = 0; r < 3; r++) {
        const int h = lane + 32 * r;
        if (h < n1) gather_h1(h, jcc[r]);
        if (h < n2) gather_h2(h, j22[r]);
        if (h < nhe) gather_edge(h, jee[r]);
      }
      for (int h = lane + 96; h < n1; h += 32) gather_h1(h, __ldg(&fm.hc_idx[hp + h]));
      for (int h = lane + 96; h < n2; h += 32) gather_h2(h, __ldg(&fm.h2_idx[h2p + h]));
      for (int h = lane + 96; h < nhe; h += 32) gather_edge(h, __ldg(&fm.he_idx[ep + h]));
      cp_async_mbar_arrive_noinc(&full[s]);
      if (j + (int)gridDim.x < fm.ntiles) fetch_meta(tile_id(j + gridDim.x));
    }
    return;
  }

  // ================================== consumer warps ==================================
  double dq2[4] = {0.0, 0.0, 0.0, 0.0};
  const int hst0 = 2 * fm.FW * kBlock;  // first double2 of the ring-1 face states inside X
  int it = 0;
  for (int j = blockIdx.x; j < fm.ntiles; j += gridDim.x, it++) {
    const int t = fm.tile_list ? __ldg(&fm.tile_list[j]) : j;
    const int s = it % kStages;
    const uint32_t ph = (it / kStages) & 1;
    const int c0 = t * kBlock;
    const int ncell = min(kBlock, m.n_own - c0);
    const int i = c0 + tid;
    const bool live = tid < ncell;
    // everything only this thread reads: RK data, coefficient rows, stencil slots, face words, face displacements
    double q0[4], fo[4], dl = 0.0, ivol = 1.0;
    double2 cf[kF2cRows], fd[4];
    uint32_t pk[4];
    int gsl[4];  // (the host only selects this kernel for stencils of <= 4 entries)
    const int4 hh0 = __ldg(&fm.hdr[4 * t]), hh1 = __ldg(&fm.hdr[4 * t + 1]), hh2 = __ldg(&fm.hdr[4 * t + 2]);
    const int fbase = hh1.z, fw = hh1.w, gsb = hh2.z;
    {
      const int gwh = hh2.w, twh = (kBlock + hh0.w + 7) & ~7;
#pragma unroll
      for (int r = 0; r < kF2cRows; r++) cf[r] = (live && r < gwh + F0) ? __ldg(&fm.gc2[(size_t)r * np + i]) : make_double2(0.0, 0.0);
#pragma unroll
      for (int k = 0; k < 4; k++) {
        fd[k] = (live && k < fw) ? __ldg(&fm.fdxy[(size_t)k * np + i]) : make_double2(0.0, 0.0);
        pk[k] = (live && k < fw) ? __ldg(&fm.pack2[fbase + k * kBlock + tid]) : 0xFFFEu;
        gsl[k] = (live && k < gwh) ? (int)__ldg(&fm.gslot[gsb + k * twh + tid]) : tid;
      }
    }
    if (live) {
      stage_load<UM, STEADY>(S, i, np, q, f, dtl, q0, fo, dl);
      ivol = m.ivol[i];
    }
    double2 *scg = st_cg(s);
    const double2 *spa = st_pa(s), *sph = st_ph(s), *sen = st_en(s), *shfd = st_hfd(s);
    const double *sea = st_ea(s);
    const uint32_t *shf = st_hf(s);
    const uint16_t *sgs = st_gs(s);
    mbar_wait(&full[s], ph);
    const int gw = st_misc(s)[0], n1 = st_misc(s)[1], nhf = st_misc(s)[2];
    const int tw1 = ((kBlock + n1 + 7) & ~7) - kBlock;
    auto ld_p = [&](int js, double pj[4]) {
      const double2 *a = js < kBlock ? spa + js : sph + (js - kBlock);
      const int pitch = js < kBlock ? kBlock : HP;
      const double2 u = a[0], w = a[pitch];
      pj[0] = u.x; pj[1] = u.y; pj[2] = w.x; pj[3] = w.y;
    };

    // ---- phase 1a: gradients (own cell: coefficients and slots from registers)
    double p0[4] = {0.0, 0.0, 0.0, 0.0}, gx[4] = {0.0, 0.0, 0.0, 0.0}, gy[4] = {0.0, 0.0, 0.0, 0.0};
    if (live) {
      ld_p(tid, p0);
      if (FORM == 0) {
#pragma unroll
        for (int v = 0; v < 4; v++) { gx[v] = cf[0].x * p0[v]; gy[v] = cf[0].y * p0[v]; }
      }
#pragma unroll
      for (int k = 0; k < 4; k++) {
        if (k < gw) {
          double pj[4];
          ld_p(gsl[k], pj);
#pragma unroll
          for (int v = 0; v < 4; v++) {
            const double d = FORM == 0 ? pj[v] : pj[v] - p0[v];
            gx[v] = fma(cf[k + F0].x, d, gx[v]);
            gy[v] = fma(cf[k + F0].y, d, gy[v]);
          }
        }
      }
    }
    for (int h = tid; h < n1; h += kBlock) {  // ring 1: in place over the column's coefficients (column-private)
      double ph0[4], ax[4], ay[4];
      ld_p(kBlock + h, ph0);
      if (FORM == 0) {
        const double2 cc = scg[h];
#pragma unroll
        for (int v = 0; v < 4; v++) { ax[v] = cc.x * ph0[v]; ay[v] = cc.y * ph0[v]; }
      } else {
#pragma unroll
        for (int v = 0; v < 4; v++) { ax[v] = 0.0; ay[v] = 0.0; }
      }
      for (int k = 0; k < gw; k++) {
        const int js = sgs[k * tw1 + h];
        const double2 c2 = scg[(k + F0) * H1 + h];
        double pj[4];
        ld_p(js, pj);
#pragma unroll
        for (int v = 0; v < 4; v++) {
          const double d = FORM == 0 ? pj[v] : pj[v] - ph0[v];
          ax[v] = fma(c2.x, d, ax[v]);
          ay[v] = fma(c2.y, d, ay[v]);
        }
      }
      scg[h] = make_double2(ax[0], ax[1]);
      scg[H1 + h] = make_double2(ax[2], ax[3]);
      scg[2 * H1 + h] = make_double2(ay[0], ay[1]);
      scg[3 * H1 + h] = make_double2(ay[2], ay[3]);
    }
    // every thread is past phase 2 of the previous tile (X may be rewritten) and the ring-1 gradients are in place
    asm volatile("bar.sync 1, %0;" ::"n"(kBlock) : "memory");

    // ---- phase 1b: publish the reconstructed face states
    if (live) {
#pragma unroll
      for (int k = 0; k < 4; k++) {
        if (k < fw && (pk[k] & 0xFFFFu) != 0xFFFEu) {
          const double dx = fd[k].x, dy = fd[k].y;
          sx[(2 * k) * kBlock + tid] = make_double2(recon_k0(p0[0], gx[0], gy[0], dx, dy), recon_k0(p0[1], gx[1], gy[1], dx, dy));
          sx[(2 * k + 1) * kBlock + tid] = make_double2(recon_k0(p0[2], gx[2], gy[2], dx, dy), recon_k0(p0[3], gx[3], gy[3], dx, dy));
        }
      }
    }
    for (int e = tid; e < nhf; e += kBlock) {
      const int h = shf[e] & 0xFFFFu;
      const double2 dd = shfd[e];
      const double2 a = sph[h], b = sph[HP + h];
      const double2 ga = scg[h], gb = scg[H1 + h], gc = scg[2 * H1 + h], gd = scg[3 * H1 + h];
      sx[hst0 + 2 * e] = make_double2(recon_k0(a.x, ga.x, gc.x, dd.x, dd.y), recon_k0(a.y, ga.y, gc.y, dd.x, dd.y));
      sx[hst0 + 2 * e + 1] = make_double2(recon_k0(b.x, gb.x, gd.x, dd.x, dd.y), recon_k0(b.y, gb.y, gd.y, dd.x, dd.y));
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kBlock) : "memory");

    // ---- phase 2: faces
    double acc[4] = {0.0, 0.0, 0.0, 0.0}, wsacc = 0.0;
    if (live) {
      auto face = [&](const uint32_t w, const int k, const auto bnd_tag) {
        constexpr bool BND = decltype(bnd_tag)::value;
        const int code = w & 0xFFFFu, eslot = (w >> 16) & 0xFFFu, kr = (w >> 28) & 3;
        const bool self_c1 = BND || (w >> 31) == 0;
        const double2 fn = sen[eslot];
        const double af = sea[eslot], nx = fn.x, ny = fn.y;
        const int own0 = (2 * k) * kBlock + tid, own1 = own0 + kBlock;
        double sL[4], sR[4];
        if (!BND) {
          const int nb0 = code < kBlock ? (2 * kr) * kBlock + code : hst0 + 2 * (code - kBlock);
          const int nb1 = code < kBlock ? nb0 + kBlock : nb0 + 1;
          const double2 a = sx[self_c1 ? own0 : nb0], b = sx[self_c1 ? own1 : nb1];
          const double2 c = sx[self_c1 ? nb0 : own0], d = sx[self_c1 ? nb1 : own1];
          sL[0] = a.x; sL[1] = a.y; sL[2] = b.x; sL[3] = b.y;
          sR[0] = c.x; sR[1] = c.y; sR[2] = d.x; sR[3] = d.y;
        } else {
          const double2 a = sx[own0], b2 = sx[own1];
          sL[0] = a.x; sL[1] = a.y; sL[2] = b2.x; sL[3] = b2.y;
          const int b = __ldg(&fm.t_bf[fbase + k * kBlock + tid]);
          const int type = __ldg(&m.bf_type[b]);
          if (type == 2) {  // slip wall: mirror the normal velocity (src/residual.f90:200-204)
            const double un = sL[1] * nx + sL[2] * ny;
            sR[0] = sL[0]; sR[3] = sL[3];
            sR[1] = sL[1] - 2.0 * un * nx;
            sR[2] = sL[2] - 2.0 * un * ny;
          } else {
#pragma unroll
            for (int v = 0; v < 4; v++) sR[v] = __ldg(&bc[v * m.nbf + b]);
          }
        }
        double flux[4], ws;
        roe_flux2(P, sL, sR, nx, ny, flux, ws);
        const double ha = 0.5 * af, sa = self_c1 ? ha : -ha;
#pragma unroll
        for (int v = 0; v < 4; v++) acc[v] += flux[v] * sa;
        wsacc += ws * ha;
      };
      // face pairs (0,1) and (2,3): two interior faces in one basic block so that the two flux evaluations interleave;
      // the loops stay rolled (the face words are picked out of their registers by selects) to keep the code small
      bool has_bnd = false;
#pragma unroll 1
      for (int kk = 0; kk < 2; kk++) {
        const uint32_t w0 = kk ? pk[2] : pk[0], w1 = kk ? pk[3] : pk[1];
        const uint32_t n0 = w0 & 0xFFFFu, n1_ = w1 & 0xFFFFu;
        has_bnd = has_bnd || n0 == 0xFFFFu || n1_ == 0xFFFFu;
        if (n0 < 0xFFFEu && n1_ < 0xFFFEu) {
          face(w0, 2 * kk, std::false_type{});
          face(w1, 2 * kk + 1, std::false_type{});
        } else {
#pragma unroll 1
          for (int h = 0; h < 2; h++) {
            const uint32_t w = h ? w1 : w0;
            if ((w & 0xFFFFu) < 0xFFFEu) face(w, 2 * kk + h, std::false_type{});
          }
        }
      }
      if (has_bnd) {
#pragma unroll 1
        for (int k = 0; k < 4; k++) {
          const uint32_t w = k == 0 ? pk[0] : k == 1 ? pk[1] : k == 2 ? pk[2] : pk[3];
          if ((w & 0xFFFFu) == 0xFFFFu) face(w, k, std::true_type{});
        }
      }
    }
    // (no proxy fence: the regions the bulk copies refill are never written by generic stores -- the gradients go over
    // the ring-1 coefficient block, which is refilled by cp.async, generic proxy; X is not touched by the producer)
    mbar_arrive(&empty[s]);
    if (live) stage_update_pre<UM, STEADY>(P, S, i, np, ivol, m.vol, q0, fo, dl, acc, wsacc, q, f, pout, dtl, nullptr, nullptr, dq2);
  }
  if (S.last) {
    __shared__ double red[4][kBlock / 32];
#pragma unroll
    for (int v = 0; v < 4; v++) {
      double x = dq2[v];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
      if (lane == 0) red[v][warp] = x;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kBlock) : "memory");
    if (tid < 4) {
      double ssum = 0.0;
#pragma unroll
      for (int w = 0; w < kBlock / 32; w++) ssum += red[tid][w];
      partial[blockIdx.x * 4 + tid] = ssum;
    }
  }
}

}  // namespace fvs2d
