// kernels_fused.cuh -- ONE kernel per Runge-Kutta stage: pass A (gradients) folded into the persistent pass-B pipeline
// of kernels.cuh, and -- on several GPUs -- the halo exchange folded into the same kernel as stores to peer memory.
// Applies to second-order upwind reconstruction without limiter (kappa = 0); everything else runs the two-pass path.
//
// Why: the two-pass schedule writes the gradients (64 B/cell) in pass A and reads them back in pass B, and reads the
// primitive state twice -- 160 of the 512-528 algorithmic bytes of a stage ("B_alg - 128 B" is the
// single-pass lower bound).  Here a tile stages, besides what k_flux_pipe stages,
//   * the primitive state of "ring 2" (gradient-stencil members of the tile's cells and of its halo cells = ring 1),
//   * the gradient operator (coefficient rows + a per-tile table of 16-bit stencil slots) of tile + ring 1,
// and the consumer warps run three phases per tile:
//   phase 1a  gradients: own cell -> registers; ring-1 cells -> in place over their coefficients in shared memory
//             (~35 % redundant gradient evaluations, a few per cent of the flux arithmetic)
//   barrier   (the tile's own-cell region X -- state, centroid, coefficient rows -- is dead from here on)
//   phase 1b  publish: own[k][tid] = reconstructed state of this cell at its face k (into X); hst[i] = state of the
//             ring-1 cell at the i-th tile/ring-1 face (list fz_hf, into X behind the own states).  Every face state is
//             evaluated ONCE, by the thread that holds the gradient.
//   barrier
//   phase 2   per face: own[k][tid], and own[k'][ns] (k' = index of the face in the neighbour's list, from the face
//             word) or hst[i]: two 32-byte states instead of the 14 scattered 16-byte loads per face of k_flux_pipe;
//             unit normal and length of the edge; Roe flux; accumulate.  Then the stage update of k_flux_pipe.
// The gradient of a cell is evaluated with explicit fma() in one fixed order wherever it is rebuilt, so every copy is
// bit-identical, both evaluations of an interior face flux agree bitwise (discrete conservation), and the result equals
// the two-pass path's bitwise (k_gradient's `a += c*d` contracts to the same fma).
//
// Stage layout: X [XR][kBlock] double2 (rows 0,1 state; 2 centroid; 3.. coefficient rows) | state of rings 1+2
// [2][HP] | coefficients -> gradients of ring 1 [CG][H1] | centroids of ring 1 [H1] | exy, enxy [2][E] | ea [E] |
// face words [4][kBlock] | 8 ints | stencil slots [W][TW] u16 | tile/ring-1 faces [HF] u32
//
// Reference: src/gradient_ggcb.f90:116-138, src/gradient_ggnb.f90:183-210, src/gradient_lsq.f90:393-401 (phase 1);
// src/residual.f90:66-166, src/flux_invscid.f90:37-136, src/runge_kutta.f90:156-162,225-226,299-313,383-387 (phase 2).
// (Variants measured and removed in round 2, profiles/r2a_fused_ncu.md: gathering state/gradient/centroid per face in
// phase 2 (+2-5 %), every face flux once with a third barrier (+12-16 %), own-cell operands straight into registers to fit
// three CTAs per SM on quadrilateral tiles (+8-14 %: long-scoreboard stalls replace the shared-memory traffic).)
#pragma once
#include "kernels.cuh"

namespace fvs2d {

struct FusedMeta {
  const int4 *hdr;  // 4 x int4 per tile: {es, ne, hc_ptr, n1}, {he_ptr, n_he, fbase, fw}, {h2_ptr, n2, gs_base, gw}, {hf_ptr, n_hf, 0, 0}
  const int *hc_idx, *he_idx, *h2_idx;
  const uint32_t *pack2, *hf;
  const int *t_bf;
  const uint16_t *gslot;  // per tile gw rows of pitch TW = roundup8(kBlock + n1)
  const double2 *gc2;     // gradient coefficients (cx, cy), rows of pitch np: [c0 (FORM 0)], entry 0, entry 1, ...
  int H1, HP, E, TW, W, CG, XR, FW, HF;  // pitches: ring 1, rings 1+2, edges, slot table; max stencil entries; rows of the
                                         // ring-1 block and of X; max faces per cell; max tile/ring-1 faces (multiple of 4)
  int ntiles;
  const int *tile_list;  // null: tiles 0..ntiles-1; else the ntiles tile ids to process (a launch over a subset of the tiles)
};
__host__ __device__ inline size_t fused_stage_bytes(const FusedMeta &f) {
  return (size_t)f.XR * kBlock * 16 + (size_t)2 * f.HP * 16 + (size_t)f.CG * f.H1 * 16 + (size_t)f.H1 * 16 + (size_t)f.E * 40 +
         4 * kBlock * sizeof(uint32_t) + 32 + (((size_t)f.W * f.TW * 2 + 15) & ~(size_t)15) + (size_t)f.HF * 4;
}

// ---- halo exchange inside the stage kernel (several ranks; one process per GPU, peers' arrays mapped with CUDA IPC) ----
// A rank's tiles are launched boundary tiles first (tiles that read a ghost cell or hold a cell a peer needs), then the
// interior ones.  Stage number e (counted over the whole run) on every rank:
//   wait    the producer warp of a CTA whose first tile is a boundary tile spins until every peer's flag says e-1
//           (their stage e-1 results are in this rank's ghost slots, and they are done reading the ghost slots of the
//           buffer this stage overwrites on their side);
//   send    the stage update of a cell a peer needs also stores the new primitive state into that peer's ghost slot
//           (st.global to peer memory over NVLink: the transfer overlaps the tile-by-tile compute, no pack, no NCCL);
//   signal  each CTA, done with its boundary tiles, fences at system scope and bumps a counter; the last one stores e
//           into this rank's flag word at every peer (release, system scope) -- then the CTAs go on with interior tiles.
// The peers' results arrive while the interior tiles run; the next stage finds its flags already set unless a peer is
// more than an interior phase behind.  No kernel of the time loop depends on the host, so the step replays as a CUDA
// graph on several ranks too.  (SURVEY 8e variant 2a: with the deep ghost layers only the state travels, once per stage.)
constexpr int kMaxPeers = 8;
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

struct HaloP2P {
  int n_peers;                     // 0: single rank, nothing of the above happens
  int n_bnd;                       // the first n_bnd entries of tile_list are the boundary tiles
  int n_bnd_ctas;                  // CTAs of this launch that own at least one boundary tile = min(n_bnd, grid)
  int stage;                       // epoch e = clk->epoch0 + 4 * clk->istep + stage + 1
  const StepClock *clk;
  double2 *peer_out[kMaxPeers];    // the peer's array this stage writes (its `pout`), pitch peer_np
  int peer_np[kMaxPeers];
  unsigned *peer_flag[kMaxPeers];  // this rank's flag word in the peer's memory
  const unsigned *my_flag[kMaxPeers];  // the peer's flag word in this rank's memory
  const uint32_t *rs_word;         // per boundary tile (position in tile_list) and thread: first entry << 3 | count
  const int2 *rs_ent;              // (peer slot, ghost index at the peer)
  unsigned *done_ctr;              // CTAs done with their boundary tiles (reset by the last one)
  int *timed_out;                  // set when a wait gave up (the ranks do not run the same sequence of calls): the host
                                   // turns it into an error instead of a hung GPU
};
constexpr unsigned long long kPeerWaitNs = 20ull * 1000000000ull;  // no exchange of a correct run takes 20 s
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// spin until *flag has reached `need` (counters wrap: signed distance); gives up after kPeerWaitNs
__device__ __forceinline__ void wait_flag(const unsigned *flag, unsigned need, int *timed_out) {
  if ((int)(ld_acquire_sys(flag) - need) >= 0) return;
  if (*reinterpret_cast<volatile int *>(timed_out)) return;  // already given up once: do not wait 20 s per stage
  const unsigned long long t0 = global_ns();
  while ((int)(ld_acquire_sys(flag) - need) < 0) {
    __nanosleep(64);
    if (global_ns() - t0 > kPeerWaitNs) { *timed_out = 1; return; }
  }
}
template <int UM, bool STEADY, int FORM, int CTAS, bool P2P>
__global__ void __launch_bounds__(kPipeThreads, CTAS) k_stage_fused(const DevMesh m, const FusedMeta fm, const Phys P, const StageParams S,
                                                                     const double *__restrict__ p, const double *__restrict__ bc,
                                                                     double *__restrict__ q, double *__restrict__ f,
                                                                     double *__restrict__ pout, double *__restrict__ dtl,
                                                                     double *__restrict__ partial, const HaloP2P hx) {
  pdl_entry();
  constexpr int F0 = FORM == 0 ? 1 : 0;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int H1 = fm.H1, HP = fm.HP, EE = fm.E, CG = fm.CG, np = m.np;
  const size_t stage_bytes = fused_stage_bytes(fm);
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
    if ((P2P && hx.n_peers > 0) && (int)blockIdx.x < hx.n_bnd) {
      // boundary tiles come first: before the first ghost is read, every peer must have delivered the previous stage
      const unsigned need = hx.clk->epoch0 + 4u * (unsigned)hx.clk->istep + (unsigned)hx.stage;
      if (lane < hx.n_peers) wait_flag(hx.my_flag[lane], need, hx.timed_out);
      __syncwarp();
    }
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
  bool bnd_open = (P2P && hx.n_peers > 0) && (int)blockIdx.x < hx.n_bnd;  // this CTA still owes its "boundary tiles done"
  auto signal_peers = [&]() {
    __threadfence_system();  // this thread's stores to peer memory are visible system-wide ...
    asm volatile("bar.sync 1, %0;" ::"n"(kBlock) : "memory");  // ... for all consumer threads of the CTA
    if (tid == 0) {
      const unsigned e = hx.clk->epoch0 + 4u * (unsigned)hx.clk->istep + (unsigned)hx.stage + 1u;
      if (atomicAdd(hx.done_ctr, 1u) + 1u == (unsigned)hx.n_bnd_ctas) {  // last CTA: every boundary tile of the rank is done
        *hx.done_ctr = 0u;
        __threadfence_system();
        for (int k = 0; k < hx.n_peers; k++) st_release_sys(hx.peer_flag[k], e);
      }
    }
  };
  int it = 0;
  // the id of a tile is loaded one iteration ahead: the addresses of the cell's Runge-Kutta data depend on it, and an L2
  // round trip at the top of every tile was the largest single long-scoreboard stall of the kernel
  int t_next = ((int)blockIdx.x < fm.ntiles && fm.tile_list) ? __ldg(&fm.tile_list[blockIdx.x]) : (int)blockIdx.x;
  for (int j = blockIdx.x; j < fm.ntiles; j += gridDim.x, it++) {
    const int t = t_next;
    if (j + (int)gridDim.x < fm.ntiles) t_next = fm.tile_list ? __ldg(&fm.tile_list[j + gridDim.x]) : j + (int)gridDim.x;
    const int s = it % kStages;
    const uint32_t ph = (it / kStages) & 1;
    const int c0 = t * kBlock;
    const int ncell = min(kBlock, m.n_own - c0);
    const int i = c0 + tid;
    const bool live = tid < ncell;
    double q0[4], fo[4], dl = 0.0, ivol = 1.0;
    uint32_t rsw = 0;
    if (live) {
      stage_load<UM, STEADY>(S, i, np, q, f, dtl, q0, fo, dl);
      ivol = m.ivol[i];
      if ((P2P && hx.n_peers > 0) && j < hx.n_bnd) rsw = __ldg(&hx.rs_word[(size_t)j * kBlock + tid]);
    }
    if ((P2P && hx.n_peers > 0) && bnd_open && j >= hx.n_bnd) { bnd_open = false; signal_peers(); }  // first interior tile of this CTA
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
    if (live) {
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
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores before the refilling bulk copies
    mbar_arrive(&empty[s]);
    if (live) {
      double2 pr[2];
      stage_update_pre<UM, STEADY>(P, S, i, np, ivol, m.vol, q0, fo, dl, acc, wsacc, q, f, pout, dtl, nullptr, nullptr, dq2, pr);
      for (uint32_t c = 0, e0 = rsw >> 3; c < (rsw & 7u); c++) {  // cells a peer needs: straight into its ghost slot
        const int2 en = __ldg(&hx.rs_ent[e0 + c]);
        double2 *dst = hx.peer_out[en.x] + en.y;
        dst[0] = pr[0];
        dst[hx.peer_np[en.x]] = pr[1];
      }
    }
  }
  if ((P2P && hx.n_peers > 0) && bnd_open) signal_peers();  // this CTA had boundary tiles only
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
