// kernels_coop.cuh -- whole time steps in ONE cooperative kernel, for meshes small enough that every 128-cell tile has its
// own resident CTA (C1: 57 tiles, C2 NACA: 512 tiles; up to 4 x 148 = 592 on a B200).
//
// Why: such meshes live in L2 (NACA: 65 536 cells x ~700 B = 46 MB of the 126 MB) and a time step is ~10 dependent
// kernels of a few microseconds each -- round 1 measured 29 us of kernel time inside an 88 us step even as a CUDA graph.
// Here the kernel sequence of runge_kutta.f90's step loop becomes phases of one launch separated by grid-wide barriers:
//     per step:  4 x [ boundary states | gradient (+ limiter) | barrier | face fluxes + stage update | barrier ]
//                [ vortex errors | barrier ] finish (CTA 0, overlapped with the next step's first phase)
// and fvs2d_gpu_time_integration(t1, nsub) is a single launch for all nsub steps: no launch latency, no host round trip.
// The phases call the same device functions as k_gradient / k_flux_rk / k_bc_state / k_vortex_err / k_finish_step, with
// the same per-CTA reduction order, so results are bitwise those of the two-pass path.
//
// Memory visibility across the phases: the barrier is the cooperative-groups pattern (bar.sync; one thread: fence, arrive,
// spin, fence; bar.sync).  Arrays written inside the kernel reach the phase functions through plain (non-const,
// non-restrict) kernel parameters, so no load of them is compiled to the non-coherent path (checked in the SASS:
// LDG.E.CONSTANT only for mesh tables).
// Reference: src/runge_kutta.f90:120-418 (the step loops), :424-437, :169-184; src/residual.f90:23-177.
#pragma once
#include "kernels.cuh"

namespace fvs2d {

struct CoopArgs {
  double *pa, *pb, *g, *phi, *bc, *q, *f, *dtl;   // state buffers (pa / pb swap every stage inside the kernel)
  double *partial, *vpartial;                     // per-CTA partial sums of the step norms / vortex errors
  int *vbest, *vbest_loc;
  double *logbuf;
  int *logid;
  StepClock *clk;
  unsigned long long *bar;                        // grid barrier counter, zeroed by the host before the launch
  StageParams S[4];
  int nsteps, limiter_type, log_stride, vort, bc_time_dep, bc_done;
};

__device__ __forceinline__ void grid_barrier(unsigned long long *ctr, unsigned long long &target, const unsigned nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += nblocks;
    __threadfence();
    atomicAdd(ctr, 1ull);
    while (*reinterpret_cast<volatile unsigned long long *>(ctr) < target) {}
    __threadfence();
  }
  __syncthreads();
}

// sum of NV per-thread values over the CTA in the order of block_sum_store (warp shuffles, then warps in order)
template <int NV>
__device__ __forceinline__ void cta_sum_store(double val[NV], double *__restrict__ out) {
  __shared__ double sm[NV][kBlock / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int v = 0; v < NV; v++) {
    double x = val[v];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if (lane == 0) sm[v][wid] = x;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kBlock / 32; w++) s += sm[threadIdx.x][w];
    out[threadIdx.x] = s;
  }
  __syncthreads();
}

template <int FORM, bool LIM, int UM, bool STEADY, int RC>
__global__ void __launch_bounds__(kBlock, 4) k_step_coop(const DevMesh m, const Phys P, CoopArgs A) {
  const unsigned nb = gridDim.x;
  const int tid = threadIdx.x, i = blockIdx.x * kBlock + tid;  // one tile per CTA
  const bool live = i < m.n_own;
  unsigned long long target = 0;
  double *pa = A.pa, *pb = A.pb;
  const double t1 = A.clk->t1, dt = A.clk->dt;
  const int istep0 = A.clk->istep;
  for (int step = 0; step < A.nsteps; step++) {
    const double told = t1 + (double)(istep0 + step) * dt;
    double dq2[4] = {0.0, 0.0, 0.0, 0.0};
    for (int rk = 0; rk < 4; rk++) {
      // boundary states for this stage's time (freestream / MMS states are computed once)
      if (A.bc_time_dep || (!A.bc_done && step == 0 && rk == 0))
        for (int b = blockIdx.x * kBlock + tid; b < m.nbf; b += nb * kBlock) bc_state_one(m, P, told + A.clk->off[rk], b, A.bc);
      if (RC != RC_FIRST && live) gradient_cell<FORM, LIM>(m, A.limiter_type, pa, A.g, A.phi, i);
      grid_barrier(A.bar, target, nb);
      if (live) flux_rk_cell<UM, STEADY, RC>(m, P, A.S[rk], pa, A.g, A.phi, A.bc, A.q, A.f, pb, A.dtl, nullptr, nullptr, i, dq2);
      if (rk == 3) cta_sum_store<4>(dq2, A.partial + 4 * (size_t)blockIdx.x);
      grid_barrier(A.bar, target, nb);
      double *t = pa; pa = pb; pb = t;
    }
    if (A.vort) {
      // src/mms.f90:315-361 over this CTA's cells, as k_vortex_err does per CTA (one tile per CTA here)
      const double time = told + A.clk->off_end;
      double d[4] = {0, 0, 0, 0};
      bool in = false;
      int oid = 0x7fffffff;
      if (live) {
        in = m.is_intr[i] != 0;
        const double2 cc = m.xy[i];
        const int np = m.np;
        const double a0 = A.q[i], a1 = A.q[np + i], a2 = A.q[2 * np + i], a3 = A.q[3 * np + i];
        double pv[4];
        vortex_exact(P, time, cc.x, cc.y, pv);
        const double ex0 = pv[0], ex1 = pv[0] * pv[1], ex2 = pv[0] * pv[2];
        const double ex3 = pv[3] / (P.gamma - 1.0) + 0.5 * pv[0] * (pv[1] * pv[1] + pv[2] * pv[2]);
        d[0] = fabs(a0 - ex0); d[1] = fabs(a1 - ex1); d[2] = fabs(a2 - ex2); d[3] = fabs(a3 - ex3);
        oid = m.orig_id[i];
      }
      __shared__ double smx[4][kBlock / 32], s1[4][kBlock / 32], s2[4][kBlock / 32], sb[kBlock / 32];
      __shared__ int sid[kBlock / 32], sil[kBlock / 32];
      const int lane = tid & 31, wid = tid >> 5;
      double bv = in ? d[0] : -1.0;
      int bi = in ? oid : 0x7fffffff, bl = in ? i : 0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_down_sync(0xffffffffu, bv, o);
        const int oi = __shfl_down_sync(0xffffffffu, bi, o), ol = __shfl_down_sync(0xffffffffu, bl, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; bl = ol; }
      }
#pragma unroll
      for (int v = 0; v < 4; v++) {
        double mx = in ? d[v] : 0.0, a = in ? d[v] : 0.0, b = in ? d[v] * d[v] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o));
          a += __shfl_down_sync(0xffffffffu, a, o);
          b += __shfl_down_sync(0xffffffffu, b, o);
        }
        if (lane == 0) { smx[v][wid] = mx; s1[v][wid] = a; s2[v][wid] = b; }
      }
      if (lane == 0) { sb[wid] = bv; sid[wid] = bi; sil[wid] = bl; }
      __syncthreads();
      if (tid < 4) {
        double mx = 0, a = 0, b = 0;
        for (int w = 0; w < kBlock / 32; w++) { mx = fmax(mx, smx[tid][w]); a += s1[tid][w]; b += s2[tid][w]; }
        A.vpartial[(size_t)blockIdx.x * 13 + tid] = mx;
        A.vpartial[(size_t)blockIdx.x * 13 + 4 + tid] = a;
        A.vpartial[(size_t)blockIdx.x * 13 + 8 + tid] = b;
      }
      if (tid == 0) {
        double bb = sb[0]; int ii = sid[0], ll = sil[0];
        for (int w = 1; w < kBlock / 32; w++)
          if (sb[w] > bb || (sb[w] == bb && sid[w] < ii)) { bb = sb[w]; ii = sid[w]; ll = sil[w]; }
        A.vpartial[(size_t)blockIdx.x * 13 + 12] = bb;
        A.vbest[blockIdx.x] = ii;
        A.vbest_loc[blockIdx.x] = ll;
      }
      grid_barrier(A.bar, target, nb);
    }
    // finish (CTA 0): the step's row of the device log, same lane-strided sums + shuffle trees as k_finish_step.  The
    // other CTAs go on with the next step; `partial` / `vpartial` are rewritten only after several more barriers.
    if (blockIdx.x == 0) {
      const int warp = tid >> 5, lane = tid & 31;
      double *row = A.logbuf + (size_t)A.log_stride * (istep0 + step);
      for (int qn = warp; qn < 17; qn += kBlock / 32) {
        if (qn < 4) {
          double s = 0.0;
          for (unsigned b = lane; b < nb; b += 32) s += A.partial[(size_t)b * 4 + qn];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
          if (lane == 0) row[qn] = s;
        } else if (!A.vort) {
          continue;
        } else if (qn < 16) {
          const int v = qn - 4;
          double s = 0.0;
          for (unsigned b = lane; b < nb; b += 32) {
            const double x = A.vpartial[(size_t)b * 13 + v];
            s = v < 4 ? fmax(s, x) : s + x;
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const double x = __shfl_down_sync(0xffffffffu, s, o);
            s = v < 4 ? fmax(s, x) : s + x;
          }
          if (lane == 0) row[4 + v] = s;
        } else {
          double bv = -1.0;
          int bi = 0x7fffffff, bl = 0;
          for (unsigned b = lane; b < nb; b += 32) {
            const double x = A.vpartial[(size_t)b * 13 + 12];
            const int id = A.vbest[b];
            if (x > bv || (x == bv && id < bi)) { bv = x; bi = id; bl = A.vbest_loc[b]; }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const double x = __shfl_down_sync(0xffffffffu, bv, o);
            const int id = __shfl_down_sync(0xffffffffu, bi, o), il = __shfl_down_sync(0xffffffffu, bl, o);
            if (x > bv || (x == bv && id < bi)) { bv = x; bi = id; bl = il; }
          }
          if (lane == 0) {
            const double2 c = m.xy[bl];
            row[16] = bv; row[17] = c.x; row[18] = c.y; A.logid[istep0 + step] = bi;
          }
        }
      }
    }
  }
  if (blockIdx.x == 0 && tid == 0) A.clk->istep = istep0 + A.nsteps;  // (every CTA read istep before its first barrier)
}

}  // namespace fvs2d
