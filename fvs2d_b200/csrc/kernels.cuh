// kernels.cuh -- sm_100a kernels of the fvs2d hot path (fp64, HBM-bound, no tensor cores).
//
//   k_gradient      pass A: cell-parallel gradient (+ limiter) from the primitive state        [reference K1-K3]
//   k_flux_pipe     pass B: face-flux gather + residual + RK stage update; persistent, warp-specialised
//                   TMA / cp.async shared-memory pipeline (production path)                    [reference K4-K9]
//   k_flux_rk       pass B, direct global gathers (fallback for meshes whose tiles do not fit the pipeline)
//   (kernels_fused.cuh: k_stage_fused / k_stage_fused2 / k_stage_fused2c -- pass A folded into the pass-B pipeline,
//    one kernel per stage, option "fuse")
//   k_bc_state      Dirichlet / freestream ghost states of the boundary faces for one stage time
//   k_prim          conserved -> primitive
//   k_vortex_err    isentropic-vortex error norms                                             [reference K10]
//   k_finish_step   fixed-order final reductions of the per-CTA partials (no fp atomics anywhere) and the
//                   device-side step counter (lets a time step be replayed as a CUDA graph)
//   k_pack, k_scatter_in, k_gather_out   halo packing and AoS <-> device-layout copies
//   k_cell2node     cell -> node inverse-distance interpolation of the primitive state (output path, per save)
//   k_wall_values   wall pressure / normal velocity at the boundary-edge centres (output path, per save)
//
// Data layout: struct-of-arrays with pitch `np` (cells padded to a multiple of 32).  The arrays that other
// cells gather -- primitive state p (4 vars), gradients g (8 vars: gx0-3, gy0-3), centroids xy, edge centres
// exy and normals enxy -- are PAIR-interleaved: variables 2k and 2k+1 of cell i form the double2
// a2[k*np + i], so every gather is a 16-byte access (LDG.128 / LDS.128 / L1-bypassing cp.async.cg) and
// own-cell runs stay contiguous for the TMA bulk copies.  q, f (touched only by the owning thread) are
// plain SoA a[v*np + i].  One thread per cell; per-cell lists are sliced ELL (see layout.hpp) so list
// reads of a warp are coalesced.  Face fluxes are evaluated in the edge's own orientation (c1 -> c2) by both
// cells, so the two evaluations are bit-identical and the scheme stays discretely conservative.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

namespace fvs2d {

#ifndef FVS2D_TILE
#define FVS2D_TILE 128
#endif
#ifndef FVS2D_PIPE_CTAS
#define FVS2D_PIPE_CTAS 3
#endif
constexpr int kBlock = FVS2D_TILE;  // threads per CTA for the cell-parallel kernels = cells per tile
constexpr int kPadNbr = INT32_MIN;  // == kFacePad

enum UpdateMode { UM_RESID = 0, UM_RK = 1, UM_SSPRK = 2 };
// reconstruction variants: 0 first order (phi=0, no gradients); 1 kappa=0 and phi==1;
// 2 kappa=0 with a limiter array; 3 general kappa with a phi array
enum ReconMode { RC_FIRST = 0, RC_K0 = 1, RC_K0_PHI = 2, RC_GENERAL = 3 };

struct DevMesh {
  int n_own, n_loc, np;  // np: SoA pitch
  const int *f_off, *f_nbr, *f_edge;
  const double2 *exy, *enxy;  // edge centre (x,y), unit normal (nx,ny) c1 -> c2
  const double *ea;           // edge length
  const double2 *xy;          // cell centroid
  const double *vol;
  const double *ivol;         // 1/vol, rounded once on the host: -R/vol is a multiplication in the stage update
  const int *g_off, *g_idx;
  const double *g_cx, *g_cy, *c0x, *c0y;
  const int *bf_type, *bf_edge;
  const unsigned char *is_intr;
  const int *orig_id;
  int nbf;
};

struct Phys {
  double gamma, kappa, cfl;
  double gm1, gog;  // gamma-1, gamma/(gamma-1)
  double pinf[4];
  double vpos[2], vkap, vinf[4];
  double mms[4][4];
  int lvortex, limiter;
  int pow2n, pad;   // n when 2/(gamma-1) is an integer (gamma = 1.4 -> 5, 5/3 -> 3), else 0
};

// Time of the current step / stage, kept in device memory so that the kernel sequence of one time step has no
// host-side parameter that changes from step to step (it can be replayed as a CUDA graph):
// told = t1 + istep*dt (src/runge_kutta.f90:135), stage time = told + off[stage], end-of-step time = told + off_end
struct StepClock {
  double t1, dt, off[4], off_end;
  int istep;
  unsigned epoch0;  // Runge-Kutta stages completed before this call (numbering of the in-kernel halo exchange)
};
__device__ __forceinline__ double clock_told(const StepClock *c) { return c->t1 + (double)c->istep * c->dt; }

struct StageParams {
  int stage;        // 0..3
  int last;         // stage == nstages-1
  double h;         // h_rk(stage)   (unsteady)  |  cst (steady: multiplies dt_local)
  double c;         // rk_coef(stage)
  double dt;        // global dt
};

// Programmatic dependent launch (the kernels of a time step form one dependent chain; on small meshes the ~2.5 us between
// two kernels of a graph are a third of the step): every kernel of the chain waits here for its predecessor's results
// before its first global access, and at once lets its own successor be scheduled, so that the successor's launch latency
// and prologue hide behind this kernel.  Without the launch attribute both instructions do nothing.
__device__ __forceinline__ void pdl_entry() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// variable v of cell i in a pair-interleaved array (as doubles)
__host__ __device__ __forceinline__ size_t pidx(int v, int np, int i) { return ((size_t)(v >> 1) * np + i) * 2 + (v & 1); }
__device__ __forceinline__ void load4(const double2 *__restrict__ a2, int np, int i, double out[4]) {
  const double2 a = a2[i], b = a2[np + i];
  out[0] = a.x; out[1] = a.y; out[2] = b.x; out[3] = b.y;
}

__device__ __forceinline__ void mms_exact(const Phys &P, double x, double y, double pv[4]) {
  // src/mms.f90:137-146 (solution only; nx+ny==0 branch of manufactured_sol :203)
#pragma unroll
  for (int v = 0; v < 4; v++) pv[v] = P.mms[v][0] + P.mms[v][1] * sin(P.mms[v][2] * x + P.mms[v][3] * y);
}

// Branch-free fp64 reciprocal and reciprocal square root for normal, positive arguments (densities, w+z, a^2):
// hardware seed (MUFU.RCP64H / MUFU.RSQ64H, 20 mantissa bits) + ONE third-order step -> rounding-level error
// (residual e ~ 2^-19, truncation ~ e^3 < 2^-56), no slow path.  The IEEE division / sqrt sequences cost twice
// the instructions plus a branch, and two Newton steps are a dependent chain twice as long as this.
__device__ __forceinline__ double fast_rcp(const double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x, y, 1.0);   // 1/x = y / (1-e) = y (1 + e + e^2 + ...)
  return fma(y, fma(e, e, e), y);
}
__device__ __forceinline__ double fast_rsqrt(const double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x * y, y, 1.0);  // x^-1/2 = y (1-e)^-1/2 = y (1 + e/2 + 3e^2/8 + ...)
  return fma(y * e, fma(0.375, e, 0.5), y);
}
// sqrt(x) and 1/x from one rsqrt: s = x*r corrected by one Newton step, 1/x = r*r
__device__ __forceinline__ void fast_sqrt_rcp(const double x, double &s, double &inv) {
  const double r = fast_rsqrt(x);
  double t = x * r;
  t = fma(fma(-t, t, x), 0.5 * r, t);
  s = t;
  inv = r * r;
}

// rho = temp^(1/(gamma-1)) (src/mms.f90:251).  For the usual gas constants 2/(gamma-1) is an integer n, so the power
// is sqrt(temp)^n: one rsqrt chain and a few multiplications instead of log + exp (the exponent the reference
// forms, 1/(gamma-1) in fp64, differs from n/2 by an ulp: a relative change of 1e-16*|log temp|).
__device__ __forceinline__ double pow_inv_gm1(const Phys &P, const double temp) {
  if (P.pow2n > 0) {
    double s, inv;
    fast_sqrt_rcp(temp, s, inv);
    double r = (P.pow2n & 1) ? s : 1.0, b = temp;
    for (int n = P.pow2n >> 1; n; n >>= 1) {
      if (n & 1) r *= b;
      b *= b;
    }
    return r;
  }
  return exp(log(temp) / P.gm1);
}

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void vortex_exact(const Phys &P, double t, double x, double y, double pv[4]) {
  // src/mms.f90:219-265.  exp(1-r^2) = exp((1-r^2)/2)^2; rho = temp^(1/(gamma-1)) once (pow_inv_gm1),
  // p = rho^gamma = rho * temp (two pow calls cost three times as much)
  const double pi = 3.141592653589793238462643383279502884;
  const double rho_inf = P.vinf[0], u_inf = P.vinf[1], v_inf = P.vinf[2], p_inf = P.vinf[3];
  const double T_inf = p_inf / rho_inf;
  const double xc = P.vpos[0] + u_inf * t, yc = P.vpos[1] + v_inf * t;
  const double dx = x - xc, dy = y - yc;
  const double r2 = dx * dx + dy * dy;
  const double kk = P.vkap / (2.0 * pi);
  const double e1 = exp(0.5 * (1.0 - r2));
  pv[1] = u_inf - kk * dy * e1;
  pv[2] = v_inf + kk * dx * e1;
  const double temp = T_inf - kk * kk * (P.gamma - 1.0) / (2.0 * P.gamma) * (e1 * e1);
  const double rho = pow_inv_gm1(P, temp);
  pv[0] = rho;
  pv[3] = rho * temp;  // rho^gamma = rho * rho^(gamma-1) = rho * temp
}

// Roe flux with Harten's entropy fix, primitive inputs (src/flux_invscid.f90:37-136).
// aL, aR only ever appear squared in the reference (HL = aL*aL/(gamma-1)+kL), so c2 = gamma*p/rho is
// used directly; divisions are reciprocals shared between quotients; gog = gamma/(gamma-1).
// Returns TWICE the flux and TWICE ws_max: the reference's factors 0.5 (:118-133) are folded by the callers into the
// face area (0.5*a is exact, so the products are bitwise those of 0.5*(...)*a).
__device__ __forceinline__ void roe_flux2(const Phys &P, const double L[4], const double R[4], const double nx,
                                         const double ny, double flux[4], double &ws_max) {
  const double gm1 = P.gm1, gog = P.gog;
  const double tx = -ny, ty = nx;
  const double rhoL = L[0], uL = L[1], vL = L[2], pL = L[3];
  const double rhoR = R[0], uR = R[1], vR = R[2], pR = R[3];
  // Roe averages with w = sqrt(rhoL), z = sqrt(rhoR): RT = z/w, 1/(1+RT) = w/(w+z), rho~ = w*z.  The two rsqrt chains
  // are independent (the reference's sqrt(rhoR/rhoL) -> 1/(1+RT) chain is serial) and give 1/rho = rs^2 for free.
  const double rsL = fast_rsqrt(rhoL), rsR = fast_rsqrt(rhoR);
  const double w = rhoL * rsL, z = rhoR * rsR;
  const double irL = rsL * rsL, irR = rsR * rsR;
  const double unL = uL * nx + vL * ny, unR = uR * nx + vR * ny;
  const double utL = uL * tx + vL * ty, utR = uR * tx + vR * ty;
  const double kL = 0.5 * (uL * uL + vL * vL), kR = 0.5 * (uR * uR + vR * vR);
  const double HL = gog * pL * irL + kL;
  const double HR = gog * pR * irR + kR;
  const double rho = w * z;
  const double iw = fast_rcp(w + z);
  const double u = (w * uL + z * uR) * iw;
  const double v = (w * vL + z * vR) * iw;
  const double H = (w * HL + z * HR) * iw;
  const double tke = 0.5 * (u * u + v * v);
  const double a2 = gm1 * (H - tke);
  double a, ia2;
  fast_sqrt_rcp(a2, a, ia2);
  const double un = u * nx + v * ny, ut = u * tx + v * ty;
  const double drho = rhoR - rhoL, dp = pR - pL, dun = unR - unL, dut = utR - utL;
  const double rad = rho * a * dun, hia2 = 0.5 * ia2;
  const double l1 = (dp - rad) * hia2;
  const double l2 = rho * dut;
  const double l3 = drho - dp * ia2;
  const double l4 = (dp + rad) * hia2;
  double w1 = fabs(un - a), w2 = fabs(un), w4 = fabs(un + a);
  const double dws = 1.0 / 5.0;
  // Harten's entropy fix (src/flux_invscid.f90:97-101): 0.5*(w*w/dws + dws) = fma(w*w, 2.5, 0.5*dws) -- scaling by
  // a power of two commutes with the rounding, so this is bitwise 0.5*fma(w*w, 5, dws)
  w1 = w1 < dws ? fma(w1 * w1, 2.5, 0.5 * dws) : w1;
  w4 = w4 < dws ? fma(w4 * w4, 2.5, 0.5 * dws) : w4;
  const double s1 = w1 * l1, s2 = w2 * l2, s3 = w2 * l3, s4 = w4 * l4;
  // diss_i = sum_j ws_j LdU_j R_ij, j = 1..4 in order (src/flux_invscid.f90:111-116)
  const double anx = a * nx, any = a * ny, una = un * a;
  const double d0 = s1 + s3 + s4;
  const double d1 = s1 * (u - anx) + s2 * tx + s3 * u + s4 * (u + anx);
  const double d2 = s1 * (v - any) + s2 * ty + s3 * v + s4 * (v + any);
  const double d3 = s1 * (H - una) + s2 * ut + s3 * tke + s4 * (H + una);
  const double mL = rhoL * unL, mR = rhoR * unR;
  flux[0] = mL + mR - d0;
  flux[1] = mL * uL + pL * nx + (mR * uR + pR * nx) - d1;
  flux[2] = mL * vL + pL * ny + (mR * vR + pR * ny) - d2;
  flux[3] = mL * HL + mR * HR - d3;
  ws_max = fabs(un) + a;
}

// second-order upwind state (kappa = 0, no limiter): p + (x_f - x_c) . grad p as two fused multiply-adds
__device__ __forceinline__ double recon_k0(const double p, const double gx, const double gy, const double dx, const double dy) {
  return fma(dy, gy, fma(dx, gx, p));
}

// limiter function (src/gradient_limiter.f90:103-134); eps2 is precomputed per cell
__device__ __forceinline__ double limiter_fn(int type, double a, double b, double eps2) {
  if (type == 1) return ((a * a + eps2) + 2.0 * b * a) / (a * a + 2.0 * (b * b) + a * b + eps2);
  if (type == 2) return fmin(1.0, a / b);
  const double l = ((b * b + eps2) * a + (a * a + eps2) * b) / (a * a + b * b + 2.0 * eps2);
  return l / (b + eps2);
}

// ------------------------------------------------------------------------------------------------
// K1: conserved -> primitive (src/data_solution.f90:72-86)
__global__ void __launch_bounds__(256) k_prim(int n, int np, double gamma, const double *__restrict__ q, double *__restrict__ p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double r = q[i], ru = q[np + i], rv = q[2 * np + i], re = q[3 * np + i];
  const double ir = fast_rcp(r);  // the same reciprocal as the stage update, so a restart reproduces the run bitwise
  const double u = ru * ir, v = rv * ir;
  double2 *p2 = reinterpret_cast<double2 *>(p);
  p2[i] = make_double2(r, u);
  p2[np + i] = make_double2(v, (gamma - 1.0) * (re - 0.5 * r * (u * u + v * v)));
}

// ------------------------------------------------------------------------------------------------
// pass A: gradient of the primitive variables (+ limiter) for the owned cells
//   FORM 0: grad = c0*p_i + sum c_k p_k   (GGCB src/gradient_ggcb.f90:116-138, GGNB src/gradient_ggnb.f90:183-210)
//   FORM 1: grad = sum c_k (p_k - p_i)    (LSQ src/gradient_lsq.f90:393-401)
//   LIM: phi_i = min over vars and faces (src/gradient_limiter.f90:47-91); min/max over the stencil
template <int FORM, bool LIM>
__device__ __forceinline__ void gradient_cell(const DevMesh &m, const int limiter_type, const double *__restrict__ p,
                                              double *__restrict__ g, double *__restrict__ phi, const int i) {
  const int np = m.np, lane = i & 31, sl = i >> 5;
  const int off = __ldg(&m.g_off[sl]);
  const int w = (__ldg(&m.g_off[sl + 1]) - off) >> 5;
  const double2 *p2 = reinterpret_cast<const double2 *>(p);
  double p0[4], ax[4], ay[4], pmin[4], pmax[4];
  load4(p2, np, i, p0);
  if (FORM == 0) {
    const double c0x = m.c0x[i], c0y = m.c0y[i];
#pragma unroll
    for (int v = 0; v < 4; v++) { ax[v] = c0x * p0[v]; ay[v] = c0y * p0[v]; }
  } else {
#pragma unroll
    for (int v = 0; v < 4; v++) { ax[v] = 0.0; ay[v] = 0.0; }
  }
  if (LIM) {
#pragma unroll
    for (int v = 0; v < 4; v++) { pmin[v] = p0[v]; pmax[v] = p0[v]; }
  }
  for (int k = 0; k < w; k++) {
    const int e = off + 32 * k + lane;
    const int j = __ldg(&m.g_idx[e]);
    const double cx = __ldg(&m.g_cx[e]), cy = __ldg(&m.g_cy[e]);
    double pjv[4];
    load4(p2, np, j, pjv);  // two 16-byte gathers per stencil member
#pragma unroll
    for (int v = 0; v < 4; v++) {
      const double pj = pjv[v];
      const double d = FORM == 0 ? pj : pj - p0[v];
      ax[v] += cx * d;
      ay[v] += cy * d;
      if (LIM) { pmin[v] = fmin(pmin[v], pj); pmax[v] = fmax(pmax[v], pj); }
    }
  }
  double2 *g2 = reinterpret_cast<double2 *>(g);
  g2[i] = make_double2(ax[0], ax[1]);
  g2[np + i] = make_double2(ax[2], ax[3]);
  g2[2 * np + i] = make_double2(ay[0], ay[1]);
  g2[3 * np + i] = make_double2(ay[2], ay[3]);
  if (LIM) {
    const double pi = 3.141592653589793238462643383279502884;
    const double xc = m.xy[i].x, yc = m.xy[i].y;
    const double h = 2.0 * sqrt(m.vol[i] / pi);
    const double kh = (limiter_type == 1 ? 5.0 : 0.3) * h;
    const double eps2 = kh * kh * kh;
    const int foff = __ldg(&m.f_off[sl]);
    const int fw = (__ldg(&m.f_off[sl + 1]) - foff) >> 5;
    double ph = 1.0;  // min(1, ...) over faces and variables
    for (int k = 0; k < fw; k++) {
      const int e = foff + 32 * k + lane;
      if (__ldg(&m.f_nbr[e]) == kPadNbr) continue;
      const int ed = __ldg(&m.f_edge[e]) >> 1;
      const double2 ec = __ldg(&m.exy[ed]);
      const double dx = ec.x - xc, dy = ec.y - yc;
#pragma unroll
      for (int v = 0; v < 4; v++) {
        const double pf = p0[v] + dx * ax[v] + dy * ay[v];
        const double diff = pf - p0[v];
        // one evaluation with the selected extremum instead of two divergent calls (the same function of the same
        // arguments: bitwise the reference's value); diff == 0 -> 1 (src/gradient_limiter.f90:80-86)
        const double f = limiter_fn(limiter_type, (diff > 0.0 ? pmax[v] : pmin[v]) - p0[v], diff, eps2);
        ph = fmin(ph, diff != 0.0 ? f : 1.0);
      }
    }
    phi[i] = ph;
  }
}

template <int FORM, bool LIM>
__global__ void __launch_bounds__(kBlock) k_gradient(const DevMesh m, const int limiter_type, const double *__restrict__ p,
                                                     double *__restrict__ g, double *__restrict__ phi,
                                                     const int *__restrict__ tile_list) {
  pdl_entry();
  // one CTA = one 128-cell tile; tile_list (or null = all tiles in order) selects the interior / boundary subset
  const int i = (tile_list ? __ldg(&tile_list[blockIdx.x]) : (int)blockIdx.x) * kBlock + threadIdx.x;
  if (i >= m.n_own) return;
  gradient_cell<FORM, LIM>(m, limiter_type, p, g, phi, i);
}

// ------------------------------------------------------------------------------------------------
// ghost states of the boundary faces that do not depend on the interior state
// (src/residual.f90:195-217: freestream -> pvar_inf, dirichlet -> vortex(t) or MMS at the face centre)
__device__ __forceinline__ void bc_state_one(const DevMesh &m, const Phys &P, const double time, const int b, double *__restrict__ bc) {
  const int type = m.bf_type[b];
  double pv[4] = {0, 0, 0, 0};
  if (type == 1) {
#pragma unroll
    for (int v = 0; v < 4; v++) pv[v] = P.pinf[v];
  } else if (type == 4) {
    const int ed = m.bf_edge[b];
    const double2 ec = m.exy[ed];
    if (P.lvortex) vortex_exact(P, time, ec.x, ec.y, pv);
    else mms_exact(P, ec.x, ec.y, pv);
  }
#pragma unroll
  for (int v = 0; v < 4; v++) bc[v * m.nbf + b] = pv[v];
}
__global__ void __launch_bounds__(128) k_bc_state(const DevMesh m, const Phys P, const StepClock *__restrict__ clk, const int stage,
                                                  double *__restrict__ bc /* [4][nbf] */) {
  pdl_entry();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= m.nbf) return;
  bc_state_one(m, P, clock_told(clk) + clk->off[stage], b, bc);
}

// ------------------------------------------------------------------------------------------------
// block-wide sum of NV values per thread -> out[blockIdx.x*NV + v] (fixed order, deterministic)
template <int NV>
__device__ __forceinline__ void block_sum_store(double val[NV], double *__restrict__ out) {
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
    out[blockIdx.x * NV + threadIdx.x] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// pass B building blocks, shared by the direct-gather kernel (k_flux_rk) and the shared-memory tile
// kernel (k_flux_pipe).
//   interior faces  src/residual.f90:66-103     boundary faces  src/residual.f90:111-157
//   -R/vol          src/residual.f90:164-166    local dt        src/runge_kutta.f90:424-437
//   RK update       src/runge_kutta.f90:156-162 (RK), :225-226 (SSPRK), :299-313, :383-387 (steady)
//   norms           src/runge_kutta.f90:169-184 (sum of (q-q0)^2 per CTA -> partial)

// interior face: p0/me/phi0 = this cell's state, reconstruction increment and limiter; pj/ot/phij the
// neighbour's.  The flux is evaluated in the edge's own orientation (L = c1, R = c2).
// RC_K0: me / ot are the reconstructed STATES (recon_k0), for the other modes the increments (x_f - x_c) . grad p.
template <int RC>
__device__ __forceinline__ void interior_states(const Phys &P, const bool self_c1, const double p0[4], const double me[4],
                                                const double phi0, const double pj[4], const double ot[4], const double phij,
                                                double sL[4], double sR[4]) {
  const double kap = P.kappa;
#pragma unroll
  for (int v = 0; v < 4; v++) {
    const double pL = self_c1 ? p0[v] : pj[v], pR = self_c1 ? pj[v] : p0[v];
    if (RC == RC_FIRST) { sL[v] = pL; sR[v] = pR; }
    else {
      const double gL = self_c1 ? me[v] : ot[v], gR = self_c1 ? ot[v] : me[v];
      const double fL = self_c1 ? phi0 : phij, fR = self_c1 ? phij : phi0;
      if (RC == RC_K0) { sL[v] = gL; sR[v] = gR; }
      else if (RC == RC_K0_PHI) { sL[v] = pL + fL * gL; sR[v] = pR + fR * gR; }
      else {
        const double gC = pR - pL;
        sL[v] = pL + fL * (kap / 2.0 * gC + (1.0 - kap) * gL);
        sR[v] = pR + fR * (-kap / 2.0 * gC + (1.0 - kap) * gR);
      }
    }
  }
}
template <int RC>
__device__ __forceinline__ void interior_face(const Phys &P, const bool self_c1, const double p0[4], const double me[4],
                                              const double phi0, const double pj[4], const double ot[4], const double phij,
                                              const double nx, const double ny, const double af, double acc[4], double &wsacc) {
  double sL[4], sR[4];
  interior_states<RC>(P, self_c1, p0, me, phi0, pj, ot, phij, sL, sR);
  double flux[4], ws;
  roe_flux2(P, sL, sR, nx, ny, flux, ws);
  const double ha = 0.5 * af, sa = self_c1 ? ha : -ha;
#pragma unroll
  for (int v = 0; v < 4; v++) acc[v] += flux[v] * sa;
  wsacc += ws * ha;
}

// boundary face: this cell is c1 (src/residual.f90:125-155); bcv = ghost state for freestream/dirichlet
template <int RC>
__device__ __forceinline__ void boundary_states(const int type, const double p0[4], const double me[4], const double phi0,
                                                const double bcv[4], const double nx, const double ny, double sL[4], double sR[4]) {
#pragma unroll
  for (int v = 0; v < 4; v++) sL[v] = (RC == RC_FIRST) ? p0[v] : (RC == RC_K0) ? me[v] : p0[v] + phi0 * me[v];
  if (type == 2) {  // slip wall: mirror the normal velocity
    const double un = sL[1] * nx + sL[2] * ny;
    sR[0] = sL[0]; sR[3] = sL[3];
    sR[1] = sL[1] - 2.0 * un * nx;
    sR[2] = sL[2] - 2.0 * un * ny;
  } else {
#pragma unroll
    for (int v = 0; v < 4; v++) sR[v] = bcv[v];
  }
}
template <int RC>
__device__ __forceinline__ void boundary_face(const Phys &P, const int type, const double p0[4], const double me[4],
                                              const double phi0, const double bcv[4], const double nx, const double ny,
                                              const double af, double acc[4], double &wsacc) {
  double sL[4], sR[4];
  boundary_states<RC>(type, p0, me, phi0, bcv, nx, ny, sL, sR);
  double flux[4], ws;
  roe_flux2(P, sL, sR, nx, ny, flux, ws);
  const double ha = 0.5 * af;
#pragma unroll
  for (int v = 0; v < 4; v++) acc[v] += flux[v] * ha;
  wsacc += ws * ha;
}

// residual -> stage update of cell i with the cell's RK data already in registers (q0 = state at the
// start of the step, fo = accumulated stage residuals, dl = dt_local of stages > 0);
// returns (q_new - q0)^2 in dq2 on the last stage
// ivol = 1/vol (rounded once on the host), vol_arr = the volumes (read only by the steady stage-0 local time step)
template <int UM, bool STEADY>
__device__ __forceinline__ void stage_update_pre(const Phys &P, const StageParams &S, const int i, const int np, const double ivol,
                                                 const double *__restrict__ vol_arr,
                                                 const double q0[4], const double fo[4], const double dl_in, const double acc[4],
                                                 const double wsacc, double *__restrict__ q, double *__restrict__ f,
                                                 double *__restrict__ pout, double *__restrict__ dtl, double *__restrict__ resid_out,
                                                 double *__restrict__ ws_out, double dq2[4], double2 *prim_new = nullptr) {
  double R[4];
  const double niv = -ivol;
#pragma unroll
  for (int v = 0; v < 4; v++) R[v] = acc[v] * niv;
  if (UM == UM_RESID) {
#pragma unroll
    for (int v = 0; v < 4; v++) resid_out[v * np + i] = R[v];
    if (ws_out) ws_out[i] = wsacc;
    return;
  }
  double h = S.h;
  if (STEADY) {
    double dl = dl_in;
    if (S.stage == 0) { dl = P.cfl * vol_arr[i] / (0.5 * wsacc); dtl[i] = dl; }
    h = dl * S.h;
  }
  double qn[4];
  if (UM == UM_RK) {
#pragma unroll
    for (int v = 0; v < 4; v++) {
      const double fn = fo[v] + S.c * R[v];
      qn[v] = S.last ? q0[v] + h * fn : q0[v] + h * R[v];
      if (!S.last) f[v * np + i] = fn;
    }
  } else {
#pragma unroll
    for (int v = 0; v < 4; v++) {
      qn[v] = q0[v] + h * (S.c * R[v] + fo[v]);
      if (!S.last) f[v * np + i] = fo[v] + R[v];
    }
  }
  // primitive state for the next stage (cvar2pvar of the next compute_residual); 1/rho by fast_rcp
  const double ir = fast_rcp(qn[0]);
  const double u = qn[1] * ir, vv = qn[2] * ir;
  double2 *po = reinterpret_cast<double2 *>(pout);
  const double2 pr0 = make_double2(qn[0], u), pr1 = make_double2(vv, (P.gamma - 1.0) * (qn[3] - 0.5 * qn[0] * (u * u + vv * vv)));
  po[i] = pr0;
  po[np + i] = pr1;
  if (prim_new) { prim_new[0] = pr0; prim_new[1] = pr1; }  // for the in-kernel halo exchange (kernels_fused.cuh)
  if (S.last) {
#pragma unroll
    for (int v = 0; v < 4; v++) {
      q[v * np + i] = qn[v];
      const double d = fabs(qn[v] - q0[v]);
      dq2[v] += d * d;
    }
  }
}

template <int UM, bool STEADY>
__device__ __forceinline__ void stage_load(const StageParams &S, const int i, const int np, const double *__restrict__ q,
                                           const double *__restrict__ f, const double *__restrict__ dtl, double q0[4],
                                           double fo[4], double &dl) {
  dl = 0.0;
#pragma unroll
  for (int v = 0; v < 4; v++) { q0[v] = 0.0; fo[v] = 0.0; }
  if constexpr (UM != UM_RESID) {
#pragma unroll
    for (int v = 0; v < 4; v++) q0[v] = q[v * np + i];
    if (S.stage != 0) {
#pragma unroll
      for (int v = 0; v < 4; v++) fo[v] = f[v * np + i];
      if (STEADY) dl = dtl[i];
    }
  }
}

template <int UM, bool STEADY>
__device__ __forceinline__ void stage_update(const Phys &P, const StageParams &S, const int i, const int np, const double ivol,
                                             const double *__restrict__ vol_arr,
                                             const double acc[4], const double wsacc, double *__restrict__ q,
                                             double *__restrict__ f, double *__restrict__ pout, double *__restrict__ dtl,
                                             double *__restrict__ resid_out, double *__restrict__ ws_out, double dq2[4]) {
  double q0[4], fo[4], dl;
  stage_load<UM, STEADY>(S, i, np, q, f, dtl, q0, fo, dl);
  stage_update_pre<UM, STEADY>(P, S, i, np, ivol, vol_arr, q0, fo, dl, acc, wsacc, q, f, pout, dtl, resid_out, ws_out, dq2);
}

// ------------------------------------------------------------------------------------------------
// pass B, direct-gather variant: neighbour data straight from global memory (works for any mesh
// numbering; also the fallback when a tile's halo does not fit the packed 16-bit slots).
// one cell of the direct-gather pass B: faces from global memory, then the stage update
template <int UM, bool STEADY, int RC>
__device__ __forceinline__ void flux_rk_cell(const DevMesh &m, const Phys &P, const StageParams &S, const double *__restrict__ p,
                                             const double *__restrict__ g, const double *__restrict__ phi, const double *__restrict__ bc,
                                             double *__restrict__ q, double *__restrict__ f, double *__restrict__ pout,
                                             double *__restrict__ dtl, double *__restrict__ resid_out, double *__restrict__ ws_out,
                                             const int i, double dq2[4]) {
  const int np = m.np, lane = i & 31;
  const double2 *p2 = reinterpret_cast<const double2 *>(p), *g2 = reinterpret_cast<const double2 *>(g);
  const int sl = i >> 5;
  const int off = __ldg(&m.f_off[sl]);
  const int w = (__ldg(&m.f_off[sl + 1]) - off) >> 5;
  double p0[4], g0x[4], g0y[4];
  load4(p2, np, i, p0);
  if (RC != RC_FIRST) { load4(g2, np, i, g0x); load4(g2 + 2 * (size_t)np, np, i, g0y); }
  const double2 c0 = m.xy[i];
  const double phi0 = (RC >= RC_K0_PHI) ? phi[i] : 1.0;
  // the cell's Runge-Kutta data and all of its face words are requested up front: one memory round trip instead of one per
  // face (on meshes that live in L2 a stage is bound by these dependent latencies, not by bandwidth)
  double q0[4], fo[4], dl;
  stage_load<UM, STEADY>(S, i, np, q, f, dtl, q0, fo, dl);
  const double ivol = m.ivol[i];
  int nbv[4], fev[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    nbv[k] = k < w ? __ldg(&m.f_nbr[off + 32 * k + lane]) : kPadNbr;
    fev[k] = k < w ? __ldg(&m.f_edge[off + 32 * k + lane]) : 0;
  }
  double acc[4] = {0.0, 0.0, 0.0, 0.0}, wsacc = 0.0;
  // twice the flux of face k in the edge's orientation, its wave speed, and the signed half length it is added with
  auto face = [&](const int k, double fl[4], double &ws, double &ha, double &sa) {
    const int nb = nbv[k], fe = fev[k];
    const int ed = fe >> 1;
    const bool self_c1 = (fe & 1) == 0;
    const double2 fc = __ldg(&m.exy[ed]), fn = __ldg(&m.enxy[ed]);
    const double af = __ldg(&m.ea[ed]);
    double me[4] = {0.0, 0.0, 0.0, 0.0};  // this cell's reconstruction increment (x_f - x_c) . grad p (RC_K0: the state)
    if (RC != RC_FIRST) {
      const double dx = fc.x - c0.x, dy = fc.y - c0.y;
#pragma unroll
      for (int v = 0; v < 4; v++) me[v] = RC == RC_K0 ? recon_k0(p0[v], g0x[v], g0y[v], dx, dy) : dx * g0x[v] + dy * g0y[v];
    }
    double sL[4], sR[4];
    if (nb >= 0) {
      double pj[4], ot[4] = {0.0, 0.0, 0.0, 0.0};
      load4(p2, np, nb, pj);
      double phij = 1.0;
      if (RC != RC_FIRST) {
        double gjx[4], gjy[4];
        load4(g2, np, nb, gjx);
        load4(g2 + 2 * (size_t)np, np, nb, gjy);
        const double2 cj = m.xy[nb];
        const double dx = fc.x - cj.x, dy = fc.y - cj.y;
#pragma unroll
        for (int v = 0; v < 4; v++) ot[v] = RC == RC_K0 ? recon_k0(pj[v], gjx[v], gjy[v], dx, dy) : dx * gjx[v] + dy * gjy[v];
        if (RC >= RC_K0_PHI) phij = phi[nb];
      }
      interior_states<RC>(P, self_c1, p0, me, phi0, pj, ot, phij, sL, sR);
      sa = self_c1 ? 0.5 * af : -(0.5 * af);
    } else {
      const int b = -1 - nb;
      const int type = __ldg(&m.bf_type[b]);
      double bcv[4];
#pragma unroll
      for (int v = 0; v < 4; v++) bcv[v] = bc[v * m.nbf + b];  // (plain load: the cooperative step kernel writes bc itself)
      boundary_states<RC>(type, p0, me, phi0, bcv, fn.x, fn.y, sL, sR);
      sa = 0.5 * af;
    }
    roe_flux2(P, sL, sR, fn.x, fn.y, fl, ws);
    ha = 0.5 * af;
  };
  auto add = [&](const double fl[4], const double ws, const double ha, const double sa) {
#pragma unroll
    for (int v = 0; v < 4; v++) acc[v] += fl[v] * sa;
    wsacc += ws * ha;
  };
  {
    // two faces per basic block, so that their independent flux evaluations interleave; accumulated in face order
    int k = 0;
#pragma unroll 1
    for (; k + 1 < w; k += 2) {
      const bool v0 = nbv[k] != kPadNbr, v1 = nbv[k + 1] != kPadNbr;
      if (v0 && v1) {
        double fa[4], fb[4], wa, wb, ha, hb, sa_, sb_;
        face(k, fa, wa, ha, sa_);
        face(k + 1, fb, wb, hb, sb_);
        add(fa, wa, ha, sa_);
        add(fb, wb, hb, sb_);
      } else if (v0 || v1) {
        double fa[4], wa, ha, sa_;
        face(v0 ? k : k + 1, fa, wa, ha, sa_);
        add(fa, wa, ha, sa_);
      }
    }
    if (k < w && nbv[k] != kPadNbr) {
      double fa[4], wa, ha, sa_;
      face(k, fa, wa, ha, sa_);
      add(fa, wa, ha, sa_);
    }
  }
  stage_update_pre<UM, STEADY>(P, S, i, np, ivol, m.vol, q0, fo, dl, acc, wsacc, q, f, pout, dtl, resid_out, ws_out, dq2);
}

template <int UM, bool STEADY, int RC>
__global__ void __launch_bounds__(kBlock) k_flux_rk(const DevMesh m, const Phys P, const StageParams S,
                                                    const double *__restrict__ p, const double *__restrict__ g,
                                                    const double *__restrict__ phi, const double *__restrict__ bc,
                                                    double *__restrict__ q, double *__restrict__ f, double *__restrict__ pout,
                                                    double *__restrict__ dtl, double *__restrict__ resid_out,
                                                    double *__restrict__ ws_out, double *__restrict__ partial) {
  pdl_entry();
  const int i = blockIdx.x * kBlock + threadIdx.x;
  double dq2[4] = {0.0, 0.0, 0.0, 0.0};
  if (i < m.n_own) flux_rk_cell<UM, STEADY, RC>(m, P, S, p, g, phi, bc, q, f, pout, dtl, resid_out, ws_out, i, dq2);
  if (UM != UM_RESID && S.last) block_sum_store<4>(dq2, partial);
}

// ------------------------------------------------------------------------------------------------
// Small meshes: TWO threads per cell.  A mesh of a few ten thousand cells fills only a fraction of the machine's
// thread slots with one thread per cell, and a pass then lasts as long as the dependent chain of a single thread (C2:
// ~10 us per pass, whatever launches it -- profiles/r2_small_meshes.md).  The two lanes of a pair split the chain:
//   pass A  lane h takes the variable pair (rho,u) / (v,p) -- the pair-interleaved layout makes that one 16-byte load per
//           stencil member -- and half of the limiter evaluations; phi = fmin of the two lanes' minima (exact);
//   pass B  lane h takes the faces k = h, h + 2; lane 1's fluxes travel to lane 0 by shuffles and are added there in
//           face order with the same fused multiply-adds as k_flux_rk, so the state is bitwise the one-thread kernel's;
//           lane 0 applies the stage update.
template <int FORM, bool LIM>
__global__ void __launch_bounds__(kBlock) k_gradient2(const DevMesh m, const int limiter_type, const double *__restrict__ p,
                                                      double *__restrict__ g, double *__restrict__ phi) {
  pdl_entry();
  const int t = blockIdx.x * kBlock + threadIdx.x, half = t & 1;
  const bool live = (t >> 1) < m.n_own;
  const int i = live ? t >> 1 : m.n_own - 1;  // (idle pairs redo the last cell and store nothing: the shuffles stay full-warp)
  const int np = m.np, lane = i & 31, sl = i >> 5;
  const int off = __ldg(&m.g_off[sl]);
  const int w = (__ldg(&m.g_off[sl + 1]) - off) >> 5;
  const double2 *p2 = reinterpret_cast<const double2 *>(p) + (size_t)half * np;
  const double2 q0 = p2[i];
  double ax0, ax1, ay0, ay1, mn0 = q0.x, mx0 = q0.x, mn1 = q0.y, mx1 = q0.y;
  if (FORM == 0) {
    const double c0x = m.c0x[i], c0y = m.c0y[i];
    ax0 = c0x * q0.x; ay0 = c0y * q0.x; ax1 = c0x * q0.y; ay1 = c0y * q0.y;
  } else {
    ax0 = ax1 = ay0 = ay1 = 0.0;
  }
  for (int k = 0; k < w; k++) {
    const int e = off + 32 * k + lane;
    const int j = __ldg(&m.g_idx[e]);
    const double cx = __ldg(&m.g_cx[e]), cy = __ldg(&m.g_cy[e]);
    const double2 pj = p2[j];
    const double d0 = FORM == 0 ? pj.x : pj.x - q0.x, d1 = FORM == 0 ? pj.y : pj.y - q0.y;
    ax0 += cx * d0;
    ay0 += cy * d0;
    ax1 += cx * d1;
    ay1 += cy * d1;
    if (LIM) { mn0 = fmin(mn0, pj.x); mx0 = fmax(mx0, pj.x); mn1 = fmin(mn1, pj.y); mx1 = fmax(mx1, pj.y); }
  }
  if (live) {
    double2 *g2 = reinterpret_cast<double2 *>(g);
    g2[(size_t)half * np + i] = make_double2(ax0, ax1);
    g2[(size_t)(2 + half) * np + i] = make_double2(ay0, ay1);
  }
  if (LIM) {
    const double pi = 3.141592653589793238462643383279502884;
    const double xc = m.xy[i].x, yc = m.xy[i].y;
    const double h = 2.0 * sqrt(m.vol[i] / pi);
    const double kh = (limiter_type == 1 ? 5.0 : 0.3) * h;
    const double eps2 = kh * kh * kh;
    const int foff = __ldg(&m.f_off[sl]);
    const int fw = (__ldg(&m.f_off[sl + 1]) - foff) >> 5;
    double ph = 1.0;
    for (int k = 0; k < fw; k++) {
      const int e = foff + 32 * k + lane;
      if (__ldg(&m.f_nbr[e]) == kPadNbr) continue;
      const int ed = __ldg(&m.f_edge[e]) >> 1;
      const double2 ec = __ldg(&m.exy[ed]);
      const double dx = ec.x - xc, dy = ec.y - yc;
      {
        const double pf = q0.x + dx * ax0 + dy * ay0;
        const double diff = pf - q0.x;
        const double f = limiter_fn(limiter_type, (diff > 0.0 ? mx0 : mn0) - q0.x, diff, eps2);
        ph = fmin(ph, diff != 0.0 ? f : 1.0);
      }
      {
        const double pf = q0.y + dx * ax1 + dy * ay1;
        const double diff = pf - q0.y;
        const double f = limiter_fn(limiter_type, (diff > 0.0 ? mx1 : mn1) - q0.y, diff, eps2);
        ph = fmin(ph, diff != 0.0 ? f : 1.0);
      }
    }
    ph = fmin(ph, __shfl_xor_sync(0xffffffffu, ph, 1));
    if (live && half == 0) phi[i] = ph;
  }
}

template <int UM, bool STEADY, int RC>
__global__ void __launch_bounds__(kBlock) k_flux_rk2(const DevMesh m, const Phys P, const StageParams S,
                                                     const double *__restrict__ p, const double *__restrict__ g,
                                                     const double *__restrict__ phi, const double *__restrict__ bc,
                                                     double *__restrict__ q, double *__restrict__ f, double *__restrict__ pout,
                                                     double *__restrict__ dtl, double *__restrict__ resid_out,
                                                     double *__restrict__ ws_out, double *__restrict__ partial) {
  pdl_entry();
  const int t = blockIdx.x * kBlock + threadIdx.x, half = t & 1;
  const bool live = (t >> 1) < m.n_own;
  const int i = live ? t >> 1 : m.n_own - 1;
  const int np = m.np, lane = i & 31, sl = i >> 5;
  const double2 *p2 = reinterpret_cast<const double2 *>(p), *g2 = reinterpret_cast<const double2 *>(g);
  const int off = __ldg(&m.f_off[sl]);
  const int w = (__ldg(&m.f_off[sl + 1]) - off) >> 5;
  double q0[4], fo[4], dl = 0.0, ivol = 1.0;
  if (half == 0) { stage_load<UM, STEADY>(S, i, np, q, f, dtl, q0, fo, dl); ivol = m.ivol[i]; }
  double p0[4], g0x[4], g0y[4];
  load4(p2, np, i, p0);
  if (RC != RC_FIRST) { load4(g2, np, i, g0x); load4(g2 + 2 * (size_t)np, np, i, g0y); }
  const double2 c0 = m.xy[i];
  const double phi0 = (RC >= RC_K0_PHI) ? phi[i] : 1.0;
  int nbv[2], fev[2];
#pragma unroll
  for (int r = 0; r < 2; r++) {
    const int k = 2 * r + half;
    nbv[r] = k < w ? __ldg(&m.f_nbr[off + 32 * k + lane]) : kPadNbr;
    fev[r] = k < w ? __ldg(&m.f_edge[off + 32 * k + lane]) : 0;
  }
  double acc[4] = {0.0, 0.0, 0.0, 0.0}, wsacc = 0.0;
#pragma unroll
  for (int r = 0; r < 2; r++) {  // round r: faces 2r (lane 0) and 2r + 1 (lane 1)
    const int nb = nbv[r], fe = fev[r];
    const bool valid = nb != kPadNbr;
    double fl[4] = {0.0, 0.0, 0.0, 0.0}, ws = 0.0, ha = 0.0, sa = 0.0;
    if (valid) {
      const int ed = fe >> 1;
      const bool self_c1 = (fe & 1) == 0;
      const double2 fc = __ldg(&m.exy[ed]), fn = __ldg(&m.enxy[ed]);
      const double af = __ldg(&m.ea[ed]);
      double me[4] = {0.0, 0.0, 0.0, 0.0};
      if (RC != RC_FIRST) {
        const double dx = fc.x - c0.x, dy = fc.y - c0.y;
#pragma unroll
        for (int v = 0; v < 4; v++) me[v] = RC == RC_K0 ? recon_k0(p0[v], g0x[v], g0y[v], dx, dy) : dx * g0x[v] + dy * g0y[v];
      }
      double sL[4], sR[4];
      if (nb >= 0) {
        double pj[4], ot[4] = {0.0, 0.0, 0.0, 0.0};
        load4(p2, np, nb, pj);
        double phij = 1.0;
        if (RC != RC_FIRST) {
          double gjx[4], gjy[4];
          load4(g2, np, nb, gjx);
          load4(g2 + 2 * (size_t)np, np, nb, gjy);
          const double2 cj = m.xy[nb];
          const double dx = fc.x - cj.x, dy = fc.y - cj.y;
#pragma unroll
          for (int v = 0; v < 4; v++) ot[v] = RC == RC_K0 ? recon_k0(pj[v], gjx[v], gjy[v], dx, dy) : dx * gjx[v] + dy * gjy[v];
          if (RC >= RC_K0_PHI) phij = phi[nb];
        }
        interior_states<RC>(P, self_c1, p0, me, phi0, pj, ot, phij, sL, sR);
        sa = self_c1 ? 0.5 * af : -(0.5 * af);
      } else {
        const int b = -1 - nb;
        const int type = __ldg(&m.bf_type[b]);
        double bcv[4];
#pragma unroll
        for (int v = 0; v < 4; v++) bcv[v] = bc[v * m.nbf + b];
        boundary_states<RC>(type, p0, me, phi0, bcv, fn.x, fn.y, sL, sR);
        sa = 0.5 * af;
      }
      roe_flux2(P, sL, sR, fn.x, fn.y, fl, ws);
      ha = 0.5 * af;
    }
    // lane 1's face joins lane 0's sum: face 2r first, then face 2r + 1 (a skipped face adds nothing, as in k_flux_rk)
    double ofl[4];
#pragma unroll
    for (int v = 0; v < 4; v++) ofl[v] = __shfl_xor_sync(0xffffffffu, fl[v], 1);
    const double ows = __shfl_xor_sync(0xffffffffu, ws, 1), oha = __shfl_xor_sync(0xffffffffu, ha, 1), osa = __shfl_xor_sync(0xffffffffu, sa, 1);
    const bool ovalid = __shfl_xor_sync(0xffffffffu, valid ? 1 : 0, 1) != 0;
    if (half == 0) {
      if (valid) {
#pragma unroll
        for (int v = 0; v < 4; v++) acc[v] += fl[v] * sa;
        wsacc += ws * ha;
      }
      if (ovalid) {
#pragma unroll
        for (int v = 0; v < 4; v++) acc[v] += ofl[v] * osa;
        wsacc += ows * oha;
      }
    }
  }
  double dq2[4] = {0.0, 0.0, 0.0, 0.0};
  if (live && half == 0) stage_update_pre<UM, STEADY>(P, S, i, np, ivol, m.vol, q0, fo, dl, acc, wsacc, q, f, pout, dtl, resid_out, ws_out, dq2);
  if (UM != UM_RESID && S.last) block_sum_store<4>(dq2, partial);
}

// ------------------------------------------------------------------------------------------------
// PTX helpers of the shared-memory pipeline: mbarrier, TMA bulk copy (cp.async.bulk), cp.async
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async8(void *dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(void *dst, const void *src) {  // L1-bypassing (LDGSTS.BYPASS)
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}

// ------------------------------------------------------------------------------------------------
// pass B, persistent warp-specialised pipeline (the production path).
// CTA = kBlock/32 consumer warps (one thread per cell of a tile) + 1 producer warp, looping over tiles
// blockIdx.x, blockIdx.x + gridDim.x, ...  A 2-stage shared-memory ring decouples them:
//   producer: waits empty[s]; one lane issues the TMA bulk copies (cp.async.bulk, complete_tx on full[s]) for
//             the tile's contiguous runs -- own cells of p (2 x double2), g (4), xy (1) [, phi], own edges
//             exy / enxy / ea, and the face table; all lanes gather the halo cells / edges listed for the
//             tile with 16-byte L1-bypassing cp.async.cg (8 bytes for phi / ea) and hand their completion to
//             full[s] (cp.async.mbarrier.arrive.noinc); the next tile's header and index lists are prefetched;
//   consumer: prefetches its cell's RK data (registers), waits full[s], computes its faces out of shared
//             memory with 16-byte LDS (left/right states are addressed by slot, so no operand swapping),
//             arrives on empty[s], applies the stage update.
// Global-memory latency is only ever seen by the producer warp.
// (Measured on B200: with 8-byte cp.async, which can only go THROUGH L1, the gathers were throttled by L1
// line allocation -- pass B ran 25 % slower whenever 3 CTAs' shared memory pushed the carve-out from 196 to
// 228 KB; the pair-interleaved layout exists to make every gather a 16-byte bypass copy.)
constexpr int kPipeThreads = kBlock + 32;
#ifndef FVS2D_STAGES
#define FVS2D_STAGES 2
#endif
#ifndef FVS2D_FACE2
#define FVS2D_FACE2 1   // two interior faces per loop iteration of a consumer thread (0: one; measured 2-3 % slower)
#endif
constexpr int kStages = FVS2D_STAGES;


struct PipeMeta {
  const int4 *hdr;  // 2 x int4 per tile: {es, ne, hc_ptr, n_hc}, {he_ptr, n_he, fbase, fw}
  const int *hc_idx, *he_idx;
  const uint32_t *t_pack;
  const int *t_bf;
  int S, E, ntiles;      // smem strides in elements (even): cell slots (kBlock + max halo), edge slots
  const int *tile_list;  // null: tiles 0..ntiles-1; else the ntiles tile ids to process (interior / boundary subset)
};

// bytes of one stage: double2 cell arrays [NC2][S] | phi [S] (limited variants) | exy, enxy [2][E] | ea [E] |
// face table [4][kBlock] | {fw, fbase}
template <int RC>
__host__ __device__ constexpr int pipe_nc2() { return RC == RC_FIRST ? 2 : 7; }
template <int RC>
__host__ __device__ inline size_t pipe_stage_bytes(int S, int E) {
  return (size_t)pipe_nc2<RC>() * S * 16 + (RC >= RC_K0_PHI ? (size_t)S * 8 : 0) + (size_t)E * 40 + 4 * kBlock * sizeof(uint32_t) + 16;
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t *bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int UM, bool STEADY, int RC>
__global__ void __launch_bounds__(kPipeThreads, FVS2D_PIPE_CTAS) k_flux_pipe(const DevMesh m, const PipeMeta pm, const Phys P, const StageParams S,
                                                               const double *__restrict__ p, const double *__restrict__ g,
                                                               const double *__restrict__ phi, const double *__restrict__ bc,
                                                               double *__restrict__ q, double *__restrict__ f,
                                                               double *__restrict__ pout, double *__restrict__ dtl,
                                                               double *__restrict__ resid_out, double *__restrict__ ws_out,
                                                               double *__restrict__ partial) {
  pdl_entry();
  constexpr int NC2 = pipe_nc2<RC>();
  constexpr bool PHI = RC >= RC_K0_PHI;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int SS = pm.S, EE = pm.E, np = m.np;
  const size_t stage_bytes = pipe_stage_bytes<RC>(SS, EE);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + kStages * stage_bytes);
  uint64_t *empty = full + kStages;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // stage layout helpers
  auto st_c2 = [&](int s) { return reinterpret_cast<double2 *>(smem_raw + s * stage_bytes); };
  auto st_phi = [&](int s) { return reinterpret_cast<double *>(st_c2(s) + NC2 * SS); };
  auto st_e2 = [&](int s) { return reinterpret_cast<double2 *>(st_phi(s) + (PHI ? SS : 0)); };
  auto st_ea = [&](int s) { return reinterpret_cast<double *>(st_e2(s) + 2 * EE); };
  auto st_f = [&](int s) { return reinterpret_cast<uint32_t *>(st_ea(s) + EE); };
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStages; s++) { mbar_init(&full[s], 33); mbar_init(&empty[s], kBlock); }
  }
  __syncthreads();

  if (warp == kBlock / 32) {
    // ================================ producer warp ================================
    const double2 *p2 = reinterpret_cast<const double2 *>(p), *g2 = reinterpret_cast<const double2 *>(g);
    auto cell_src = [&](int a) -> const double2 * { return a < 2 ? p2 + (size_t)a * np : a < 6 ? g2 + (size_t)(a - 2) * np : m.xy; };
    // the header and the halo index lists of the NEXT tile are fetched while the current one is in
    // flight, so only one memory latency (the data itself) sits between "stage free" and "stage full"
    int4 h0 = make_int4(0, 0, 0, 0), h1 = make_int4(0, 0, 0, 0);
    int jc[3] = {0, 0, 0}, je[3] = {0, 0, 0};
    auto fetch_meta = [&](int t) {
      h0 = __ldg(&pm.hdr[2 * t]);
      h1 = __ldg(&pm.hdr[2 * t + 1]);
#pragma unroll
      for (int r = 0; r < 3; r++) {
        jc[r] = (lane + 32 * r < h0.w) ? __ldg(&pm.hc_idx[h0.z + lane + 32 * r]) : 0;
        je[r] = (lane + 32 * r < h1.y) ? __ldg(&pm.he_idx[h1.x + lane + 32 * r]) : 0;
      }
    };
    auto tile_id = [&](int j) { return pm.tile_list ? __ldg(&pm.tile_list[j]) : j; };
    auto gather_cell = [&](int s, int h, int j) {
      double2 *c2 = st_c2(s);
#pragma unroll
      for (int a = 0; a < NC2; a++) cp_async16(c2 + a * SS + kBlock + h, cell_src(a) + j);
      if (PHI) cp_async8(st_phi(s) + kBlock + h, phi + j);
    };
    auto gather_edge = [&](int s, int ne, int h, int j) {
      double2 *e2 = st_e2(s);
      cp_async16(e2 + ne + h, m.exy + j);
      cp_async16(e2 + EE + ne + h, m.enxy + j);
      cp_async8(st_ea(s) + ne + h, m.ea + j);
    };
    if ((int)blockIdx.x < pm.ntiles) fetch_meta(tile_id(blockIdx.x));
    int it = 0;
    for (int j = blockIdx.x; j < pm.ntiles; j += gridDim.x, it++) {
      const int t = tile_id(j);
      const int s = it % kStages;
      const uint32_t ph = (it / kStages) & 1;
      const int es = h0.x, ne = h0.y, hp = h0.z, nh = h0.w, ep = h1.x, nhe = h1.y, fbase = h1.z, fw = h1.w;
      const int jcc[3] = {jc[0], jc[1], jc[2]}, jee[3] = {je[0], je[1], je[2]};
      mbar_wait(&empty[s], ph ^ 1);
      const int c0 = t * kBlock;
      const int ncell = min(kBlock, m.n_own - c0);
      if (lane == 0) {
        double2 *c2 = st_c2(s), *e2 = st_e2(s);
        uint32_t *sf = st_f(s);
        const uint32_t bytes_c = (uint32_t)ncell * 16u, bytes_phi = PHI ? (uint32_t)((ncell + 1) & ~1) * 8u : 0u;
        const uint32_t bytes_e = (uint32_t)ne * 16u, bytes_ea = (uint32_t)ne * 8u, bytes_f = (uint32_t)fw * kBlock * 4u;
        int *sh = reinterpret_cast<int *>(sf + 4 * kBlock);
        sh[0] = fw; sh[1] = fbase;  // published to the consumers by the arrive below (release)
        mbar_expect_tx(&full[s], NC2 * bytes_c + bytes_phi + 2u * bytes_e + bytes_ea + bytes_f);
#pragma unroll
        for (int a = 0; a < NC2; a++) bulk_g2s(c2 + a * SS, cell_src(a) + c0, bytes_c, &full[s]);
        if (PHI) bulk_g2s(st_phi(s), phi + c0, bytes_phi, &full[s]);
        if (ne > 0) {
          bulk_g2s(e2, m.exy + es, bytes_e, &full[s]);
          bulk_g2s(e2 + EE, m.enxy + es, bytes_e, &full[s]);
          bulk_g2s(st_ea(s), m.ea + es, bytes_ea, &full[s]);
        }
        if (fw > 0) bulk_g2s(sf, pm.t_pack + fbase, bytes_f, &full[s]);
      }
#pragma unroll
      for (int r = 0; r < 3; r++) {
        const int h = lane + 32 * r;
        if (h < nh) gather_cell(s, h, jcc[r]);
        if (h < nhe) gather_edge(s, ne, h, jee[r]);
      }
      for (int h = lane + 96; h < nh; h += 32) gather_cell(s, h, __ldg(&pm.hc_idx[hp + h]));  // rare: more than 96 halo cells
      for (int h = lane + 96; h < nhe; h += 32) gather_edge(s, ne, h, __ldg(&pm.he_idx[ep + h]));
      cp_async_mbar_arrive_noinc(&full[s]);
      if (j + (int)gridDim.x < pm.ntiles) fetch_meta(tile_id(j + gridDim.x));
    }
    return;
  }

  // ================================== consumer warps ==================================
  double dq2[4] = {0.0, 0.0, 0.0, 0.0};
  int it = 0;
  for (int j = blockIdx.x; j < pm.ntiles; j += gridDim.x, it++) {
    const int t = pm.tile_list ? __ldg(&pm.tile_list[j]) : j;
    const int s = it % kStages;
    const uint32_t ph = (it / kStages) & 1;
    const int c0 = t * kBlock;
    const int ncell = min(kBlock, m.n_own - c0);
    const int i = c0 + tid;
    const bool live = tid < ncell;
    double q0[4], fo[4], dl = 0.0, ivol = 1.0;
    if (live) {  // RK data of this cell: in flight while the faces are computed
      stage_load<UM, STEADY>(S, i, np, q, f, dtl, q0, fo, dl);
      ivol = m.ivol[i];
    }
    const double2 *c2 = st_c2(s), *e2 = st_e2(s);
    const double *sphi = st_phi(s), *sea = st_ea(s);
    const uint32_t *sf = st_f(s);
    mbar_wait(&full[s], ph);
    const int fw = reinterpret_cast<const int *>(sf + 4 * kBlock)[0], fbase = reinterpret_cast<const int *>(sf + 4 * kBlock)[1];
    double acc[4] = {0.0, 0.0, 0.0, 0.0}, wsacc = 0.0;
    if (live) {
      // one face: reconstruct both sides from shared memory (addressed by slot -- no operand swapping), Roe flux
      // in the edge's own orientation (L = c1, R = c2), accumulate.  BND selects the boundary variant (this cell
      // is c1, right state from the boundary condition); boundary faces come last in a cell's list, so handling
      // them in a second, rarely taken loop keeps the reference's accumulation order and keeps the
      // boundary-condition loads out of the common path.
      auto face = [&](const uint32_t pk, const int k, const auto bnd_tag) {
        constexpr bool BND = decltype(bnd_tag)::value;
        const int ns = pk & 0xFFFFu, eslot = (pk >> 16) & 0x7FFF;
        const bool self_c1 = BND || (pk >> 31) == 0;
        const double2 fc = e2[eslot], fn = e2[EE + eslot];
        const double af = sea[eslot], nx = fn.x, ny = fn.y;
        const int sl_ = self_c1 ? tid : ns, sr_ = (self_c1 && !BND) ? ns : tid;
        const double2 *cl = c2 + sl_, *cr = c2 + sr_;
        double sL[4], sR[4];
        {
          const double2 a = cl[0], b = cl[SS], c = cr[0], d = cr[SS];
          sL[0] = a.x; sL[1] = a.y; sL[2] = b.x; sL[3] = b.y;
          sR[0] = c.x; sR[1] = c.y; sR[2] = d.x; sR[3] = d.y;
        }
        if (RC != RC_FIRST) {
          const double2 xl = cl[6 * SS], xr = cr[6 * SS];
          const double dxL = fc.x - xl.x, dyL = fc.y - xl.y, dxR = fc.x - xr.x, dyR = fc.y - xr.y;
          const double fL = PHI ? sphi[sl_] : 1.0, fR = PHI ? sphi[sr_] : 1.0;
          double gL[4], gR[4];
          {
            const double2 a = cl[2 * SS], b = cl[3 * SS], c = cl[4 * SS], d = cl[5 * SS];
            if (RC == RC_K0) {
              gL[0] = recon_k0(sL[0], a.x, c.x, dxL, dyL); gL[1] = recon_k0(sL[1], a.y, c.y, dxL, dyL);
              gL[2] = recon_k0(sL[2], b.x, d.x, dxL, dyL); gL[3] = recon_k0(sL[3], b.y, d.y, dxL, dyL);
            } else {
              gL[0] = dxL * a.x + dyL * c.x; gL[1] = dxL * a.y + dyL * c.y;
              gL[2] = dxL * b.x + dyL * d.x; gL[3] = dxL * b.y + dyL * d.y;
            }
          }
          {
            const double2 a = cr[2 * SS], b = cr[3 * SS], c = cr[4 * SS], d = cr[5 * SS];
            if (RC == RC_K0) {
              gR[0] = recon_k0(sR[0], a.x, c.x, dxR, dyR); gR[1] = recon_k0(sR[1], a.y, c.y, dxR, dyR);
              gR[2] = recon_k0(sR[2], b.x, d.x, dxR, dyR); gR[3] = recon_k0(sR[3], b.y, d.y, dxR, dyR);
            } else {
              gR[0] = dxR * a.x + dyR * c.x; gR[1] = dxR * a.y + dyR * c.y;
              gR[2] = dxR * b.x + dyR * d.x; gR[3] = dxR * b.y + dyR * d.y;
            }
          }
#pragma unroll
          for (int v = 0; v < 4; v++) {
            const double pL = sL[v], pR = sR[v];
            if (RC == RC_K0) { sL[v] = gL[v]; sR[v] = gR[v]; }  // already the states (recon_k0 above)
            else if (RC == RC_K0_PHI) { sL[v] = pL + fL * gL[v]; sR[v] = pR + fR * gR[v]; }
            else {
              // boundary faces carry no kappa term (src/residual.f90:128)
              const double gC = BND ? 0.0 : pR - pL, k1 = BND ? 1.0 : 1.0 - P.kappa;
              sL[v] = pL + fL * (P.kappa / 2.0 * gC + k1 * gL[v]);
              sR[v] = pR + fR * (-P.kappa / 2.0 * gC + k1 * gR[v]);
            }
          }
        }
        if (BND) {
          const int b = __ldg(&pm.t_bf[fbase + k * kBlock + tid]);
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
#if FVS2D_FACE2
      {
        // two interior faces per iteration in one basic block, so the two independent flux evaluations interleave
        int k = 0;
#pragma unroll 1
        for (; k + 1 < fw; k += 2) {
          const uint32_t pk0 = sf[k * kBlock + tid], pk1 = sf[(k + 1) * kBlock + tid];
          const uint32_t n0 = pk0 & 0xFFFFu, n1 = pk1 & 0xFFFFu;
          has_bnd = has_bnd || n0 == 0xFFFFu || n1 == 0xFFFFu;
          if (n0 < 0xFFFEu && n1 < 0xFFFEu) {
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
#else
      uint32_t pk_next = fw > 0 ? sf[tid] : 0xFFFEu;
#pragma unroll 1
      for (int k = 0; k < fw; k++) {
        const uint32_t pk = pk_next;
        if (k + 1 < fw) pk_next = sf[(k + 1) * kBlock + tid];  // next face's table word: its latency hides behind this face
        const uint32_t ns = pk & 0xFFFFu;
        if (ns == 0xFFFEu) continue;
        if (ns == 0xFFFFu) { has_bnd = true; continue; }
        face(pk, k, std::false_type{});
      }
#endif
      if (has_bnd) {
#pragma unroll 1
        for (int k = 0; k < fw; k++) {
          const uint32_t pk = sf[k * kBlock + tid];
          if ((pk & 0xFFFFu) == 0xFFFFu) face(pk, k, std::true_type{});
        }
      }
    }
    mbar_arrive(&empty[s]);  // this thread is done reading stage s
    if (live) stage_update_pre<UM, STEADY>(P, S, i, np, ivol, m.vol, q0, fo, dl, acc, wsacc, q, f, pout, dtl, resid_out, ws_out, dq2);
  }
  if (UM != UM_RESID && S.last) {
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
constexpr int kVortexCtas = 6;  // resident CTAs per SM of k_vortex_err; its grid is exactly one such wave
// K10: vortex error norms over the interior cells (src/mms.f90:315-361).  Per CTA:
// partial[b*13 + 0..3] = max |dq_v|, [4..7] = sum |dq_v|, [8..11] = sum dq_v^2, [12] = best rho error;
// best_id[b] = original id of the first cell attaining it.
__global__ void __launch_bounds__(kBlock, kVortexCtas) k_vortex_err(const DevMesh m, const Phys P, const StepClock *__restrict__ clk,
                                                          const double *__restrict__ q, double *__restrict__ partial,
                                                          int *__restrict__ best_id, int *__restrict__ best_loc) {
  pdl_entry();
  const double time = clock_told(clk) + clk->off_end;
  // grid-stride over the owned cells with a fixed grid (one resident wave): the partial count is small and the
  // summation order is a function of the launch configuration only (deterministic)
  const int np = m.np;
  double dmx[4] = {0, 0, 0, 0}, ds1[4] = {0, 0, 0, 0}, ds2[4] = {0, 0, 0, 0};
  double bv = -1.0;
  int bi = 0x7fffffff, bl = 0;  // best: value, original id (tie-break: the reference's maxloc takes the first), local id
  for (int i = blockIdx.x * kBlock + threadIdx.x; i < m.n_own; i += gridDim.x * kBlock) {
    // all loads of the cell are issued together; boundary cells (few) are computed and masked out
    const bool in = m.is_intr[i] != 0;
    const double2 cc = m.xy[i];
    const double a0 = q[i], a1 = q[np + i], a2 = q[2 * np + i], a3 = q[3 * np + i];
    double pv[4];
    vortex_exact(P, time, cc.x, cc.y, pv);
    const double ex0 = pv[0], ex1 = pv[0] * pv[1], ex2 = pv[0] * pv[2];
    const double ex3 = pv[3] / (P.gamma - 1.0) + 0.5 * pv[0] * (pv[1] * pv[1] + pv[2] * pv[2]);
    double d[4];
    d[0] = fabs(a0 - ex0);
    d[1] = fabs(a1 - ex1);
    d[2] = fabs(a2 - ex2);
    d[3] = fabs(a3 - ex3);
    if (in) {
#pragma unroll
      for (int v = 0; v < 4; v++) { dmx[v] = fmax(dmx[v], d[v]); ds1[v] += d[v]; ds2[v] += d[v] * d[v]; }
      if (d[0] >= bv) {  // rare after the first few cells: only then is the original id needed
        const int oid = m.orig_id[i];
        if (d[0] > bv || oid < bi) { bv = d[0]; bi = oid; bl = i; }
      }
    }
  }
  __shared__ double smx[4][kBlock / 32], s1[4][kBlock / 32], s2[4][kBlock / 32], sb[kBlock / 32];
  __shared__ int sid[kBlock / 32], sil[kBlock / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_down_sync(0xffffffffu, bv, o);
    const int oi = __shfl_down_sync(0xffffffffu, bi, o), ol = __shfl_down_sync(0xffffffffu, bl, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; bl = ol; }
  }
#pragma unroll
  for (int v = 0; v < 4; v++) {
    double mx = dmx[v], a = ds1[v], b = ds2[v];
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
  if (threadIdx.x < 4) {
    const int v = threadIdx.x;
    double mx = 0, a = 0, b = 0;
    for (int w = 0; w < kBlock / 32; w++) { mx = fmax(mx, smx[v][w]); a += s1[v][w]; b += s2[v][w]; }
    partial[(size_t)blockIdx.x * 13 + v] = mx;
    partial[(size_t)blockIdx.x * 13 + 4 + v] = a;
    partial[(size_t)blockIdx.x * 13 + 8 + v] = b;
  }
  if (threadIdx.x == 0) {
    double bb = sb[0]; int ii = sid[0], ll = sil[0];
    for (int w = 1; w < kBlock / 32; w++)
      if (sb[w] > bb || (sb[w] == bb && sid[w] < ii)) { bb = sb[w]; ii = sid[w]; ll = sil[w]; }
    partial[(size_t)blockIdx.x * 13 + 12] = bb;
    best_id[blockIdx.x] = ii;
    best_loc[blockIdx.x] = ll;
  }
}

// Last kernel of a time step: the fixed-order final reductions of the per-CTA partials into this step's row of the
// device log, then the step counter advances.  One warp per quantity (lane-strided partial sums, shuffle tree), so
// there is no block-wide synchronisation until the clock update:
//   warps 0..3   row[v]      = sum_b partial[b*4+v]        Sigma (q-q0)^2 of the step (src/runge_kutta.f90:169-184)
//   warps 4..15  row[4+v]    = max (v<4) / sum of vpartial[b*13+v]                     (src/mms.f90:315-361)
//   warp  16     row[16], logid = largest rho error and the original id of the first cell attaining it;
//                row[17], row[18] = that cell's centroid (src/mms.f90:363), so the host needs no global coordinate array
constexpr int kFinishThreads = 17 * 32;
__global__ void __launch_bounds__(kFinishThreads) k_finish_step(const double *__restrict__ partial, const int nparts,
                                                                const double *__restrict__ vpartial, const int *__restrict__ vbest,
                                                                const int *__restrict__ vbest_loc, const double2 *__restrict__ xy,
                                                                const int nvparts, double *__restrict__ logbuf, const int stride,
                                                                int *__restrict__ logid, StepClock *__restrict__ clk) {
  pdl_entry();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int istep = clk->istep;
  double *row = logbuf + (size_t)stride * istep;
  if (warp < 4) {
    double s = 0.0;
    for (int b = lane; b < nparts; b += 32) s += partial[(size_t)b * 4 + warp];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) row[warp] = s;
  } else if (warp < 16) {
    const int v = warp - 4;
    if (nvparts > 0) {
      double s = 0.0;
      for (int b = lane; b < nvparts; b += 32) {
        const double x = vpartial[(size_t)b * 13 + v];
        s = v < 4 ? fmax(s, x) : s + x;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double x = __shfl_down_sync(0xffffffffu, s, o);
        s = v < 4 ? fmax(s, x) : s + x;
      }
      if (lane == 0) row[4 + v] = s;
    }
  } else if (nvparts > 0) {
    double bv = -1.0;
    int bi = 0x7fffffff, bl = 0;
    for (int b = lane; b < nvparts; b += 32) {
      const double x = vpartial[(size_t)b * 13 + 12];
      const int id = vbest[b];
      if (x > bv || (x == bv && id < bi)) { bv = x; bi = id; bl = vbest_loc[b]; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double x = __shfl_down_sync(0xffffffffu, bv, o);
      const int id = __shfl_down_sync(0xffffffffu, bi, o), il = __shfl_down_sync(0xffffffffu, bl, o);
      if (x > bv || (x == bv && id < bi)) { bv = x; bi = id; bl = il; }
    }
    if (lane == 0) {
      const double2 c = xy[bl];
      row[16] = bv; row[17] = c.x; row[18] = c.y; logid[istep] = bi;
    }
  }
  __syncthreads();  // every warp has read istep
  if (threadIdx.x == 0) clk->istep = istep + 1;
}

// ------------------------------------------------------------------------------------------------
// halo pack: buf[v*n + k] = a[v*np + idx[k]] for nv arrays of elements T (double2 for the pair-interleaved
// arrays p and g, double for phi)
template <class T>
__global__ void __launch_bounds__(256) k_pack(int n, int nv, int np, const int *__restrict__ idx, const T *__restrict__ a,
                                               T *__restrict__ buf) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int i = idx[k];
  for (int v = 0; v < nv; v++) buf[(size_t)v * n + k] = a[(size_t)v * np + i];
}

// permuting copies between the caller's (nvar,ncells) AoS arrays (original numbering, staged on the device)
// and the device arrays (local numbering): pair != 0 selects the pair-interleaved layout (variables v0..),
// orig_id == null means the caller's array is already in local order
__global__ void __launch_bounds__(256) k_scatter_in(int n, int np, int nvar, const int *__restrict__ orig_id,
                                                     const double *__restrict__ aos, double *__restrict__ soa, int pair) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t o = orig_id ? orig_id[i] : i;
  for (int v = 0; v < nvar; v++) soa[pair ? pidx(v, np, i) : (size_t)v * np + i] = aos[o * nvar + v];
}
__global__ void __launch_bounds__(256) k_gather_out(int n, int np, int nvar, const int *__restrict__ orig_id,
                                                     const double *__restrict__ soa, double *__restrict__ aos, int pair, int v0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t o = orig_id ? orig_id[i] : i;
  for (int v = 0; v < nvar; v++) aos[o * nvar + v] = soa[pair ? pidx(v0 + v, np, i) : (size_t)(v0 + v) * np + i];
}


// ------------------------------------------------------------------------------------------------
// Output path (per save interval, not per stage): the node-interpolated primitive variables of write_inst_ios
// (src/io.f90:122-150, src/interpolation.f90:107-123).  One thread per node; the node's cells are visited in
// ascending ORIGINAL cell id -- the reference's summation order -- through their local ids, with the
// inverse-distance weights of cell2node_idw_setup (src/interpolation.f90:62-101) precomputed on upload.
// mask bit v selects primitive variable v; the k-th selected variable goes to fv[k*nnodes + node].
__global__ void __launch_bounds__(256) k_cell2node(const int nnodes, const int np, const int mask, const int *__restrict__ n2c_ptr,
                                                    const int *__restrict__ n2c, const double *__restrict__ idw,
                                                    const double *__restrict__ p, double *__restrict__ fv) {
  const int in = blockIdx.x * blockDim.x + threadIdx.x;
  if (in >= nnodes) return;
  const double2 *p2 = reinterpret_cast<const double2 *>(p);
  double a[4] = {0.0, 0.0, 0.0, 0.0};
  const int j1 = __ldg(&n2c_ptr[in + 1]);
  for (int j = __ldg(&n2c_ptr[in]); j < j1; j++) {
    const int ic = __ldg(&n2c[j]);
    const double w = __ldg(&idw[j]);
    double pv[4];
    load4(p2, np, ic, pv);
#pragma unroll
    for (int v = 0; v < 4; v++) a[v] = a[v] + w * pv[v];
  }
  int k = 0;
#pragma unroll
  for (int v = 0; v < 4; v++)
    if (mask >> v & 1) { fv[(size_t)k * nnodes + in] = a[v]; k++; }
}

// Output path: the numerical part of write_inst_cp_un (src/io.f90:340-449).  One thread per boundary edge of the
// requested boundary: p, u, v of the edge's cell (c1) extrapolated to the edge centre with the cell's unlimited
// gradient (pass A has just been run on the current state), out[4*i..] = x_f, p_w, p_cell, u_n.
__global__ void __launch_bounds__(128) k_wall_values(const int n, const int np, const int *__restrict__ cell,
                                                      const double2 *__restrict__ exy, const double2 *__restrict__ enxy,
                                                      const double2 *__restrict__ xy, const double *__restrict__ p,
                                                      const double *__restrict__ g, double *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int ic = cell[i];
  const double2 *p2 = reinterpret_cast<const double2 *>(p), *g2 = reinterpret_cast<const double2 *>(g);
  double pv[4], gx[4], gy[4];
  load4(p2, np, ic, pv);
  load4(g2, np, ic, gx);
  load4(g2 + 2 * (size_t)np, np, ic, gy);
  const double2 fc = exy[i], fn = enxy[i], cc = xy[ic];
  const double dx = fc.x - cc.x, dy = fc.y - cc.y;
  const double pw = pv[3] + dx * gx[3] + dy * gy[3];
  const double uw = pv[1] + dx * gx[1] + dy * gy[1];
  const double vw = pv[2] + dx * gx[2] + dy * gy[2];
  out[4 * (size_t)i + 0] = fc.x;
  out[4 * (size_t)i + 1] = pw;
  out[4 * (size_t)i + 2] = pv[3];
  out[4 * (size_t)i + 3] = uw * fn.x + vw * fn.y;
}

}  // namespace fvs2d
