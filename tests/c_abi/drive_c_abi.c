/* drive_c_abi.c -- a plain C host over include/fvs2d_gpu.h (no Python, no C++): what the Fortran shim of INTEGRATION.md
 * does, in C.  Builds an nx x ny split-quad triangle mesh on [0,20]x[0,10] (all-Dirichlet, isentropic vortex), then
 *     fvs2d_gpu_init -> fvs2d_gpu_set_mesh -> fvs2d_gpu_initialize_solution -> fvs2d_gpu_time_integration(nsteps)
 *     -> fvs2d_gpu_get_state -> fvs2d_gpu_compute_residual -> fvs2d_gpu_finalize
 * and writes the state (binary doubles) plus the residual history to the files given on the command line, for
 * tests/test_gpu_c_abi.py to compare with the CPU oracle.   usage: drive_c_abi nx ny nsteps state.bin log_res.txt */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fvs2d_gpu.h"

#define CHECK(call)                                                        \
  do {                                                                     \
    if ((call) != 0) {                                                     \
      fprintf(stderr, "%s failed: %s\n", #call, fvs2d_gpu_last_error());   \
      return 2;                                                            \
    }                                                                      \
  } while (0)

int main(int argc, char **argv) {
  if (argc < 6) { fprintf(stderr, "usage: %s nx ny nsteps state.bin log_res.txt\n", argv[0]); return 1; }
  const int nx = atoi(argv[1]), ny = atoi(argv[2]), nsteps = atoi(argv[3]);
  const int nnodes = (nx + 1) * (ny + 1), ntri = 2 * nx * ny, nbc = 2 * (nx + ny);
  double *xy = (double *)malloc(sizeof(double) * 2 * nnodes);
  int *cptr = (int *)malloc(sizeof(int) * (ntri + 1)), *cnode = (int *)malloc(sizeof(int) * 3 * ntri);
  int *bcell = (int *)malloc(sizeof(int) * nbc);
  const double hx = 20.0 / nx, hy = 10.0 / ny;
  for (int j = 0; j <= ny; j++)
    for (int i = 0; i <= nx; i++) {
      double x = i * hx, y = j * hy;
      if (i > 0 && i < nx && j > 0 && j < ny) {  /* a smooth, deterministic distortion of the interior nodes */
        x += 0.15 * hx * sin(1.3 * i + 0.7 * j);
        y += 0.15 * hy * cos(0.9 * i - 1.1 * j);
      }
      xy[2 * (j * (nx + 1) + i)] = x;
      xy[2 * (j * (nx + 1) + i) + 1] = y;
    }
  /* two counter-clockwise triangles per quad along (i,j)->(i+1,j+1); the diagonal is flipped in the lower-right and
   * upper-left corner quads so that no triangle owns two boundary edges (src/residual.f90:112,125) */
  for (int j = 0; j < ny; j++)
    for (int i = 0; i < nx; i++) {
      const int q = j * nx + i, n00 = j * (nx + 1) + i, n10 = n00 + 1, n01 = n00 + nx + 1, n11 = n01 + 1;
      const int flip = (i == nx - 1 && j == 0) || (i == 0 && j == ny - 1);
      int *lo = cnode + 6 * q, *hi = lo + 3;
      if (!flip) { lo[0] = n00; lo[1] = n10; lo[2] = n11; hi[0] = n00; hi[1] = n11; hi[2] = n01; }
      else       { lo[0] = n00; lo[1] = n10; lo[2] = n01; hi[0] = n10; hi[1] = n11; hi[2] = n01; }
    }
  for (int c = 0; c <= ntri; c++) cptr[c] = 3 * c;
  /* boundary cells in edge-walk order: bottom (lower triangles), right, top (upper triangles), left */
  int nb = 0;
  for (int i = 0; i < nx; i++) bcell[nb++] = 2 * (0 * nx + i);
  for (int j = 0; j < ny; j++) { const int q = j * nx + nx - 1; bcell[nb++] = (j == 0) ? 2 * q + 1 : 2 * q; }
  for (int i = nx - 1; i >= 0; i--) bcell[nb++] = 2 * ((ny - 1) * nx + i) + 1;
  for (int j = ny - 1; j >= 0; j--) { const int q = j * nx; bcell[nb++] = (j == ny - 1) ? 2 * q : 2 * q + 1; }

  fvs2d_config c;
  memset(&c, 0, sizeof c);
  c.gamma = 1.4; c.dt = 0.005; c.cfl_user = 1.25; c.umuscl_cst = 0.0; c.lsq_pow = 0.0;
  c.grad_method = 1; c.lsq_stencil = 0; c.limiter = 0; c.recon = 2; c.flux = 1;
  c.rk_nstages = 4; c.rk_order = 4; c.ssprk = 0; c.steady = 0; c.lvortex = 1; c.ntstart = 1;
  c.pvar_inf[0] = 1.0; c.pvar_inf[1] = 0.8; c.pvar_inf[2] = 0.0; c.pvar_inf[3] = 1.0 / 1.4;
  c.vortex_pos[0] = 5.0; c.vortex_pos[1] = 5.0; c.vortex_kappa = 1.0;
  c.vortex_inf[0] = 1.0; c.vortex_inf[1] = 0.2; c.vortex_inf[2] = 0.0; c.vortex_inf[3] = 1.0;
  c.ngpus = 1;

  CHECK(fvs2d_gpu_init(&c, 0));
  const int b_ncells[1] = {nbc}, b_type[1] = {FVS2D_BC_DIRICHLET};
  CHECK(fvs2d_gpu_set_mesh(nnodes, ntri, 0, xy, cptr, cnode, 1, b_ncells, b_type, bcell));
  CHECK(fvs2d_gpu_initialize_solution());
  double *res = (double *)malloc(sizeof(double) * 4 * nsteps), *verr = (double *)malloc(sizeof(double) * 14 * nsteps);
  double *vxy = (double *)malloc(sizeof(double) * 2 * nsteps);
  CHECK(fvs2d_gpu_time_integration(0.0, nsteps, res, verr, vxy));
  double *q = (double *)malloc(sizeof(double) * 4 * ntri), *r = (double *)malloc(sizeof(double) * 4 * ntri);
  CHECK(fvs2d_gpu_get_state(q));
  CHECK(fvs2d_gpu_compute_residual(nsteps * c.dt, r, NULL));
  int sizes[10];
  CHECK(fvs2d_gpu_sizes(sizes));
  FILE *f = fopen(argv[4], "wb");
  if (!f) return 3;
  fwrite(q, sizeof(double), 4 * (size_t)ntri, f);
  fwrite(r, sizeof(double), 4 * (size_t)ntri, f);
  fclose(f);
  f = fopen(argv[5], "w");
  if (!f) return 3;
  for (int s = 0; s < nsteps; s++) fprintf(f, "%d %.17e %.17e %.17e %.17e %.17e\n", s + 1, res[4 * s], res[4 * s + 1], res[4 * s + 2], res[4 * s + 3], verr[14 * s + 3]);
  fclose(f);
  printf("drive_c_abi: nnodes %d ncells %d nedges %d (%d boundary) steps %d ok\n", sizes[0], sizes[1], sizes[2], sizes[4], nsteps);
  CHECK(fvs2d_gpu_finalize());
  return 0;
}
