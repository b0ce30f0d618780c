/*
 * fvs2d_gpu.h -- C-ABI of the B200-native fvs2d hot path (libfvs2d_gpu.so).
 *
 * The reference (shirzadgit/fvs2d, Fortran 90) has no FFI; its seam is two module procedures that
 * work on module globals:
 *     call time_integration(t1, nsubsteps(it))      src/fvs2d.f90:146 -> src/runge_kutta.f90:94
 *     call compute_residual(time)                   src/runge_kutta.f90:154,223,296,377; src/test.f90:494
 * This header is what an ISO_C_BINDING interface module on the Fortran side binds instead (the
 * binding is shown in INTEGRATION.md).  Conventions:
 *   - every function returns 0 on success; non-zero -> fvs2d_gpu_last_error() holds the message the
 *     Fortran side prints before `stop` (the reference's error convention is print + stop);
 *   - all pointers are caller-owned HOST memory, copied during the call; the library never frees them;
 *   - reals are C double (the reference is built with -r8), integers are int32 (default INTEGER);
 *   - cell / node ids are 0-based here (the Fortran shim subtracts 1 while flattening cell(:)%node);
 *   - arrays shaped (4,ncells) in Fortran are 4 contiguous doubles per cell here (same memory);
 *   - not thread-safe; one context per process (one process per GPU).
 * There is no CPU fallback: every entry point that computes fails when no CUDA device is usable.
 */
#ifndef FVS2D_GPU_H
#define FVS2D_GPU_H

#ifdef __cplusplus
extern "C" {
#endif

/* boundary types (strings in <grid>.bc; src/residual.f90:195-217) */
enum { FVS2D_BC_FREESTREAM = 1, FVS2D_BC_SLIP_WALL = 2, FVS2D_BC_SOLID_WALL = 3, FVS2D_BC_DIRICHLET = 4 };

/* Mirrors the module state of src/input.f90, src/data_solution.f90:55-63 and src/mms.f90:55-60,80-101. */
typedef struct fvs2d_config {
  double gamma, dt, cfl_user, umuscl_cst, lsq_pow;
  int grad_method;  /* 1 GGCB, 2 GGNB, 3 LSQ                      src/input.f90:190-215 */
  int lsq_stencil;  /* 0 fn (face neighbours), 1 nn (node neighbours)                    */
  int limiter;      /* 0 none, 1 Venkatakrishnan, 2 Barth-Jespersen, 3 van Albada  :221-238 */
  int recon;        /* 1 upwind-1st, 2 upwind-2nd (kappa forced 0), 3 UMUSCL       :244-263 */
  int flux;         /* 1 Roe                                                      :269-277 */
  int rk_nstages, rk_order, ssprk, steady, lvortex, ntstart;
  double pvar_inf[4];                 /* rho,u,v,p freestream      src/data_solution.f90:55-63 */
  double vortex_pos[2], vortex_kappa, vortex_inf[4];            /* src/mms.f90:55-60 */
  double mms_c[4][4];                 /* (c0,cs,cx,cy) for rho,u,v,p   src/mms.f90:80-101 */
  int ngpus;                          /* informational; the communicator decides (fvs2d_gpu_comm_init) */
} fvs2d_config;

/* ---- life cycle -------------------------------------------------------------------------- */

/* Replaces: input_read + data_solution_init + runge_kutta_init state (src/input.f90:64,
 * src/data_solution.f90:41, src/runge_kutta.f90:25).  Selects the CUDA device `device` (pass -1 for
 * LOCAL_RANK from the environment, else device 0) and validates the scheme combination. */
int fvs2d_gpu_init(const fvs2d_config *cfg, int device);

/* Optional, before fvs2d_gpu_set_mesh, one process per GPU: join an NCCL communicator.  Rank 0 calls
 * fvs2d_gpu_comm_unique_id and the host program broadcasts the 128 bytes (MPI_Bcast on the Fortran
 * side, torch.distributed in bench.py).  The reference itself aborts unless nproc==1
 * (src/fvs2d.f90:35-43); this is the domain-decomposed extension SURVEY section 8(e) asks for. */
int fvs2d_gpu_comm_unique_id(char id[128]);
int fvs2d_gpu_comm_init(int rank, int nranks, const char id[128]);

/* Replaces: grid_data + grid_data_verify + gradient_init (src/grid_procs.f90:170-794,
 * src/gradient.f90:18).  Takes what grid_read / grid_bc_read produce (src/grid_procs.f90:63-164):
 * node coordinates, CSR cell->node (triangles first, counter-clockwise), and the boundary cell lists.
 * The library rebuilds connectivity, geometry and gradient coefficients with the reference's entity
 * numbering, renumbers cells along a Hilbert curve, and uploads SoA / sliced-ELL arrays.
 * With a communicator every rank passes the same global mesh and keeps its own partition. */
int fvs2d_gpu_set_mesh(int nnodes, int ncells_tri, int ncells_quad, const double *node_xy /* 2*nnodes */,
                       const int *cell_ptr /* ncells+1 */, const int *cell_node /* CSR */, int nbndries,
                       const int *bndry_ncells, const int *bndry_type /* FVS2D_BC_* */,
                       const int *bndry_cell /* concatenated */);

/* Optional, after fvs2d_gpu_set_mesh (grad_method 3 only): replace the library's least-squares operator by the caller's --
 * the reference's public table `lsq(:)` (src/gradient_lsq.f90:16-27: %ncells, %cell, %w, %coef(2,:)) flattened to CSR in the
 * ORIGINAL cell numbering with 0-based cell ids: entries ptr[ic] .. ptr[ic+1]-1 belong to cell ic, coef holds coef(1,k),
 * coef(2,k) per entry.  The Fortran host's own coefficients (kd-tree tie-breaks of the boundary stencils included) are then
 * used bit for bit: grad = sum_k coef(:,k) * w(k) * (p(cell(k)) - p(ic)), src/gradient_lsq.f90:393-401.  The stencil is also
 * the limiter's min/max set (src/gradient_limiter.f90:54-58).  The call repeats the partitioning and the device upload of
 * fvs2d_gpu_set_mesh; the state must be set (again) afterwards.  The reference's linear-exactness check
 * (src/gradient_lsq.f90:490-529, 1e-10) is applied to the supplied table. */
int fvs2d_gpu_set_lsq(const int *ptr /* ncells+1 */, const int *cell, const double *w, const double *coef /* 2 per entry */);

/* Replaces: initialize_solution for ntstart<=1 (src/initialize.f90:35-53,79-84): freestream, isentropic
 * vortex at t=(ntstart-1)*dt, or the manufactured solution (ntstart==0).  Restart (ntstart>1) is
 * fvs2d_gpu_set_state with the cvar read from cont.s8. */
int fvs2d_gpu_initialize_solution(void);

/* cvar(4,ncells) in the ORIGINAL cell numbering.  Under a communicator the arrays are still global:
 * set_state reads the entries of the cells this rank owns or ghosts; get_state writes only the
 * entries of the cells this rank owns and leaves the others untouched. */
int fvs2d_gpu_set_state(const double *cvar);
int fvs2d_gpu_get_state(double *cvar);

/* Distributed hosts: the same for the owned cells only, cvar_own(4, ncells_own) in the library's local
 * cell order (fvs2d_gpu_mesh_array("orig_id") lists the original id of every local cell, owned first).
 * No permutation, no global array: a plain contiguous copy per rank; ghost copies are refreshed by a
 * halo exchange inside set_state_local. */
int fvs2d_gpu_set_state_local(const double *cvar_own);
int fvs2d_gpu_get_state_local(double *cvar_own);

/* ---- the hot path ------------------------------------------------------------------------ */

/* Replaces: time_integration(t1, ntimes_sub) (src/runge_kutta.f90:94-418), incl. the per-step
 * residual norms written to log_res.plt (:169-184) and error_isentropic_vortex (src/mms.f90:271-365).
 *   res_l2        4*nsub  : sqrt(sum((q-q0)^2)/ncells) for rho, rho*u, rho*v, rho*E per step, or NULL
 *   vortex_err    14*nsub : the 14 columns of log_vortex_err.plt per step (lvortex only), or NULL
 *   vortex_err_xy 2*nsub  : centroid of the max-density-error cell per step, or NULL
 * State stays resident on the device between calls; nothing but these logs is copied back. */
int fvs2d_gpu_time_integration(double t1, int nsub, double *res_l2, double *vortex_err, double *vortex_err_xy);

/* Replaces: compute_residual(time) (src/residual.f90:23-177).  resid(4,ncells) = -R/vol and
 * ws_nrml(ncells) (may be NULL) in the original numbering. */
int fvs2d_gpu_compute_residual(double time, double *resid, double *ws_nrml);

/* Side products of a residual evaluation (src/data_solution.f90:16-21) for the CURRENT state: pvar(4,nc),
 * grad(4,nc,2) in Fortran order [idim][ic][ivar], phi_lim(nc).  Any pointer may be NULL.  The gradients and the limiter are
 * re-evaluated from the current primitive state by the call (pass A), so the three arrays always belong to one state --
 * right after fvs2d_gpu_compute_residual they are that call's; after fvs2d_gpu_time_integration they belong to the final
 * state (the reference leaves those of the last stage's input state there).  First-order reconstruction: grad = 0, as
 * src/gradient.f90:49 leaves it. */
int fvs2d_gpu_get_aux(double *pvar, double *grad, double *phi_lim);

/* test_resid (src/test.f90:481-519): one compute_residual(0) and the L2/Linf norms of
 * resid+mms_source over interior cells.  corrected!=0 uses r*ux in the continuity source instead of
 * the reference's u*rx (src/mms.f90:169). */
int fvs2d_gpu_test_resid(int corrected, double l2[4], double linf[4]);

/* ---- output path (per save interval) ------------------------------- */

/* Replaces: the numerical part of write_inst_ios (src/io.f90:122-150) = cvar2pvar + interpolate_cell2node
 * (src/interpolation.f90:62-123: linear inverse-distance weights, summed over node%cell in ascending cell id) for
 * every primitive variable v (0 rho, 1 u, 2 v, 3 p) with select[v] != 0 -- the lw_inst flags of fvs2d.input
 * line 17.  fnode holds nselected records of nnodes doubles in variable order, ready for writed
 * (src/ios_unstrc.f90:300-404).  Only the node records cross PCIe instead of cvar(4,ncells).
 * Under a communicator every rank gets its SHARE of every node value -- the weighted sum over the cells it owns, with
 * weights normalised by all cells around the node -- and zeros elsewhere: the node values are the sum of the ranks' arrays
 * (MPI_Allreduce / MPI_Reduce with MPI_SUM on the host side), which differs from the one-rank result only by the summation
 * order at nodes on partition interfaces. */
int fvs2d_gpu_interpolate_cell2node(const int select[4], double *fnode /* nselected * nnodes */);

/* Replaces: the numerical part of write_inst_cp_un (src/io.f90:340-449) for boundary ib (0-based, .bc order): the
 * unlimited gradient of the current primitive state (gradient_cellcntr_1var, src/gradient.f90:74-96) and, per edge i
 * of bndry(ib)%edge, vals[4*i..4*i+3] = x_f, p_w, p_cell, u_n (pressure and normal velocity extrapolated to the edge
 * centre from the edge's cell, the cell pressure).  cp, the |V_n| norms and cl/cd are sums of these in edge order, left
 * to the caller exactly as the reference forms them.  Under a communicator a rank fills the entries of the edges whose cell it
 * owns and leaves the others untouched (zero the array first and sum over the ranks); this needs one boundary edge per
 * listed boundary cell, which the reference's own boundary loop assumes as well (src/residual.f90:112-125). */
int fvs2d_gpu_wall_values(int ib, double *vals /* 4 * nedges(ib) */);

/* ---- queries ----------------------------------------------------------------------------- */

/* sizes[0..9] = nnodes, ncells, nedges, nedges_intr, nedges_bndr, ncells_intr, ncells_bndr,
 *               cells owned by this rank, owned + ghost cells, local edges */
int fvs2d_gpu_sizes(int sizes[10]);
/* scalars[0..5] = heff1, heff2, sum(vol), Green-theorem volume, LSQ verify max error, device HBM bytes in use */
int fvs2d_gpu_scalars(double scalars[6]);
/* Host copies of mesh products in the reference's numbering, for verification against grid_data:
 *   double: xc yc vol ex ey ea enx eny grad_cx grad_cy grad_c0x grad_c0y
 *   int:    en1 en2 ec1 ec2 cedge nghbre cell_intr b_edge b_edge_ptr grad_ptr grad_idx perm
 * and of the device layout (local numbering, sliced ELL; see fvs2d_b200/csrc/layout.hpp):
 *   int: f_off f_nbr f_edge g_off g_idx orig_id loc2new bf_type bf_edge peers send_ptr send_idx
 *        recv_begin recv_count;  double: g_cx g_cy lex ley;  unsigned char: is_intr
 * Several ranks (except the least-squares stencil over face neighbours): the pre-processing is partition-local -- the
 * "reference numbering" arrays above then describe this rank's SUBMESH (its cells plus two rings of node-adjacent cells,
 * ids local to it, in ascending original id); int sub_orig / sub_new_id give the original id and the global Hilbert id
 * of every submesh cell (both empty after a whole-mesh build).
 * Call with out==NULL to get the element count. Returns the count, or -1 for an unknown name. */
long fvs2d_gpu_mesh_array(const char *name, void *out);

/* Host half of fvs2d_gpu_set_mesh only (connectivity, geometry, gradient operator, Hilbert
 * renumbering, partition `rank` of `nranks`, halo plan): touches no CUDA API, so the pre-processing
 * can be verified on a machine without a GPU.  The result is visible through fvs2d_gpu_sizes,
 * fvs2d_gpu_scalars and fvs2d_gpu_mesh_array; every computing entry point still needs
 * fvs2d_gpu_init + fvs2d_gpu_set_mesh. */
int fvs2d_host_build(const fvs2d_config *cfg, int rank, int nranks, int nnodes, int ncells_tri, int ncells_quad,
                     const double *node_xy, const int *cell_ptr, const int *cell_node, int nbndries,
                     const int *bndry_ncells, const int *bndry_type, const int *bndry_cell);

/* Kernel timing of the last fvs2d_gpu_time_integration call, measured with CUDA events on the
 * library's stream: ms[0] = whole call, ms[1] = gradient(+limiter) kernels, ms[2] = flux+residual+RK
 * kernels, ms[3] = reductions/boundary/halo; launches = kernels launched by the call. */
int fvs2d_gpu_last_timing(double ms[4], long *launches);

/* Tuning knobs (kernel variant selection for benchmarking); unknown keys are an error.
 *   "timing"   1: CUDA event pair around every kernel launch (fills ms[1..3] of fvs2d_gpu_last_timing; disables the
 *              step graph); 0 (default): only the whole call is timed
 *   "tile"     2 (default): pass B = persistent shared-memory pipeline k_flux_pipe; 0: direct-gather kernel k_flux_rk
 *              (also chosen automatically when a mesh's tiles do not fit the pipeline's shared memory)
 *   "fuse"     one kernel per Runge-Kutta stage (k_stage_fused: gradients rebuilt in shared memory inside the pass-B
 *              pipeline; on several ranks the halo exchange happens inside the same kernel as stores to peer memory).
 *              Applies to second-order upwind reconstruction without limiter when the gradient tables fit shared memory;
 *              results are bitwise those of the two-pass path.  -1 (default) automatic: used wherever it applies, tiles split
 *              into two launches per stage by shared-memory need (triangle tiles at three CTAs per SM, quadrilateral tiles
 *              at two); 2: a single launch per stage; 0: never (two-pass path, NCCL send/recv on several ranks).
 *              On several ranks the option must be set before fvs2d_gpu_set_mesh (it decides the ghost layers).
 *   "graph"    1 (default): steps 2..nsub of a call replay a captured CUDA graph (one GPU, or the fused path on several:
 *              its time step contains no NCCL call); 0: every step eager
 *   "pair"     two threads per cell in both passes (k_gradient2 / k_flux_rk2: the variable pairs in pass A, the faces in pass B, fluxes
 *              joined by shuffles in face order -- the state is bitwise the one-thread kernels'): -1 (default) automatic = on one
 *              GPU for meshes of up to 512 cells per SM (75 776 on a B200: the reference's shipped examples), where one thread
 *              per cell cannot fill the machine's thread slots and a pass lasts as long as one thread's dependent chain (measured
 *              47.1 -> 38.9 us per step at 7 k cells, 86.0 -> 76.4 us at 65 k cells); 0 never; 1 always
 *   "pdl"      1 (default): on one GPU the kernels of a time step are launched with programmatic stream serialisation -- each one
 *              waits for its predecessor's results at its first instruction (griddepcontrol.wait) and lets its successor be
 *              scheduled at once, so launch latency hides behind the running kernel; 0: plain launches
 *   "overlap"  1 (default): multi-GPU halo exchange on a second stream, overlapped with interior-tile work
 *   "ctas"     resident CTAs per SM of the persistent kernels (0 = as many as the occupancy API reports, at most 3) */
int fvs2d_gpu_set_option(const char *key, int value);

const char *fvs2d_gpu_last_error(void);
int fvs2d_gpu_finalize(void);

#ifdef __cplusplus
}
#endif
#endif /* FVS2D_GPU_H */
