/*
 * fvs2d_oracle.c -- CPU restatement of the fvs2d residual + Runge-Kutta hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for the CUDA path in
 * fvs2d_b200/.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  The product never links or calls it.
 *
 * PARITY UNPINNED: the reference (shirzadgit/fvs2d, Fortran 90 + Intel extensions) ships no
 * numeric golden vectors and cannot be compiled in this image (no Fortran compiler).  The oracle
 * is pinned only through the reference's own analytic self-checks (LSQ linear exactness
 * src/gradient_lsq.f90:490-529, sum-of-volumes src/grid_procs.f90:824-840, boundary counts
 * src/grid_procs.f90:722-728,785-791) -- see tests/test_oracle_pins.py -- and, since round 2, against an independent
 * numpy transcription of the same Fortran sources (tests/golden/ref_numpy.py, fixtures tests/golden/ref_*.npz,
 * tests/test_oracle_vs_ref_numpy.py: fields <= 1e-12, log_res <= 1e-10).  Nothing here is an output of the reference itself.
 *
 * Serial, IEEE fp64, compiled with -ffp-contract=off.  Loop order follows the reference
 * (edge-scatter residual, per-cell stencils).  All indices are 0-based here; "no neighbour /
 * boundary" is -1 where the Fortran uses 0.  Citations are path:line under /root/reference/.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define NVAR 4
/* -DORC_OMP builds the "all host cores" context variant (liboracle_omp.so): cell loops run under OpenMP and the
 * interior-edge scatter, which races in the reference (src/residual.f90:65), becomes an edge-parallel flux pass plus
 * a cell-parallel gather.  Same arithmetic per face, different summation order: NOT the reference algorithm and
 * never used for parity. */
#ifdef ORC_OMP
#include <omp.h>
#define OMP_FOR _Pragma("omp parallel for schedule(static)")
#define OMP_FOR_N4 _Pragma("omp parallel for schedule(static)")
#else
#define OMP_FOR
#define OMP_FOR_N4
#endif
enum { BC_FREESTREAM = 1, BC_SLIP_WALL = 2, BC_SOLID_WALL = 3, BC_DIRICHLET = 4 };

typedef struct {
  double gamma, dt, cfl_user, umuscl_cst, lsq_pow;
  int grad_method;  /* 1 GGCB, 2 GGNB, 3 LSQ              src/input.f90:190-215 */
  int lsq_stencil;  /* 0 fn, 1 nn */
  int limiter;      /* 0 none, 1 venk, 2 barth, 3 albada  src/input.f90:221-238 */
  int recon;        /* 1 upwind1, 2 upwind2, 3 umuscl     src/input.f90:244-263 */
  int flux;         /* 1 Roe */
  int rk_nstages, rk_order, ssprk, steady, lvortex, ntstart;
  double pvar_inf[4];
  double vortex_pos[2], vortex_kappa, vortex_inf[4];
  double mms_c[4][4]; /* (c0,cs,cx,cy) for rho,u,v,p      src/mms.f90:80-101 */
  int ngpus;
} orc_config;

typedef struct {
  /* raw mesh */
  int nnodes, ntri, nquad, ncells, nedges;
  double *xn, *yn;
  int *cptr, *cnode; /* CSR cell -> node */
  /* derived, per cell-slot (index cptr[ic]+k) */
  int *nghbr, *nghbre, *cedge;
  double *nrmlsign;
  double *xc, *yc, *vol;
  int *n2c_ptr, *n2c;
  int *en1, *en2, *ec1, *ec2;
  double *ex, *ey, *ea, *enx, *eny;
  int ncells_intr, ncells_bndr, *cell_intr;
  int nedges_intr, nedges_bndr, *edge_intr;
  int nb, *b_ncells, *b_type, *b_cell_ptr, *b_cell, *b_nedges, *b_edge_ptr, *b_edge;
  double heff1, heff2, vol_sum, vol_green;
  /* config + solution */
  orc_config cfg;
  double *pvar, *cvar, *grad /* [idim][ic][ivar] */, *resid, *phi_lim, *ws_nrml;
  double *dt_local;
  double rk_coef[4], h_rk[4], dts[4], dte[4];
  /* gradient operators */
  double *gg_coef0;   /* 2 per cell (ggcb or ggnb) */
  double *ggcb_coefnb;/* per cell-slot, 2 */
  int *ggcb_ptr;      /* per cell-slot */
  int *ggnb_ptr;      /* ncells+1, into ggnb_coefnb */
  double *ggnb_coefnb, *ggnb_coefedg; /* coefedg per cell-slot, 2 */
  int *lsq_ptr, *lsq_cell;
  double *lsq_w, *lsq_coef; /* coef: 2 per entry */
  int lsq_verified;
  double lsq_verify_err;
  /* mms */
  double *mms_sol, *mms_source, *mms_source_fixed;
  /* timers: grad, limiter, flux, rk (seconds), src/mainparam.f90:15-16 */
  double cput[4];
  double *eflux; /* ORC_OMP only: flux*a (4) and ws*a per edge */
  char err[256];
} orc;

static double wtime(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* ---------------------------------------------------------------------------------------------
 * mesh: src/grid_procs.f90:170-794
 * ------------------------------------------------------------------------------------------- */
static double tri_area(double x1, double x2, double x3, double y1, double y2, double y3) {
  /* src/grid_procs.f90:800-808 */
  return 0.5 * (x1 * (y2 - y3) + x2 * (y3 - y1) + x3 * (y1 - y2));
}

static int nvrt(const orc *m, int ic) { return m->cptr[ic + 1] - m->cptr[ic]; }

/* local edge index (0-based) that corresponds to nghbr slot k: tri_v2e/qud_v2e, src/grid_procs.f90:491-492 */
static int slot2edge(int nv, int k) {
  /* 1-based: tri (2,3,1), quad (3,4,1,2)  ==> edge = slot - 2 wrapped (0-based: (k+nv-2)%nv) */
  return (k + nv - 2) % nv;
}

static int grid_data(orc *m) {
  int nc = m->ncells, nn = m->nnodes;
  int nslots = m->cptr[nc];
  m->xc = calloc(nc, 8); m->yc = calloc(nc, 8); m->vol = calloc(nc, 8);
  /* centroids: src/grid_procs.f90:185-208 */
  for (int ic = 0; ic < nc; ic++) {
    int nv = nvrt(m, ic);
    double xc = 0, yc = 0;
    for (int iv = 0; iv < nv; iv++) {
      int vp = m->cnode[m->cptr[ic] + iv];
      xc = xc + m->xn[vp];
      yc = yc + m->yn[vp];
    }
    m->xc[ic] = xc / (double)nv;
    m->yc[ic] = yc / (double)nv;
  }
  /* volumes: src/grid_procs.f90:214-227 */
  for (int ic = 0; ic < nc; ic++) {
    const int *nd = &m->cnode[m->cptr[ic]];
    double x1 = m->xn[nd[0]], y1 = m->yn[nd[0]];
    double x2 = m->xn[nd[1]], y2 = m->yn[nd[1]];
    double x3 = m->xn[nd[2]], y3 = m->yn[nd[2]];
    if (nvrt(m, ic) == 3) {
      m->vol[ic] = tri_area(x1, x2, x3, y1, y2, y3);
    } else {
      double x4 = m->xn[nd[3]], y4 = m->yn[nd[3]];
      m->vol[ic] = tri_area(x1, x2, x3, y1, y2, y3) + tri_area(x1, x3, x4, y1, y3, y4);
    }
  }
  /* effective lengths: src/grid_procs.f90:233-243 */
  double at = 0;
  for (int ic = 0; ic < nc; ic++) at = at + m->vol[ic];
  m->vol_sum = at;
  m->heff1 = sqrt(at / (double)nc);
  at = 0;
  for (int ic = 0; ic < nc; ic++) at = at + sqrt(m->vol[ic]);
  m->heff2 = at / (double)nc;

  /* node -> cell: src/grid_procs.f90:275-300 (cells in ascending order per node) */
  m->n2c_ptr = calloc(nn + 1, sizeof(int));
  for (int s = 0; s < nslots; s++) m->n2c_ptr[m->cnode[s] + 1]++;
  for (int i = 0; i < nn; i++) m->n2c_ptr[i + 1] += m->n2c_ptr[i];
  m->n2c = malloc(sizeof(int) * nslots);
  int *fill = calloc(nn, sizeof(int));
  for (int ic = 0; ic < nc; ic++)
    for (int s = m->cptr[ic]; s < m->cptr[ic + 1]; s++) {
      int vp = m->cnode[s];
      m->n2c[m->n2c_ptr[vp] + fill[vp]++] = ic;
    }
  free(fill);

  /* edge neighbours: src/grid_procs.f90:330-366 */
  m->nghbr = malloc(sizeof(int) * nslots);
  for (int s = 0; s < nslots; s++) m->nghbr[s] = -1;
  for (int ic = 0; ic < nc; ic++) {
    int nv = nvrt(m, ic);
    const int *nd = &m->cnode[m->cptr[ic]];
    for (int iv = 0; iv < nv; iv++) {
      int vL = nd[(iv + 1) % nv], vR = nd[iv];
      int found = 0, jc = -1, im = 0;
      for (int j = m->n2c_ptr[vR]; j < m->n2c_ptr[vR + 1] && !found; j++) {
        jc = m->n2c[j];
        int nvj = nvrt(m, jc);
        const int *ndj = &m->cnode[m->cptr[jc]];
        for (int ii = 0; ii < nvj; ii++) {
          int v2 = ndj[(ii + nvj - 1) % nvj], v1 = ndj[ii];
          if (v1 == vR && v2 == vL) {
            found = 1;
            im = (ii + 1) % nvj;
            break;
          }
        }
      }
      int in = (iv + 2) % nv;
      if (found) {
        m->nghbr[m->cptr[ic] + in] = jc;
        m->nghbr[m->cptr[jc] + im] = ic;
      } else {
        m->nghbr[m->cptr[ic] + in] = -1;
      }
    }
  }
  /* nghbre: src/grid_procs.f90:372-397 */
  m->nghbre = malloc(sizeof(int) * nslots);
  for (int ic = 0; ic < nc; ic++) {
    int nv = nvrt(m, ic);
    for (int k = 0; k < nv; k++) m->nghbre[m->cptr[ic] + slot2edge(nv, k)] = m->nghbr[m->cptr[ic] + k];
  }
  /* count + build edges: src/grid_procs.f90:403-486.  Local edge e=(v_e,v_e+1), neighbour = nghbre[e];
     created iff neighbour > ic or boundary. */
  int ne = 0;
  for (int ic = 0; ic < nc; ic++)
    for (int s = m->cptr[ic]; s < m->cptr[ic + 1]; s++)
      if (m->nghbr[s] > ic || m->nghbr[s] < 0) ne++;
  m->nedges = ne;
  m->en1 = malloc(sizeof(int) * ne); m->en2 = malloc(sizeof(int) * ne);
  m->ec1 = malloc(sizeof(int) * ne); m->ec2 = malloc(sizeof(int) * ne);
  m->cedge = malloc(sizeof(int) * nslots);
  ne = 0;
  for (int ic = 0; ic < nc; ic++) {
    int nv = nvrt(m, ic);
    const int *nd = &m->cnode[m->cptr[ic]];
    for (int e = 0; e < nv; e++) {
      int jc = m->nghbre[m->cptr[ic] + e];
      if (jc > ic || jc < 0) {
        m->en1[ne] = nd[e];
        m->en2[ne] = nd[(e + 1) % nv];
        m->ec1[ne] = ic;
        m->ec2[ne] = jc;
        m->cedge[m->cptr[ic] + e] = ne;
        ne++;
      } else {
        /* src/grid_procs.f90:499-624: find the slot of jc pointing back to ic, take its edge id */
        int nvj = nvrt(m, jc), got = -1;
        for (int k = 0; k < nvj; k++)
          if (m->nghbr[m->cptr[jc] + k] == ic) {
            got = m->cedge[m->cptr[jc] + slot2edge(nvj, k)];
            break;
          }
        m->cedge[m->cptr[ic] + e] = got;
      }
    }
  }
  /* edge geometry: src/grid_procs.f90:630-647 */
  m->ex = malloc(8 * ne); m->ey = malloc(8 * ne); m->ea = malloc(8 * ne);
  m->enx = malloc(8 * ne); m->eny = malloc(8 * ne);
  for (int i = 0; i < ne; i++) {
    int v1 = m->en1[i], v2 = m->en2[i];
    double dx = m->xn[v2] - m->xn[v1], dy = m->yn[v2] - m->yn[v1];
    m->ea[i] = sqrt(dx * dx + dy * dy);
    m->ex[i] = 0.5 * (m->xn[v1] + m->xn[v2]);
    m->ey[i] = 0.5 * (m->yn[v1] + m->yn[v2]);
    m->enx[i] = dy / m->ea[i];
    m->eny[i] = -dx / m->ea[i];
  }
  /* normal signs: src/grid_procs.f90:672-691 */
  m->nrmlsign = malloc(8 * nslots);
  for (int ic = 0; ic < nc; ic++)
    for (int s = m->cptr[ic]; s < m->cptr[ic + 1]; s++) m->nrmlsign[s] = (m->ec1[m->cedge[s]] != ic) ? -1.0 : 1.0;
  /* interior / boundary cells: src/grid_procs.f90:697-728 */
  m->cell_intr = malloc(sizeof(int) * nc);
  m->ncells_intr = m->ncells_bndr = 0;
  for (int ic = 0; ic < nc; ic++) {
    int im = 1;
    for (int s = m->cptr[ic]; s < m->cptr[ic + 1]; s++)
      if (m->nghbre[s] < 0) im = 0;
    if (im) m->cell_intr[m->ncells_intr++] = ic; else m->ncells_bndr++;
  }
  int sumb = 0;
  for (int ib = 0; ib < m->nb; ib++) sumb += m->b_ncells[ib];
  if (m->ncells_bndr != sumb) {
    snprintf(m->err, sizeof m->err, "#s of boundary cells does not match: topology %d, .bc %d", m->ncells_bndr, sumb);
    return 1;
  }
  /* interior / boundary edges: src/grid_procs.f90:733-763 */
  m->edge_intr = malloc(sizeof(int) * ne);
  m->nedges_intr = m->nedges_bndr = 0;
  for (int ie = 0; ie < ne; ie++) {
    if (m->ec2[ie] < 0) m->nedges_bndr++; else m->edge_intr[m->nedges_intr++] = ie;
  }
  /* boundary edge lists: src/grid_procs.f90:765-791 */
  m->b_nedges = calloc(m->nb, sizeof(int));
  m->b_edge_ptr = calloc(m->nb + 1, sizeof(int));
  m->b_edge = malloc(sizeof(int) * (m->nedges_bndr + 1));
  int nbe = 0;
  for (int ib = 0; ib < m->nb; ib++) {
    m->b_edge_ptr[ib] = nbe;
    for (int i = m->b_cell_ptr[ib]; i < m->b_cell_ptr[ib + 1]; i++) {
      int ic = m->b_cell[i];
      for (int s = m->cptr[ic]; s < m->cptr[ic + 1]; s++) {
        int je = m->cedge[s];
        if (m->ec1[je] == ic && m->ec2[je] < 0) {
          if (nbe >= m->nedges_bndr) { nbe++; continue; }
          m->b_edge[nbe++] = je;
        }
      }
    }
    m->b_nedges[ib] = nbe - m->b_edge_ptr[ib];
  }
  m->b_edge_ptr[m->nb] = nbe;
  if (nbe != m->nedges_bndr) {
    snprintf(m->err, sizeof m->err, "#s of boundary edges/faces does not match: topology %d, lists %d", m->nedges_bndr, nbe);
    return 2;
  }
  /* Green-theorem volume: src/grid_procs.f90:831-840 */
  double vol2 = 0;
  for (int ic = 0; ic < nc; ic++) {
    double v = 0;
    for (int s = m->cptr[ic]; s < m->cptr[ic + 1]; s++) {
      int je = m->cedge[s];
      v = v + m->enx[je] * m->nrmlsign[s] * m->ex[je] * m->ea[je];
    }
    vol2 = vol2 + v;
  }
  m->vol_green = vol2;
  return 0;
}

/* ---------------------------------------------------------------------------------------------
 * gradient set-up
 * ------------------------------------------------------------------------------------------- */
static void ggcb_setup(orc *m) { /* src/gradient_ggcb.f90:48-110 */
  int nc = m->ncells, ns = m->cptr[nc];
  m->gg_coef0 = calloc(2 * nc, 8);
  m->ggcb_coefnb = calloc(2 * ns, 8);
  m->ggcb_ptr = calloc(ns, sizeof(int));
  for (int ic = 0; ic < nc; ic++) {
    double xc = m->xc[ic], yc = m->yc[ic];
    for (int s = m->cptr[ic]; s < m->cptr[ic + 1]; s++) {
      int je = m->cedge[s];
      int ic1 = m->ec1[je], ic2 = m->ec2[je];
      double af = m->ea[je];
      double nxf = m->enx[je] * m->nrmlsign[s], nyf = m->eny[je] * m->nrmlsign[s];
      double xf = m->ex[je], yf = m->ey[je];
      double dx = xf - xc, dy = yf - yc;
      double d0 = sqrt(dx * dx + dy * dy);
      double xc1 = xc, yc1 = yc;
      m->ggcb_ptr[s] = ic;
      if (ic1 >= 0 && ic1 != ic) { xc1 = m->xc[ic1]; yc1 = m->yc[ic1]; m->ggcb_ptr[s] = ic1; }
      else if (ic2 >= 0 && ic2 != ic) { xc1 = m->xc[ic2]; yc1 = m->yc[ic2]; m->ggcb_ptr[s] = ic2; }
      dx = xf - xc1; dy = yf - yc1;
      double d1 = sqrt(dx * dx + dy * dy);
      m->gg_coef0[2 * ic + 0] = m->gg_coef0[2 * ic + 0] + d1 / (d0 + d1) * nxf * af;
      m->gg_coef0[2 * ic + 1] = m->gg_coef0[2 * ic + 1] + d1 / (d0 + d1) * nyf * af;
      m->ggcb_coefnb[2 * s + 0] = d0 / (d0 + d1) * nxf * af;
      m->ggcb_coefnb[2 * s + 1] = d0 / (d0 + d1) * nyf * af;
    }
  }
}

static void ggnb_setup(orc *m) { /* src/gradient_ggnb.f90:49-177 */
  int nc = m->ncells, nn = m->nnodes, ns = m->cptr[nc];
  double *idw = malloc(8 * nn);
  for (int in = 0; in < nn; in++) {
    double idt = 0;
    for (int j = m->n2c_ptr[in]; j < m->n2c_ptr[in + 1]; j++) {
      int ic = m->n2c[j];
      double dx = m->xc[ic] - m->xn[in], dy = m->yc[ic] - m->yn[in];
      idt = idt + 1.0 / sqrt(dx * dx + dy * dy);
    }
    idw[in] = 1.0 / idt;
  }
  m->gg_coef0 = calloc(2 * nc, 8);
  for (int ic = 0; ic < nc; ic++)
    for (int s = m->cptr[ic]; s < m->cptr[ic + 1]; s++) {
      int je = m->cedge[s];
      double af = m->ea[je], nxf = m->enx[je] * m->nrmlsign[s], nyf = m->eny[je] * m->nrmlsign[s];
      int iv1 = m->en1[je], iv2 = m->en2[je];
      double dx = m->xc[ic] - m->xn[iv1], dy = m->yc[ic] - m->yn[iv1];
      double w1 = 1.0 / sqrt(dx * dx + dy * dy);
      dx = m->xc[ic] - m->xn[iv2]; dy = m->yc[ic] - m->yn[iv2];
      double w2 = 1.0 / sqrt(dx * dx + dy * dy);
      double wt1 = idw[iv1], wt2 = idw[iv2];
      m->gg_coef0[2 * ic + 0] = m->gg_coef0[2 * ic + 0] + af * nxf / 2.0 * (w1 * wt1 + w2 * wt2);
      m->gg_coef0[2 * ic + 1] = m->gg_coef0[2 * ic + 1] + af * nyf / 2.0 * (w1 * wt1 + w2 * wt2);
    }
  m->ggnb_ptr = calloc(nc + 1, sizeof(int));
  for (int ic = 0; ic < nc; ic++) {
    int cnt = 0;
    for (int s = m->cptr[ic]; s < m->cptr[ic + 1]; s++) cnt += m->n2c_ptr[m->cnode[s] + 1] - m->n2c_ptr[m->cnode[s]];
    m->ggnb_ptr[ic + 1] = m->ggnb_ptr[ic] + cnt;
  }
  m->ggnb_coefnb = calloc(m->ggnb_ptr[nc], 8);
  for (int ic = 0; ic < nc; ic++) {
    int i = m->ggnb_ptr[ic];
    for (int s = m->cptr[ic]; s < m->cptr[ic + 1]; s++) {
      int iv = m->cnode[s];
      for (int j = m->n2c_ptr[iv]; j < m->n2c_ptr[iv + 1]; j++, i++) {
        int jc = m->n2c[j];
        double dx = m->xc[jc] - m->xn[iv], dy = m->yc[jc] - m->yn[iv];
        double d = sqrt(dx * dx + dy * dy);
        m->ggnb_coefnb[i] = idw[iv] / d;
        if (jc == ic) m->ggnb_coefnb[i] = 0.0;
      }
    }
  }
  m->ggnb_coefedg = calloc(2 * ns, 8);
  for (int ic = 0; ic < nc; ic++)
    for (int s = m->cptr[ic]; s < m->cptr[ic + 1]; s++) {
      int iv = m->cnode[s];
      for (int t = m->cptr[ic]; t < m->cptr[ic + 1]; t++) {
        int je = m->cedge[t];
        for (int j = 0; j < 2; j++) {
          int jv = j == 0 ? m->en1[je] : m->en2[je];
          if (jv == iv) {
            double af = m->ea[je], nxf = m->enx[je] * m->nrmlsign[t], nyf = m->eny[je] * m->nrmlsign[t];
            m->ggnb_coefedg[2 * s + 0] = m->ggnb_coefedg[2 * s + 0] + af * nxf / 2.0;
            m->ggnb_coefedg[2 * s + 1] = m->ggnb_coefedg[2 * s + 1] + af * nyf / 2.0;
          }
        }
      }
    }
  free(idw);
}

typedef struct { double d2; int idx; } knn_t;

/* exact 8 nearest centroids, ascending squared distance, ties by lower index (replaces kdtree2,
   src/gradient_lsq.f90:85,108; tie order is undefined in the reference) */
static void knn8(const orc *m, int ic, knn_t out[8]) {
  for (int k = 0; k < 8; k++) { out[k].d2 = HUGE_VAL; out[k].idx = -1; }
  double x = m->xc[ic], y = m->yc[ic];
  for (int jc = 0; jc < m->ncells; jc++) {
    double dx = m->xc[jc] - x, dy = m->yc[jc] - y;
    double d2 = dx * dx + dy * dy;
    if (d2 < out[7].d2) {
      int k = 7;
      while (k > 0 && out[k - 1].d2 > d2) { out[k] = out[k - 1]; k--; }
      out[k].d2 = d2; out[k].idx = jc;
    }
  }
}

static void lsq_coefficients(orc *m) { /* common tail of setup_fn/setup_nn: src/gradient_lsq.f90:137-203, 281-347 */
  int nc = m->ncells;
  int tot = m->lsq_ptr[nc];
  m->lsq_w = calloc(tot, 8);
  m->lsq_coef = calloc(2 * tot, 8);
  double lsq_p = m->cfg.lsq_pow;
  for (int ic = 0; ic < nc; ic++) {
    double xc = m->xc[ic], yc = m->yc[ic];
    int b = m->lsq_ptr[ic], n = m->lsq_ptr[ic + 1] - b;
    double g[2][2] = {{0, 0}, {0, 0}};
    double *d = malloc(16 * (n > 0 ? n : 1));
    for (int i = 0; i < n; i++) {
      int jc = m->lsq_cell[b + i];
      double xc1 = m->xc[jc], yc1 = m->yc[jc];
      double dis = sqrt((xc1 - xc) * (xc1 - xc) + (yc1 - yc) * (yc1 - yc));
      double w = 0.0;
      if (dis > 0.0) { w = 1.0 / pow(dis, lsq_p); m->lsq_w[b + i] = w; }
      d[2 * i + 0] = w * (xc1 - xc);
      d[2 * i + 1] = w * (yc1 - yc);
    }
    for (int i = 0; i < 2; i++)
      for (int j = 0; j < 2; j++)
        for (int k = 0; k < n; k++) g[i][j] = g[i][j] + d[2 * k + i] * d[2 * k + j];
    double det = g[0][0] * g[1][1] - g[0][1] * g[1][0];
    double gi[2][2];
    gi[0][0] = 1.0 / det * g[1][1];
    gi[1][1] = 1.0 / det * g[0][0];
    gi[0][1] = -1.0 / det * g[0][1];
    gi[1][0] = -1.0 / det * g[1][0];
    for (int i = 0; i < 2; i++)
      for (int k2 = 0; k2 < n; k2++) {
        double acc = 0.0;
        for (int k = 0; k < 2; k++) acc = acc + gi[i][k] * d[2 * k2 + k];
        m->lsq_coef[2 * (b + k2) + i] = acc;
      }
    free(d);
  }
}

static int lsq_setup_fn(orc *m) { /* src/gradient_lsq.f90:70-131 */
  int nc = m->ncells;
  m->lsq_ptr = malloc(sizeof(int) * (nc + 1));
  for (int ic = 0; ic <= nc; ic++) m->lsq_ptr[ic] = m->cptr[ic];
  m->lsq_cell = malloc(sizeof(int) * m->cptr[nc]);
  for (int ic = 0; ic < nc; ic++) {
    int nv = nvrt(m, ic), b = m->cptr[ic];
    int tmpi[4] = {0, 0, 0, 0}, izb = 0;
    for (int in = 0; in < nv; in++) {
      int jc = m->nghbr[b + in];
      m->lsq_cell[b + in] = -1;
      if (jc < 0) tmpi[izb++] = in; else m->lsq_cell[b + in] = jc;
    }
    if (izb > 0 && izb < 3) {
      knn_t near[8];
      knn8(m, ic, near);
      int nt = 0;
      for (int i = 0; i < 8 && nt < izb; i++) {
        int jc = near[i].idx;
        if (jc < 0 || jc == ic) continue;
        int isnb = 0;
        for (int in = 0; in < nv; in++)
          if (m->nghbr[b + in] == jc) isnb = 1;
        if (!isnb) m->lsq_cell[b + tmpi[nt++]] = jc;
      }
      if (nt < izb) { snprintf(m->err, sizeof m->err, "lsq fn: not enough nearest cells for cell %d", ic); return 3; }
    } else if (izb > 2) {
      snprintf(m->err, sizeof m->err, "error in gradient_lsq, sub: setup_fn: #s of edges on the boundary>2!");
      return 4;
    }
  }
  lsq_coefficients(m);
  return 0;
}

static int cmp_int(const void *a, const void *b) { return *(const int *)a - *(const int *)b; }

static int lsq_setup_nn(orc *m) { /* src/gradient_lsq.f90:214-271: sorted unique vertex neighbours */
  int nc = m->ncells;
  m->lsq_ptr = calloc(nc + 1, sizeof(int));
  int cap = 16 * nc, tot = 0;
  m->lsq_cell = malloc(sizeof(int) * cap);
  int tmp[256];
  for (int ic = 0; ic < nc; ic++) {
    int nt = 0;
    for (int s = m->cptr[ic]; s < m->cptr[ic + 1]; s++) {
      int iv = m->cnode[s];
      for (int j = m->n2c_ptr[iv]; j < m->n2c_ptr[iv + 1]; j++)
        if (m->n2c[j] != ic && nt < 256) tmp[nt++] = m->n2c[j];
    }
    qsort(tmp, nt, sizeof(int), cmp_int);
    int nu = 0;
    for (int i = 0; i < nt; i++)
      if (i == 0 || tmp[i] != tmp[i - 1]) tmp[nu++] = tmp[i];
    if (tot + nu > cap) { cap = 2 * cap + nu; m->lsq_cell = realloc(m->lsq_cell, sizeof(int) * cap); }
    memcpy(&m->lsq_cell[tot], tmp, sizeof(int) * nu);
    tot += nu;
    m->lsq_ptr[ic + 1] = tot;
  }
  lsq_coefficients(m);
  return 0;
}

static int lsq_verify(orc *m) { /* src/gradient_lsq.f90:490-529 */
  const double cstx = 2.0, csty = 1.0;
  m->lsq_verified = 1;
  m->lsq_verify_err = 0;
  for (int ic = 0; ic < m->ncells; ic++) {
    double xc = m->xc[ic], yc = m->yc[ic], dfx = 0, dfy = 0;
    for (int i = m->lsq_ptr[ic]; i < m->lsq_ptr[ic + 1]; i++) {
      int jc = m->lsq_cell[i];
      double diff = csty * m->yc[jc] + cstx * m->xc[jc] - (csty * yc + cstx * xc);
      dfx = dfx + m->lsq_coef[2 * i + 0] * diff * m->lsq_w[i];
      dfy = dfy + m->lsq_coef[2 * i + 1] * diff * m->lsq_w[i];
    }
    double e = fmax(fabs(dfx - cstx), fabs(dfy - csty));
    if (e > m->lsq_verify_err) m->lsq_verify_err = e;
    if (fabs(dfx - cstx) > 1.0e-10 || fabs(dfy - csty) > 1.0e-10) m->lsq_verified = 0;
  }
  if (!m->lsq_verified) { snprintf(m->err, sizeof m->err, " LSQ coefficients are not correct"); return 5; }
  return 0;
}

/* ---------------------------------------------------------------------------------------------
 * analytic fields: src/mms.f90
 * ------------------------------------------------------------------------------------------- */
static void isentropic_vortex(const orc_config *c, double t, double x, double y, double pv[4]) {
  /* src/mms.f90:219-265 */
  const double pi = acos(-1.0);
  double K = c->vortex_kappa;
  double rho_inf = c->vortex_inf[0], u_inf = c->vortex_inf[1], v_inf = c->vortex_inf[2], p_inf = c->vortex_inf[3];
  double T_inf = p_inf / rho_inf;
  double xc = c->vortex_pos[0] + u_inf * t, yc = c->vortex_pos[1] + v_inf * t;
  double dx = x - xc, dy = y - yc;
  double r = sqrt(dx * dx + dy * dy);
  double u = u_inf - K / (2.0 * pi) * dy * exp(0.5 * (1.0 - r * r));
  double v = v_inf + K / (2.0 * pi) * dx * exp(0.5 * (1.0 - r * r));
  double kk = K / (2.0 * pi);
  double temp = T_inf - kk * kk * (c->gamma - 1.0) / (2.0 * c->gamma) * exp(1.0 - r * r);
  double rho = pow(temp, 1.0 / (c->gamma - 1.0));
  double p = pow(rho, c->gamma);
  pv[0] = rho; pv[1] = u; pv[2] = v; pv[3] = p;
}

static double manufactured_sol(double a0, double as, double ax, double ay, int nx, int ny, double x, double y) {
  /* src/mms.f90:187-213 */
  if (nx + ny == 0) return a0 + as * sin(ax * x + ay * y);
  double f;
  if ((nx + ny) % 2 == 0) {
    f = -(pow(ax, nx) * pow(ay, ny)) * as * sin(ax * x + ay * y);
    if ((nx + ny) % 4 == 0) f = -f;
  } else {
    f = (pow(ax, nx) * pow(ay, ny)) * as * cos(ax * x + ay * y);
    if ((nx + ny + 1) % 4 == 0) f = -f;
  }
  return f;
}

/* fixed != 0 replaces the reference's continuity typo (src/mms.f90:169 `u*rx`) by `r*ux` */
static void mms_compute_euler2d(const orc_config *c, double xc, double yc, double sol[4], double rhs[4], int fixed) {
  /* src/mms.f90:124-181 */
  const double (*k)[4] = c->mms_c;
  double g = c->gamma;
  double r = manufactured_sol(k[0][0], k[0][1], k[0][2], k[0][3], 0, 0, xc, yc);
  double u = manufactured_sol(k[1][0], k[1][1], k[1][2], k[1][3], 0, 0, xc, yc);
  double v = manufactured_sol(k[2][0], k[2][1], k[2][2], k[2][3], 0, 0, xc, yc);
  double p = manufactured_sol(k[3][0], k[3][1], k[3][2], k[3][3], 0, 0, xc, yc);
  sol[0] = r; sol[1] = u; sol[2] = v; sol[3] = p;
  double rx = manufactured_sol(k[0][0], k[0][1], k[0][2], k[0][3], 1, 0, xc, yc);
  double ux = manufactured_sol(k[1][0], k[1][1], k[1][2], k[1][3], 1, 0, xc, yc);
  double vx = manufactured_sol(k[2][0], k[2][1], k[2][2], k[2][3], 1, 0, xc, yc);
  double px = manufactured_sol(k[3][0], k[3][1], k[3][2], k[3][3], 1, 0, xc, yc);
  double ry = manufactured_sol(k[0][0], k[0][1], k[0][2], k[0][3], 0, 1, xc, yc);
  double uy = manufactured_sol(k[1][0], k[1][1], k[1][2], k[1][3], 0, 1, xc, yc);
  double vy = manufactured_sol(k[2][0], k[2][1], k[2][2], k[2][3], 0, 1, xc, yc);
  double py = manufactured_sol(k[3][0], k[3][1], k[3][2], k[3][3], 0, 1, xc, yc);
  double rH = g / (g - 1.0) * p + r * u * u / 2.0 + r * v * v / 2.0;
  double rHx = g / (g - 1.0) * px + rx * (u * u + v * v) / 2.0 + r * (u * ux + v * vx);
  double rHy = g / (g - 1.0) * py + ry * (u * u + v * v) / 2.0 + r * (u * uy + v * vy);
  if (fixed) rhs[0] = rx * u + r * ux + ry * v + r * vy;
  else       rhs[0] = rx * u + u * rx + ry * v + r * vy; /* verbatim typo, src/mms.f90:169 */
  rhs[1] = rx * u * u + 2.0 * r * u * ux + ry * u * v + r * uy * v + r * u * vy + px;
  rhs[2] = rx * u * v + r * ux * v + r * u * vx + ry * v * v + 2.0 * r * v * vy + py;
  rhs[3] = u * rHx + ux * rH + v * rHy + vy * rH;
}

/* ---------------------------------------------------------------------------------------------
 * state conversions: src/data_solution.f90:72-106
 * ------------------------------------------------------------------------------------------- */
static void cvar2pvar(orc *m) {
  double g = m->cfg.gamma;
  OMP_FOR
  for (int ic = 0; ic < m->ncells; ic++) {
    double *p = &m->pvar[4 * ic];
    const double *q = &m->cvar[4 * ic];
    p[0] = q[0];
    p[1] = q[1] / q[0];
    p[2] = q[2] / q[0];
    p[3] = (g - 1.0) * (q[3] - 0.5 * p[0] * (p[1] * p[1] + p[2] * p[2]));
  }
}
static void pvar2cvar(orc *m) {
  double g = m->cfg.gamma;
  for (int ic = 0; ic < m->ncells; ic++) {
    const double *p = &m->pvar[4 * ic];
    double *q = &m->cvar[4 * ic];
    q[0] = p[0];
    q[1] = p[0] * p[1];
    q[2] = p[0] * p[2];
    q[3] = p[3] / (g - 1.0) + 0.5 * p[0] * (p[1] * p[1] + p[2] * p[2]);
  }
}

/* ---------------------------------------------------------------------------------------------
 * gradient apply
 * ------------------------------------------------------------------------------------------- */
#define GRAD(m, idim, ic, iv) ((m)->grad[((size_t)(idim) * (m)->ncells + (ic)) * 4 + (iv)])

static void grad_ggcb(orc *m) { /* src/gradient_ggcb.f90:116-138 */
  for (int ivar = 0; ivar < NVAR; ivar++)
    OMP_FOR
    for (int ic = 0; ic < m->ncells; ic++) {
      double gx = m->gg_coef0[2 * ic + 0] * m->pvar[4 * ic + ivar];
      double gy = m->gg_coef0[2 * ic + 1] * m->pvar[4 * ic + ivar];
      for (int s = m->cptr[ic]; s < m->cptr[ic + 1]; s++) {
        int ic1 = m->ggcb_ptr[s];
        gx = gx + m->ggcb_coefnb[2 * s + 0] * m->pvar[4 * ic1 + ivar];
        gy = gy + m->ggcb_coefnb[2 * s + 1] * m->pvar[4 * ic1 + ivar];
      }
      GRAD(m, 0, ic, ivar) = gx / m->vol[ic];
      GRAD(m, 1, ic, ivar) = gy / m->vol[ic];
    }
}
static void grad_ggnb(orc *m) { /* src/gradient_ggnb.f90:183-210 */
  for (int ivar = 0; ivar < NVAR; ivar++)
    OMP_FOR
    for (int ic = 0; ic < m->ncells; ic++) {
      double gx = m->gg_coef0[2 * ic + 0] * m->pvar[4 * ic + ivar];
      double gy = m->gg_coef0[2 * ic + 1] * m->pvar[4 * ic + ivar];
      int i = m->ggnb_ptr[ic];
      for (int s = m->cptr[ic]; s < m->cptr[ic + 1]; s++) {
        int iv = m->cnode[s];
        for (int j = m->n2c_ptr[iv]; j < m->n2c_ptr[iv + 1]; j++, i++) {
          int jc = m->n2c[j];
          gx = gx + m->ggnb_coefedg[2 * s + 0] * m->ggnb_coefnb[i] * m->pvar[4 * jc + ivar];
          gy = gy + m->ggnb_coefedg[2 * s + 1] * m->ggnb_coefnb[i] * m->pvar[4 * jc + ivar];
        }
      }
      GRAD(m, 0, ic, ivar) = gx / m->vol[ic];
      GRAD(m, 1, ic, ivar) = gy / m->vol[ic];
    }
}
static void grad_lsq(orc *m) { /* src/gradient_lsq.f90:393-401 (the one loop the reference runs under OpenMP) */
  OMP_FOR
  for (int ic = 0; ic < m->ncells; ic++) {
    double gx[4] = {0, 0, 0, 0}, gy[4] = {0, 0, 0, 0};
    for (int i = m->lsq_ptr[ic]; i < m->lsq_ptr[ic + 1]; i++) {
      int jc = m->lsq_cell[i];
      for (int ivar = 0; ivar < NVAR; ivar++) {
        gx[ivar] = gx[ivar] + m->lsq_coef[2 * i + 0] * (m->pvar[4 * jc + ivar] - m->pvar[4 * ic + ivar]) * m->lsq_w[i];
        gy[ivar] = gy[ivar] + m->lsq_coef[2 * i + 1] * (m->pvar[4 * jc + ivar] - m->pvar[4 * ic + ivar]) * m->lsq_w[i];
      }
    }
    for (int ivar = 0; ivar < NVAR; ivar++) { GRAD(m, 0, ic, ivar) = gx[ivar]; GRAD(m, 1, ic, ivar) = gy[ivar]; }
  }
}
static void compute_gradient_cellcntr(orc *m) { /* src/gradient.f90:44-69 */
  if (m->cfg.recon == 1) return;
  double t1 = wtime();
  if (m->cfg.grad_method == 2) grad_ggnb(m);
  else if (m->cfg.grad_method == 1) grad_ggcb(m);
  else if (m->cfg.grad_method == 3) grad_lsq(m);
  m->cput[0] += wtime() - t1;
}

/* ---------------------------------------------------------------------------------------------
 * limiter: src/gradient_limiter.f90
 * ------------------------------------------------------------------------------------------- */
static double limiter_fn(int type, double a, double b, double vol) { /* :103-134 */
  const double pi = acos(-1.0);
  if (type == 1) {
    double ap = 5.0;
    double h = 2.0 * sqrt(vol / pi);
    double eps2 = (ap * h) * (ap * h) * (ap * h);
    return ((a * a + eps2) + 2.0 * b * a) / (a * a + 2.0 * (b * b) + a * b + eps2);
  } else if (type == 2) {
    return fmin(1.0, a / b);
  } else {
    double h = 2.0 * sqrt(vol / pi);
    double eps2 = (0.3 * h) * (0.3 * h) * (0.3 * h);
    double l = ((b * b + eps2) * a + (a * a + eps2) * b) / (a * a + b * b + 2.0 * eps2);
    return l / (b + eps2);
  }
}
static int compute_gradient_limiter(orc *m) { /* :19-97 */
  int nc = m->ncells;
  if (m->cfg.recon == 1) { for (int ic = 0; ic < nc; ic++) m->phi_lim[ic] = 0.0; return 0; }
  if (m->cfg.limiter == 0) { for (int ic = 0; ic < nc; ic++) m->phi_lim[ic] = 1.0; return 0; }
  if (!m->lsq_ptr) { snprintf(m->err, sizeof m->err, "limiter requires the LSQ stencil (src/gradient_limiter.f90:54)"); return 6; }
  double t1 = wtime();
  for (int ic = 0; ic < nc; ic++) m->phi_lim[ic] = HUGE_VAL;
  for (int ivar = 0; ivar < NVAR; ivar++)
    OMP_FOR
    for (int ic = 0; ic < nc; ic++) {
      double xc = m->xc[ic], yc = m->yc[ic];
      double pc = m->pvar[4 * ic + ivar];
      double pmin = pc, pmax = pc;
      for (int i = m->lsq_ptr[ic]; i < m->lsq_ptr[ic + 1]; i++) {
        int jc = m->lsq_cell[i];
        pmin = fmin(pmin, m->pvar[4 * jc + ivar]);
        pmax = fmax(pmax, m->pvar[4 * jc + ivar]);
      }
      double phi_edge[4] = {100.0, 100.0, 100.0, 100.0};
      int k = 0;
      for (int s = m->cptr[ic]; s < m->cptr[ic + 1]; s++, k++) {
        int je = m->cedge[s];
        double xf = m->ex[je], yf = m->ey[je];
        double pf = pc + (xf - xc) * GRAD(m, 0, ic, ivar) + (yf - yc) * GRAD(m, 1, ic, ivar);
        double dmax = pmax - pc, dmin = pmin - pc, diff = pf - pc;
        if (diff > 0.0) phi_edge[k] = limiter_fn(m->cfg.limiter, dmax, diff, m->vol[ic]);
        else if (diff < 0.0) phi_edge[k] = limiter_fn(m->cfg.limiter, dmin, diff, m->vol[ic]);
        else phi_edge[k] = 1.0;
      }
      double mn = fmin(fmin(phi_edge[0], phi_edge[1]), fmin(phi_edge[2], phi_edge[3]));
      double phi = fmin(1.0, mn);
      if (phi < m->phi_lim[ic]) m->phi_lim[ic] = phi;
    }
  m->cput[1] += wtime() - t1;
  return 0;
}

/* ---------------------------------------------------------------------------------------------
 * Roe flux: src/flux_invscid.f90:37-136
 * ------------------------------------------------------------------------------------------- */
static void flux_invscid_roe(double gamma, const double pvarL[4], const double pvarR[4], double nx, double ny,
                             double flux[4], double *ws_max) {
  double tx = -ny, ty = nx;
  double rhoL = pvarL[0], uL = pvarL[1], vL = pvarL[2], pL = pvarL[3];
  double rhoR = pvarR[0], uR = pvarR[1], vR = pvarR[2], pR = pvarR[3];
  double unL = uL * nx + vL * ny, unR = uR * nx + vR * ny;
  double utL = uL * tx + vL * ty, utR = uR * tx + vR * ty;
  double aL = sqrt(gamma * pL / rhoL), aR = sqrt(gamma * pR / rhoR);
  double kL = 0.5 * (uL * uL + vL * vL), kR = 0.5 * (uR * uR + vR * vR);
  double HL = aL * aL / (gamma - 1.0) + kL, HR = aR * aR / (gamma - 1.0) + kR;
  double RT = sqrt(rhoR / rhoL);
  double rho = RT * rhoL;
  double u = (uL + RT * uR) / (1.0 + RT);
  double v = (vL + RT * vR) / (1.0 + RT);
  double H = (HL + RT * HR) / (1.0 + RT);
  double a = sqrt((gamma - 1.0) * (H - 0.5 * (u * u + v * v)));
  double un = u * nx + v * ny, ut = u * tx + v * ty;
  double drho = rhoR - rhoL, dp = pR - pL, dun = unR - unL, dut = utR - utL;
  double LdU[4], ws[4], Rv[4][4], diss[4];
  LdU[0] = (dp - rho * a * dun) / (2.0 * a * a);
  LdU[1] = rho * dut;
  LdU[2] = drho - dp / (a * a);
  LdU[3] = (dp + rho * a * dun) / (2.0 * a * a);
  ws[0] = fabs(un - a); ws[1] = fabs(un); ws[2] = fabs(un); ws[3] = fabs(un + a);
  double dws = 1.0 / 5.0;
  if (ws[0] < dws) ws[0] = 0.5 * (ws[0] * ws[0] / dws + dws);
  if (ws[3] < dws) ws[3] = 0.5 * (ws[3] * ws[3] / dws + dws);
  double tke = 0.5 * (u * u + v * v);
  Rv[0][0] = 1.0;        Rv[0][1] = 0.0; Rv[0][2] = 1.0; Rv[0][3] = 1.0;
  Rv[1][0] = u - a * nx; Rv[1][1] = tx;  Rv[1][2] = u;   Rv[1][3] = u + a * nx;
  Rv[2][0] = v - a * ny; Rv[2][1] = ty;  Rv[2][2] = v;   Rv[2][3] = v + a * ny;
  Rv[3][0] = H - un * a; Rv[3][1] = ut;  Rv[3][2] = tke; Rv[3][3] = H + un * a;
  for (int i = 0; i < 4; i++) {
    diss[i] = 0.0;
    for (int j = 0; j < 4; j++) diss[i] = diss[i] + ws[j] * LdU[j] * Rv[i][j];
  }
  double fL[4], fR[4];
  fL[0] = rhoL * unL; fL[1] = rhoL * unL * uL + pL * nx; fL[2] = rhoL * unL * vL + pL * ny; fL[3] = rhoL * unL * HL;
  fR[0] = rhoR * unR; fR[1] = rhoR * unR * uR + pR * nx; fR[2] = rhoR * unR * vR + pR * ny; fR[3] = rhoR * unR * HR;
  for (int i = 0; i < 4; i++) flux[i] = 0.5 * (fL[i] + fR[i] - 1.0 * diss[i]);
  *ws_max = 0.5 * (fabs(un) + a);
}

/* ---------------------------------------------------------------------------------------------
 * residual: src/residual.f90:23-255
 * ------------------------------------------------------------------------------------------- */
static int bc_flux(orc *m, double time, double x, double y, double nx, double ny, int type, const double pfL[4], double pfR[4]) {
  /* src/residual.f90:183-255 */
  for (int i = 0; i < 4; i++) pfR[i] = 0.0;
  if (type == BC_FREESTREAM) {
    for (int i = 0; i < 4; i++) pfR[i] = m->cfg.pvar_inf[i];
  } else if (type == BC_SLIP_WALL) {
    for (int i = 0; i < 4; i++) pfR[i] = pfL[i];
    double un = pfL[1] * nx + pfL[2] * ny;
    pfR[1] = pfL[1] - 2.0 * un * nx;
    pfR[2] = pfL[2] - 2.0 * un * ny;
  } else if (type == BC_DIRICHLET) {
    if (m->cfg.lvortex) isentropic_vortex(&m->cfg, time, x, y, pfR);
    else { double rhs[4]; mms_compute_euler2d(&m->cfg, x, y, pfR, rhs, 0); }
  } else {
    snprintf(m->err, sizeof m->err, "Boundary condition type %d not implemented", type);
    return 7;
  }
  return 0;
}

static int compute_residual(orc *m, double time) {
  int nc = m->ncells;
  memset(m->resid, 0, 32 * (size_t)nc);
  memset(m->grad, 0, 64 * (size_t)nc);
  memset(m->phi_lim, 0, 8 * (size_t)nc);
  memset(m->ws_nrml, 0, 8 * (size_t)nc);
  cvar2pvar(m);
  compute_gradient_cellcntr(m);
  int rc = compute_gradient_limiter(m);
  if (rc) return rc;
  double t1 = wtime();
  double kap = m->cfg.umuscl_cst, gam = m->cfg.gamma;
  /* interior edges: src/residual.f90:66-103 */
#ifdef ORC_OMP
  if (!m->eflux) m->eflux = malloc(40 * (size_t)m->nedges);
  OMP_FOR
#endif
  for (int i = 0; i < m->nedges_intr; i++) {
    int ie = m->edge_intr[i];
    double xf = m->ex[ie], yf = m->ey[ie], af = m->ea[ie], nxf = m->enx[ie], nyf = m->eny[ie];
    int icL = m->ec1[ie], icR = m->ec2[ie];
    double xcL = m->xc[icL], ycL = m->yc[icL], xcR = m->xc[icR], ycR = m->yc[icR];
    double pfL[4], pfR[4], flux[4], ws_max;
    for (int v = 0; v < 4; v++) {
      double gradC = m->pvar[4 * icR + v] - m->pvar[4 * icL + v];
      double gradL = (xf - xcL) * GRAD(m, 0, icL, v) + (yf - ycL) * GRAD(m, 1, icL, v);
      double gradR = (xf - xcR) * GRAD(m, 0, icR, v) + (yf - ycR) * GRAD(m, 1, icR, v);
      pfL[v] = m->pvar[4 * icL + v] + m->phi_lim[icL] * (kap / 2.0 * gradC + (1.0 - kap) * gradL);
      pfR[v] = m->pvar[4 * icR + v] + m->phi_lim[icR] * (-kap / 2.0 * gradC + (1.0 - kap) * gradR);
    }
    flux_invscid_roe(gam, pfL, pfR, nxf, nyf, flux, &ws_max);
#ifdef ORC_OMP
    for (int v = 0; v < 4; v++) m->eflux[5 * (size_t)ie + v] = flux[v] * af;
    m->eflux[5 * (size_t)ie + 4] = ws_max * af;
#else
    for (int v = 0; v < 4; v++) {
      m->resid[4 * icL + v] = m->resid[4 * icL + v] + flux[v] * af;
      m->resid[4 * icR + v] = m->resid[4 * icR + v] - flux[v] * af;
    }
    m->ws_nrml[icL] = m->ws_nrml[icL] + ws_max * af;
    m->ws_nrml[icR] = m->ws_nrml[icR] + ws_max * af;
#endif
  }
#ifdef ORC_OMP
  OMP_FOR
  for (int ic = 0; ic < nc; ic++)
    for (int s = m->cptr[ic]; s < m->cptr[ic + 1]; s++) {
      int je = m->cedge[s];
      if (m->ec2[je] < 0) continue; /* boundary edges are added by the serial loop below */
      double sg = m->ec1[je] == ic ? 1.0 : -1.0;
      for (int v = 0; v < 4; v++) m->resid[4 * ic + v] += sg * m->eflux[5 * (size_t)je + v];
      m->ws_nrml[ic] += m->eflux[5 * (size_t)je + 4];
    }
#endif
  /* boundary edges: src/residual.f90:111-157 (left cell = edge%c1) */
  for (int ib = 0; ib < m->nb; ib++)
    for (int i = m->b_edge_ptr[ib]; i < m->b_edge_ptr[ib + 1]; i++) {
      int ie = m->b_edge[i];
      double xf = m->ex[ie], yf = m->ey[ie], af = m->ea[ie], nxf = m->enx[ie], nyf = m->eny[ie];
      int icL = m->ec1[ie];
      double xcL = m->xc[icL], ycL = m->yc[icL];
      double pfL[4], pfR[4], flux[4], ws_max;
      for (int v = 0; v < 4; v++)
        pfL[v] = m->pvar[4 * icL + v] + m->phi_lim[icL] * ((xf - xcL) * GRAD(m, 0, icL, v) + (yf - ycL) * GRAD(m, 1, icL, v));
      rc = bc_flux(m, time, xf, yf, nxf, nyf, m->b_type[ib], pfL, pfR);
      if (rc) return rc;
      flux_invscid_roe(gam, pfL, pfR, nxf, nyf, flux, &ws_max);
      for (int v = 0; v < 4; v++) m->resid[4 * icL + v] = m->resid[4 * icL + v] + flux[v] * af;
      m->ws_nrml[icL] = m->ws_nrml[icL] + ws_max * af;
    }
  m->cput[2] += wtime() - t1;
  /* dQ/dt = -R/vol: src/residual.f90:164-166 */
  OMP_FOR
  for (int ic = 0; ic < nc; ic++)
    for (int v = 0; v < 4; v++) m->resid[4 * ic + v] = -m->resid[4 * ic + v] / m->vol[ic];
  return 0;
}

/* ---------------------------------------------------------------------------------------------
 * Runge-Kutta: src/runge_kutta.f90
 * ------------------------------------------------------------------------------------------- */
static int runge_kutta_init(orc *m) { /* :25-88 */
  const orc_config *c = &m->cfg;
  double dt = c->dt;
  if (c->rk_nstages != 4) { snprintf(m->err, sizeof m->err, "only rk_nstages==4 is coded (src/runge_kutta.f90:36)"); return 8; }
  double ak = 0, bk = 0, ck = 0, dk = 0;
  if (c->rk_order == 1) { ak = 3.60897; bk = 2.04; ck = 0.34206; dk = 0.00897; }
  else if (c->rk_order == 2) { ak = 0.11; bk = 3.92; ck = 1.86; dk = 0.11; }
  else if (c->rk_order == 3) { ak = 0.65; bk = 2.7; ck = 2.0; dk = 0.65; }
  else if (c->rk_order == 4) { ak = 1.0; bk = 2.0; ck = 2.0; dk = 1.0; }
  else { snprintf(m->err, sizeof m->err, "rk_order must be 1..4"); return 9; }
  m->h_rk[0] = dt / 2.0; m->rk_coef[0] = ak;
  m->h_rk[1] = dt / 2.0; m->rk_coef[1] = bk;
  m->h_rk[2] = dt;       m->rk_coef[2] = ck;
  m->h_rk[3] = dt / 6.0; m->rk_coef[3] = dk;
  if (c->rk_nstages == 4 && c->rk_order == 2 && c->ssprk) {
    for (int k = 0; k < 3; k++) { m->h_rk[k] = dt / 3.0; m->rk_coef[k] = 1.0; }
    m->h_rk[3] = dt / 4.0; m->rk_coef[3] = 1.0;
    m->dts[0] = 0.0;            m->dte[0] = dt / 3.0;
    m->dts[1] = dt / 3.0;       m->dte[1] = dt * 2.0 / 3.0;
    m->dts[2] = dt * 2.0 / 3.0; m->dte[2] = dt;
    m->dts[3] = dt;             m->dte[3] = dt;
  } else if (c->ssprk) {
    snprintf(m->err, sizeof m->err, "SSPRK is coded only for 4 stages, order 2 (src/runge_kutta.f90:56)");
    return 10;
  }
  if (c->steady) {
    m->dt_local = malloc(8 * (size_t)m->ncells);
    for (int ic = 0; ic < m->ncells; ic++) m->dt_local[ic] = dt;
  }
  return 0;
}

static void compute_local_time(orc *m) { /* :424-437 */
  OMP_FOR
  for (int ic = 0; ic < m->ncells; ic++) m->dt_local[ic] = m->cfg.cfl_user * m->vol[ic] / (0.5 * m->ws_nrml[ic]);
}

static void error_isentropic_vortex(orc *m, double time, double out[14], double xy[2]) { /* src/mms.f90:271-365 */
  double mx[4] = {0, 0, 0, 0}, l1[4] = {0, 0, 0, 0}, l2[4] = {0, 0, 0, 0}, err_L2 = 0, g = m->cfg.gamma;
  double best = 0.0; int ibest = 0; /* maxloc(erho): first max; erho==0 on boundary cells */
  int first = 1;
#ifdef ORC_OMP
#pragma omp parallel
  {
    double tmx[4] = {0, 0, 0, 0}, tl1[4] = {0, 0, 0, 0}, tl2[4] = {0, 0, 0, 0}, tbest = 0.0;
    int tib = -1;
#pragma omp for schedule(static) nowait
    for (int i = 0; i < m->ncells_intr; i++) {
      int ic = m->cell_intr[i];
      double pv[4], ex[4];
      isentropic_vortex(&m->cfg, time, m->xc[ic], m->yc[ic], pv);
      ex[0] = pv[0]; ex[1] = pv[0] * pv[1]; ex[2] = pv[0] * pv[2];
      ex[3] = pv[3] / (g - 1.0) + 0.5 * pv[0] * (pv[1] * pv[1] + pv[2] * pv[2]);
      for (int v = 0; v < 4; v++) {
        double dv = fabs(m->cvar[4 * ic + v] - ex[v]);
        tmx[v] = fmax(tmx[v], dv); tl1[v] += dv; tl2[v] += dv * dv;
        if (v == 0 && dv > tbest) { tbest = dv; tib = ic; }
      }
    }
#pragma omp critical
    {
      for (int v = 0; v < 4; v++) { mx[v] = fmax(mx[v], tmx[v]); l1[v] += tl1[v]; l2[v] += tl2[v]; err_L2 += tl2[v]; }
      if (tib >= 0 && (tbest > best || (tbest == best && tib < ibest))) { best = tbest; ibest = tib; }
    }
  }
#else
  for (int i = 0; i < m->ncells_intr; i++) {
    int ic = m->cell_intr[i];
    double pv[4];
    isentropic_vortex(&m->cfg, time, m->xc[ic], m->yc[ic], pv);
    double ex[4];
    ex[0] = pv[0]; ex[1] = pv[0] * pv[1]; ex[2] = pv[0] * pv[2];
    ex[3] = pv[3] / (g - 1.0) + 0.5 * pv[0] * (pv[1] * pv[1] + pv[2] * pv[2]);
    double d[4];
    for (int v = 0; v < 4; v++) {
      d[v] = fabs(m->cvar[4 * ic + v] - ex[v]);
      mx[v] = fmax(mx[v], d[v]);
      l1[v] = l1[v] + d[v];
      l2[v] = l2[v] + d[v] * d[v];
    }
    if (d[0] > best) { best = d[0]; ibest = ic; first = 0; }
    err_L2 = err_L2 + d[0] * d[0] + d[1] * d[1] + d[2] * d[2] + d[3] * d[3];
  }
#endif
  (void)first;
  double n = (double)m->ncells_intr;
  out[0] = time;
  for (int v = 0; v < 4; v++) { out[1 + 3 * v] = mx[v]; out[2 + 3 * v] = l1[v] / n; out[3 + 3 * v] = sqrt(l2[v] / n); }
  out[13] = sqrt(err_L2 / n);
  xy[0] = m->xc[ibest]; xy[1] = m->yc[ibest];
}

static void residual_norms(orc *m, const double *cvar0, double out[4]) { /* src/runge_kutta.f90:169-184 */
  for (int v = 0; v < 4; v++) {
    double s = 0;
#ifdef ORC_OMP
#pragma omp parallel for reduction(+ : s)
#endif
    for (int ic = 0; ic < m->ncells; ic++) {
      double d = fabs(m->cvar[4 * ic + v] - cvar0[4 * ic + v]);
      s = s + d * d;
    }
    out[v] = sqrt(s / (double)m->ncells);
  }
}

int orc_time_integration(orc *m, double t1, int nsub, double *res_l2, double *vortex_err, double *vortex_xy) {
  /* src/runge_kutta.f90:94-114 dispatch; :120-191 RK; :197-254 SSPRK; :260-343 RK steady; :349-418 SSPRK steady */
  int nc = m->ncells;
  size_t n4 = 4 * (size_t)nc;
  double *fcvar = malloc(8 * n4), *cvar0 = malloc(8 * n4);
  const orc_config *c = &m->cfg;
  double dt = c->dt;
  int rc = 0;
  for (int istep = 1; istep <= nsub && !rc; istep++) {
    double told = t1 + (double)(istep - 1) * dt;
    double t12 = told + 0.5 * dt, tnew = told + dt, tend = tnew;
    double tstart[4] = {told, t12, t12, tnew};
    memset(fcvar, 0, 8 * n4);
    memcpy(cvar0, m->cvar, 8 * n4);
    for (int rk = 0; rk < 4; rk++) {
      double ts = c->ssprk ? told + m->dts[rk] : tstart[rk];
      if (c->ssprk) tend = told + m->dte[rk];
      rc = compute_residual(m, ts);
      if (rc) break;
      double t0 = wtime();
      if (c->steady && rk == 0) compute_local_time(m);
      if (!c->ssprk && !c->steady) {
        OMP_FOR_N4 for (size_t i = 0; i < n4; i++) fcvar[i] = fcvar[i] + m->rk_coef[rk] * m->resid[i];
        if (rk < 3) OMP_FOR_N4 for (size_t i = 0; i < n4; i++) m->cvar[i] = cvar0[i] + m->h_rk[rk] * m->resid[i];
        else        OMP_FOR_N4 for (size_t i = 0; i < n4; i++) m->cvar[i] = cvar0[i] + m->h_rk[rk] * fcvar[i];
      } else if (c->ssprk && !c->steady) {
        OMP_FOR_N4 for (size_t i = 0; i < n4; i++) m->cvar[i] = cvar0[i] + m->h_rk[rk] * (m->rk_coef[rk] * m->resid[i] + fcvar[i]);
        OMP_FOR_N4 for (size_t i = 0; i < n4; i++) fcvar[i] = fcvar[i] + m->resid[i];
      } else if (!c->ssprk && c->steady) {
        /* intended algorithm of time_integ_RK_steady (the reference double-allocates diff, Appendix C #4) */
        OMP_FOR_N4 for (size_t i = 0; i < n4; i++) fcvar[i] = fcvar[i] + m->rk_coef[rk] * m->resid[i];
        double cst = 1.0 / 2.0;
        if (rk == 2) cst = 1.0;
        if (rk == 3) cst = 1.0 / 6.0;
        if (rk < 3) OMP_FOR for (int ic = 0; ic < nc; ic++) for (int v = 0; v < 4; v++)
            m->cvar[4 * ic + v] = cvar0[4 * ic + v] + m->dt_local[ic] * cst * m->resid[4 * ic + v];
        else OMP_FOR for (int ic = 0; ic < nc; ic++) for (int v = 0; v < 4; v++)
            m->cvar[4 * ic + v] = cvar0[4 * ic + v] + m->dt_local[ic] * cst * fcvar[4 * ic + v];
      } else {
        double cst = 1.0 / 3.0;
        if (rk == 3) cst = 1.0 / 4.0;
        OMP_FOR for (int ic = 0; ic < nc; ic++) for (int v = 0; v < 4; v++)
            m->cvar[4 * ic + v] = cvar0[4 * ic + v] + m->dt_local[ic] * cst * (m->rk_coef[rk] * m->resid[4 * ic + v] + fcvar[4 * ic + v]);
        OMP_FOR_N4 for (size_t i = 0; i < n4; i++) fcvar[i] = fcvar[i] + m->resid[i];
      }
      m->cput[3] += wtime() - t0;
    }
    if (rc) break;
    if (c->lvortex && vortex_err) error_isentropic_vortex(m, tend, &vortex_err[14 * (istep - 1)], &vortex_xy[2 * (istep - 1)]);
    if (res_l2) residual_norms(m, cvar0, &res_l2[4 * (istep - 1)]);
  }
  free(fcvar); free(cvar0);
  return rc;
}

/* ---------------------------------------------------------------------------------------------
 * public API (ctypes)
 * ------------------------------------------------------------------------------------------- */
orc *orc_create(int nnodes, int ntri, int nquad, const double *node_xy, const int *cell_ptr, const int *cell_node,
                int nb, const int *b_ncells, const int *b_type, const int *b_cell) {
  orc *m = calloc(1, sizeof(orc));
  m->nnodes = nnodes; m->ntri = ntri; m->nquad = nquad; m->ncells = ntri + nquad;
  m->xn = malloc(8 * nnodes); m->yn = malloc(8 * nnodes);
  for (int i = 0; i < nnodes; i++) { m->xn[i] = node_xy[2 * i]; m->yn[i] = node_xy[2 * i + 1]; }
  m->cptr = malloc(sizeof(int) * (m->ncells + 1));
  memcpy(m->cptr, cell_ptr, sizeof(int) * (m->ncells + 1));
  m->cnode = malloc(sizeof(int) * cell_ptr[m->ncells]);
  memcpy(m->cnode, cell_node, sizeof(int) * cell_ptr[m->ncells]);
  m->nb = nb;
  m->b_ncells = malloc(sizeof(int) * nb); m->b_type = malloc(sizeof(int) * nb);
  m->b_cell_ptr = calloc(nb + 1, sizeof(int));
  for (int ib = 0; ib < nb; ib++) { m->b_ncells[ib] = b_ncells[ib]; m->b_type[ib] = b_type[ib]; m->b_cell_ptr[ib + 1] = m->b_cell_ptr[ib] + b_ncells[ib]; }
  m->b_cell = malloc(sizeof(int) * (m->b_cell_ptr[nb] + 1));
  memcpy(m->b_cell, b_cell, sizeof(int) * m->b_cell_ptr[nb]);
  int rc = grid_data(m);
  if (rc) fprintf(stderr, "oracle grid_data: %s\n", m->err);
  return m;
}

const char *orc_last_error(orc *m) { return m->err; }

int orc_setup(orc *m, const orc_config *cfg) {
  m->cfg = *cfg;
  size_t nc = m->ncells;
  m->pvar = calloc(4 * nc, 8); m->cvar = calloc(4 * nc, 8); m->grad = calloc(8 * nc, 8);
  m->resid = calloc(4 * nc, 8); m->phi_lim = calloc(nc, 8); m->ws_nrml = calloc(nc, 8);
  int rc = 0;
  if (cfg->grad_method == 1) ggcb_setup(m);
  else if (cfg->grad_method == 2) ggnb_setup(m);
  else if (cfg->grad_method == 3) {
    rc = cfg->lsq_stencil == 0 ? lsq_setup_fn(m) : lsq_setup_nn(m);
    if (!rc) rc = lsq_verify(m);
  } else { snprintf(m->err, sizeof m->err, "check cell-center gradient scheme"); rc = 11; }
  if (rc) return rc;
  if (cfg->flux != 1) { snprintf(m->err, sizeof m->err, "check inviscid flux scheme"); return 12; }
  rc = runge_kutta_init(m);
  if (rc) return rc;
  if (!cfg->lvortex) { /* mms_init: src/mms.f90:66-116 */
    m->mms_sol = malloc(32 * nc); m->mms_source = malloc(32 * nc); m->mms_source_fixed = malloc(32 * nc);
    double tmp[4];
    for (size_t ic = 0; ic < nc; ic++) {
      mms_compute_euler2d(&m->cfg, m->xc[ic], m->yc[ic], &m->mms_sol[4 * ic], &m->mms_source[4 * ic], 0);
      mms_compute_euler2d(&m->cfg, m->xc[ic], m->yc[ic], tmp, &m->mms_source_fixed[4 * ic], 1);
    }
  }
  return 0;
}

int orc_initialize_solution(orc *m) { /* src/initialize.f90:19-90 (restart handled by caller via orc_set_state) */
  const orc_config *c = &m->cfg;
  if (c->ntstart == 1) {
    for (int ic = 0; ic < m->ncells; ic++) {
      if (c->lvortex) isentropic_vortex(c, (double)(c->ntstart - 1) * c->dt, m->xc[ic], m->yc[ic], &m->pvar[4 * ic]);
      else for (int v = 0; v < 4; v++) m->pvar[4 * ic + v] = c->pvar_inf[v];
    }
  } else if (c->ntstart < 1) {
    memcpy(m->pvar, m->mms_sol, 32 * (size_t)m->ncells);
  } else return 0;
  pvar2cvar(m);
  return 0;
}

int orc_compute_residual(orc *m, double time) { return compute_residual(m, time); }

/* test_resid: src/test.f90:481-519.  which: 0 = reference source (typo kept), 1 = corrected source */
int orc_test_resid(orc *m, int which, double l2[4], double linf[4]) {
  if (!m->mms_source) { snprintf(m->err, sizeof m->err, "test_resid needs mms_source (lvortex must be F)"); return 13; }
  int rc = compute_residual(m, 0.0);
  if (rc) return rc;
  const double *src = which ? m->mms_source_fixed : m->mms_source;
  for (int v = 0; v < 4; v++) {
    l2[v] = 0; linf[v] = 0;
    for (int i = 0; i < m->ncells_intr; i++) {
      int ic = m->cell_intr[i];
      double e = m->resid[4 * ic + v] + src[4 * ic + v];
      l2[v] = l2[v] + e * e;
      linf[v] = fmax(linf[v], fabs(e));
    }
    l2[v] = sqrt(l2[v] / (double)m->ncells_intr);
  }
  return 0;
}

void orc_vortex_error(orc *m, double time, double out[14], double xy[2]) { error_isentropic_vortex(m, time, out, xy); }

/* write_inst_ios numerics (src/io.f90:122-150): cvar2pvar, then interpolate_cell2node of primitive variable ivar
 * with the linear inverse-distance weights of cell2node_idw_setup (src/interpolation.f90:62-101: d = |x_c - x_v|,
 * idw = 1/d/sum(1/d)), summed over node%cell in ascending cell id (src/interpolation.f90:107-123). */
int orc_interpolate_cell2node(orc *m, int ivar, double *fv) {
  if (ivar < 0 || ivar >= NVAR) { snprintf(m->err, sizeof m->err, "interpolate_cell2node: bad variable"); return 14; }
  cvar2pvar(m);
  for (int in = 0; in < m->nnodes; in++) {
    double idt = 0.0;
    for (int j = m->n2c_ptr[in]; j < m->n2c_ptr[in + 1]; j++) {
      int ic = m->n2c[j];
      double dx = m->xc[ic] - m->xn[in], dy = m->yc[ic] - m->yn[in];
      idt = idt + 1.0 / sqrt(dx * dx + dy * dy);
    }
    double f = 0.0;
    for (int j = m->n2c_ptr[in]; j < m->n2c_ptr[in + 1]; j++) {
      int ic = m->n2c[j];
      double dx = m->xc[ic] - m->xn[in], dy = m->yc[ic] - m->yn[in];
      double idw = 1.0 / sqrt(dx * dx + dy * dy) / idt;
      f = f + idw * m->pvar[4 * ic + ivar];
    }
    fv[in] = f;
  }
  return 0;
}

/* write_inst_cp_un numerics (src/io.f90:340-449) for boundary ib: the cell values of p, u, v extrapolated to the
 * boundary-edge centres with the UNLIMITED cell gradients of gradient_cellcntr_1var (src/gradient.f90:74-96 -- computed
 * by the selected scheme even for first-order reconstruction).  out[4*i..] = x_f, p_w, p_cell, u_n for edge i of
 * bndry(ib)%edge; the cell is edge%c1 (the reference indexes bndry%cell by the edge counter).
 * pvar is that of the current cvar (the reference calls cvar2pvar in write_inst_ios just before). */
int orc_wall_values(orc *m, int ib, double *out) {
  if (ib < 0 || ib >= m->nb) { snprintf(m->err, sizeof m->err, "wall_values: bad boundary"); return 15; }
  cvar2pvar(m);
  if (m->cfg.grad_method == 2) grad_ggnb(m);
  else if (m->cfg.grad_method == 1) grad_ggcb(m);
  else grad_lsq(m);
  for (int i = m->b_edge_ptr[ib]; i < m->b_edge_ptr[ib + 1]; i++) {
    int ie = m->b_edge[i], ic = m->ec1[ie];
    double dx = m->ex[ie] - m->xc[ic], dy = m->ey[ie] - m->yc[ic];
    double pw = m->pvar[4 * ic + 3] + dx * GRAD(m, 0, ic, 3) + dy * GRAD(m, 1, ic, 3);
    double uw = m->pvar[4 * ic + 1] + dx * GRAD(m, 0, ic, 1) + dy * GRAD(m, 1, ic, 1);
    double vw = m->pvar[4 * ic + 2] + dx * GRAD(m, 0, ic, 2) + dy * GRAD(m, 1, ic, 2);
    double *o = out + 4 * (size_t)(i - m->b_edge_ptr[ib]);
    o[0] = m->ex[ie]; o[1] = pw; o[2] = m->pvar[4 * ic + 3]; o[3] = uw * m->enx[ie] + vw * m->eny[ie];
  }
  return 0;
}

/* sizes: [nnodes,ncells,nedges,nedges_intr,nedges_bndr,ncells_intr,ncells_bndr,nslots,lsq_total,ggnb_total] */
void orc_sizes(orc *m, int out[10]) {
  out[0] = m->nnodes; out[1] = m->ncells; out[2] = m->nedges; out[3] = m->nedges_intr; out[4] = m->nedges_bndr;
  out[5] = m->ncells_intr; out[6] = m->ncells_bndr; out[7] = m->cptr[m->ncells];
  out[8] = m->lsq_ptr ? m->lsq_ptr[m->ncells] : 0; out[9] = m->ggnb_ptr ? m->ggnb_ptr[m->ncells] : 0;
}
void orc_scalars(orc *m, double out[6]) {
  out[0] = m->heff1; out[1] = m->heff2; out[2] = m->vol_sum; out[3] = m->vol_green; out[4] = m->lsq_verify_err; out[5] = (double)m->lsq_verified;
}
void orc_timers(orc *m, double out[4]) { for (int i = 0; i < 4; i++) out[i] = m->cput[i]; }
void orc_reset_timers(orc *m) { for (int i = 0; i < 4; i++) m->cput[i] = 0; }

/* generic array accessor: returns pointer + element count; kind 0 = double, 1 = int */
const void *orc_array(orc *m, const char *name, long *count, int *kind) {
  long nc = m->ncells, ns = m->cptr[m->ncells], ne = m->nedges;
#define D(nm, p, n) if (!strcmp(name, nm)) { *count = (p) ? (n) : 0; *kind = 0; return (p); }
#define I(nm, p, n) if (!strcmp(name, nm)) { *count = (p) ? (n) : 0; *kind = 1; return (p); }
  D("xc", m->xc, nc) D("yc", m->yc, nc) D("vol", m->vol, nc)
  D("ex", m->ex, ne) D("ey", m->ey, ne) D("ea", m->ea, ne) D("enx", m->enx, ne) D("eny", m->eny, ne)
  I("en1", m->en1, ne) I("en2", m->en2, ne) I("ec1", m->ec1, ne) I("ec2", m->ec2, ne)
  I("cptr", m->cptr, nc + 1) I("cnode", m->cnode, ns) I("nghbr", m->nghbr, ns) I("nghbre", m->nghbre, ns) I("cedge", m->cedge, ns)
  D("nrmlsign", m->nrmlsign, ns)
  I("n2c_ptr", m->n2c_ptr, m->nnodes + 1) I("n2c", m->n2c, ns)
  I("cell_intr", m->cell_intr, m->ncells_intr) I("edge_intr", m->edge_intr, m->nedges_intr)
  I("b_edge_ptr", m->b_edge_ptr, m->nb + 1) I("b_edge", m->b_edge, m->nedges_bndr)
  D("pvar", m->pvar, 4 * nc) D("cvar", m->cvar, 4 * nc) D("grad", m->grad, 8 * nc) D("resid", m->resid, 4 * nc)
  D("phi_lim", m->phi_lim, nc) D("ws_nrml", m->ws_nrml, nc) D("dt_local", m->dt_local, nc)
  D("gg_coef0", m->gg_coef0, 2 * nc) D("ggcb_coefnb", m->ggcb_coefnb, 2 * ns) I("ggcb_ptr", m->ggcb_ptr, ns)
  I("ggnb_ptr", m->ggnb_ptr, nc + 1) D("ggnb_coefnb", m->ggnb_coefnb, m->ggnb_ptr ? m->ggnb_ptr[nc] : 0) D("ggnb_coefedg", m->ggnb_coefedg, 2 * ns)
  I("lsq_ptr", m->lsq_ptr, nc + 1) I("lsq_cell", m->lsq_cell, m->lsq_ptr ? m->lsq_ptr[nc] : 0)
  D("lsq_w", m->lsq_w, m->lsq_ptr ? m->lsq_ptr[nc] : 0) D("lsq_coef", m->lsq_coef, m->lsq_ptr ? 2 * m->lsq_ptr[nc] : 0)
  D("mms_sol", m->mms_sol, 4 * nc) D("mms_source", m->mms_source, 4 * nc) D("mms_source_fixed", m->mms_source_fixed, 4 * nc)
  D("rk_coef", m->rk_coef, 4) D("h_rk", m->h_rk, 4) D("dts", m->dts, 4) D("dte", m->dte, 4)
#undef D
#undef I
  *count = -1; *kind = -1;
  return NULL;
}

void orc_set_state(orc *m, const double *cvar) { memcpy(m->cvar, cvar, 32 * (size_t)m->ncells); }

/* pointwise helpers for unit tests */
void orc_roe_flux(double gamma, const double *pL, const double *pR, double nx, double ny, double *flux, double *ws) {
  flux_invscid_roe(gamma, pL, pR, nx, ny, flux, ws);
}
void orc_vortex_point(const orc_config *c, double t, double x, double y, double *pv) { isentropic_vortex(c, t, x, y, pv); }
void orc_mms_point(const orc_config *c, double x, double y, double *sol, double *rhs, int fixed) { mms_compute_euler2d(c, x, y, sol, rhs, fixed); }
double orc_limiter(int type, double a, double b, double vol) { return limiter_fn(type, a, b, vol); }

void orc_destroy(orc *m) {
  if (!m) return;
  void *ptrs[] = {m->xn, m->yn, m->cptr, m->cnode, m->nghbr, m->nghbre, m->cedge, m->nrmlsign, m->xc, m->yc, m->vol,
                  m->n2c_ptr, m->n2c, m->en1, m->en2, m->ec1, m->ec2, m->ex, m->ey, m->ea, m->enx, m->eny, m->cell_intr,
                  m->edge_intr, m->b_ncells, m->b_type, m->b_cell_ptr, m->b_cell, m->b_nedges, m->b_edge_ptr, m->b_edge,
                  m->pvar, m->cvar, m->grad, m->resid, m->phi_lim, m->ws_nrml, m->dt_local, m->gg_coef0, m->ggcb_coefnb,
                  m->ggcb_ptr, m->ggnb_ptr, m->ggnb_coefnb, m->ggnb_coefedg, m->lsq_ptr, m->lsq_cell, m->lsq_w, m->lsq_coef,
                  m->mms_sol, m->mms_source, m->mms_source_fixed};
  for (size_t i = 0; i < sizeof ptrs / sizeof ptrs[0]; i++) free(ptrs[i]);
  free(m);
}
