"""Independent numpy restatement of fvs2d's residual + Runge-Kutta path (TEST INFRASTRUCTURE, fixtures only).

Written from the Fortran sources under /root/reference/src alone -- NOT from oracle/fvs2d_oracle.c and not from the
CUDA kernels -- so that agreement between this file and the C oracle is evidence about the oracle, not a tautology.
Where a mathematically identical but structurally different formulation exists it is used on purpose:

  mesh          edges found by sorting node pairs (the reference walks node->cell lists, src/grid_procs.f90:330-366);
                edge numbering therefore differs from the reference's, c1/c2 orientation does not (c1 = lower cell id,
                n1->n2 in c1's counter-clockwise node order, :413-485), nor does any cell-based quantity
  GGCB          edge-based Green-Gauss with the distance-weighted face value (src/gradient_ggcb.f90:48-138 builds the
                same thing as per-cell coefficients)
  GGNB          cell->node inverse-distance interpolation followed by the trapezoidal Green-Gauss sum
                (the reference's own alternative formulation grad_ggnb_exp, src/gradient_ggnb.f90:249-285,
                instead of the pre-expanded coefficients of :49-210)
  LSQ           batched pseudo-inverse of the weighted displacement matrix (src/gradient_lsq.f90:70-419 forms and
                inverts the 2x2 normal matrix); the face-neighbour stencil of boundary cells is completed from the 8
                nearest centroids with scipy's cKDTree (the reference uses kdtree2, :85-125)
  limiter       src/gradient_limiter.f90:19-134, vectorised over padded face / stencil arrays
  flux          src/flux_invscid.f90:37-136 (Roe, primitive input, Harten fix), vectorised over edges
  residual      src/residual.f90:23-255, edge scatter with np.add.at
  RK            src/runge_kutta.f90:25-437 (4 integrators, local time step, per-step norms)
  vortex / MMS  src/mms.f90:32-365, src/test.f90:481-519

It is the generator of tests/golden/ref_*.npz (tests/golden/make_ref_fixtures.py) and is imported by nothing else.
The reference stores no numeric goldens and cannot be compiled in this image (no Fortran compiler), so these fixtures
pin the oracle against an independent transcription of the same sources -- they are NOT outputs of the reference.
"""
from __future__ import annotations

import numpy as np
from scipy.spatial import cKDTree

PI = float(np.arccos(-1.0))


# ----------------------------------------------------------------------------------------------- mesh
class RefMesh:
    def __init__(self, node_xy, tri, quad, bndry_type, bndry_cell):
        xy = np.asarray(node_xy, dtype=np.float64)
        tri = np.asarray(tri, dtype=np.int64).reshape(-1, 3)
        quad = np.asarray(quad, dtype=np.int64).reshape(-1, 4)
        self.xn, self.yn = xy[:, 0].copy(), xy[:, 1].copy()
        self.nnodes = xy.shape[0]
        self.ntri, self.nquad = tri.shape[0], quad.shape[0]
        nc = self.nc = self.ntri + self.nquad
        self.nvrt = np.concatenate([np.full(self.ntri, 3), np.full(self.nquad, 4)]).astype(np.int64)
        cn = np.full((nc, 4), -1, dtype=np.int64)
        cn[:self.ntri, :3] = tri
        cn[self.ntri:, :] = quad
        self.cnode = cn
        x, y = self.xn, self.yn
        # centroids (src/grid_procs.f90:183-207): running sum over the nodes in order, then /nvrt
        xc, yc = np.zeros(nc), np.zeros(nc)
        for k in range(4):
            m = self.nvrt > k
            xc[m] = xc[m] + x[cn[m, k]]
            yc[m] = yc[m] + y[cn[m, k]]
        self.xc, self.yc = xc / self.nvrt, yc / self.nvrt

        def tarea(a, b, c):  # src/grid_procs.f90:800-808
            return 0.5 * (x[a] * (y[b] - y[c]) + x[b] * (y[c] - y[a]) + x[c] * (y[a] - y[b]))
        vol = np.zeros(nc)
        t, q = slice(0, self.ntri), slice(self.ntri, nc)
        vol[t] = tarea(cn[t, 0], cn[t, 1], cn[t, 2])
        vol[q] = tarea(cn[q, 0], cn[q, 1], cn[q, 2]) + tarea(cn[q, 0], cn[q, 2], cn[q, 3])
        self.vol = vol

        # half edges: local edge k of a cell joins node k and node k+1 (cyclic)  (src/grid_procs.f90:413-485)
        hc, hk, ha, hb = [], [], [], []
        for k in range(4):
            m = np.nonzero(self.nvrt > k)[0]
            nxt = np.where(k + 1 < self.nvrt[m], k + 1, 0)
            hc.append(m); hk.append(np.full(m.size, k)); ha.append(cn[m, k]); hb.append(cn[m, nxt])
        hc, hk, ha, hb = (np.concatenate(v) for v in (hc, hk, ha, hb))
        key = np.minimum(ha, hb) * np.int64(self.nnodes) + np.maximum(ha, hb)
        order = np.lexsort((hc, key))            # same key adjacent, lower cell first
        key_s = key[order]
        first = np.ones(key_s.size, dtype=bool)
        first[1:] = key_s[1:] != key_s[:-1]
        eid_s = np.cumsum(first) - 1
        ne = self.ne = int(eid_s[-1]) + 1
        i_first = order[first]
        self.ec1 = hc[i_first]
        self.en1, self.en2 = ha[i_first], hb[i_first]
        self.ec2 = np.full(ne, -1, dtype=np.int64)
        second = ~first
        self.ec2[eid_s[second]] = hc[order[second]]
        assert np.bincount(eid_s).max() <= 2
        cedge = np.full((nc, 4), -1, dtype=np.int64)
        cedge[hc[order], hk[order]] = eid_s
        self.cedge = cedge
        sign = np.zeros((nc, 4))
        for k in range(4):
            m = self.nvrt > k
            sign[m, k] = np.where(self.ec1[cedge[m, k]] == np.nonzero(m)[0], 1.0, -1.0)
        self.csign = sign                         # nrmlsign, src/grid_procs.f90:636-658
        dx, dy = x[self.en2] - x[self.en1], y[self.en2] - y[self.en1]
        self.ea = np.sqrt(dx ** 2 + dy ** 2)      # :600-617
        self.ex, self.ey = 0.5 * (x[self.en1] + x[self.en2]), 0.5 * (y[self.en1] + y[self.en2])
        self.enx, self.eny = dy / self.ea, -dx / self.ea
        self.e_intr = np.nonzero(self.ec2 >= 0)[0]
        # neighbour across local edge k (-1: boundary)
        nb = np.full((nc, 4), -1, dtype=np.int64)
        for k in range(4):
            m = np.nonzero(self.nvrt > k)[0]
            e = cedge[m, k]
            nb[m, k] = np.where(self.ec1[e] == m, self.ec2[e], self.ec1[e])
        self.nghbre = nb
        isb = np.zeros(nc, dtype=bool)
        for k in range(4):
            isb |= (self.nvrt > k) & (nb[:, k] < 0)
        self.cell_intr = np.nonzero(~isb)[0]      # :663-687
        self.ncells_bndr = int(isb.sum())
        # boundary edge lists in .bc cell order, local edge order within a cell (:749-783)
        self.btype = list(bndry_type)
        self.bedge, self.bcell = [], []
        for cells in bndry_cell:
            el, cl = [], []
            for ic in np.asarray(cells, dtype=np.int64):
                for k in range(int(self.nvrt[ic])):
                    if nb[ic, k] < 0:
                        el.append(cedge[ic, k]); cl.append(ic)
            self.bedge.append(np.array(el, dtype=np.int64)); self.bcell.append(np.array(cl, dtype=np.int64))
        # node -> cells (CSR, ascending cell id like the reference's fill loop :300-308)
        flat_c = np.repeat(np.arange(nc), 4).reshape(nc, 4)[cn >= 0]
        flat_n = cn[cn >= 0]
        o = np.lexsort((flat_c, flat_n))
        self.n2c = flat_c[o]
        self.n2c_ptr = np.concatenate([[0], np.cumsum(np.bincount(flat_n, minlength=self.nnodes))])
        self.n2c_node = flat_n[o]

    def heff(self):
        return float(np.sqrt(self.vol.sum() / self.nc))


# ----------------------------------------------------------------------------------------------- gradients
class GradGGCB:
    """src/gradient_ggcb.f90: face value = (d1*p_own + d0*p_nb)/(d0+d1), d = distance face centre -> centroid;
    boundary faces take the cell value; grad = sum_faces pf * n_out * a / vol."""

    def __init__(self, m: RefMesh):
        self.m = m
        e = m.e_intr
        c1, c2 = m.ec1[e], m.ec2[e]
        d0 = np.sqrt((m.ex[e] - m.xc[c1]) ** 2 + (m.ey[e] - m.yc[c1]) ** 2)
        d1 = np.sqrt((m.ex[e] - m.xc[c2]) ** 2 + (m.ey[e] - m.yc[c2]) ** 2)
        self.w1, self.w2 = d1 / (d0 + d1), d0 / (d0 + d1)
        self.eb = np.nonzero(m.ec2 < 0)[0]

    def __call__(self, p):
        m = self.m
        e = m.e_intr
        c1, c2 = m.ec1[e], m.ec2[e]
        gx, gy = np.zeros_like(p), np.zeros_like(p)
        pf = self.w1[:, None] * p[c1] + self.w2[:, None] * p[c2]
        fx, fy = pf * (m.enx[e] * m.ea[e])[:, None], pf * (m.eny[e] * m.ea[e])[:, None]
        np.add.at(gx, c1, fx); np.add.at(gy, c1, fy)
        np.subtract.at(gx, c2, fx); np.subtract.at(gy, c2, fy)
        eb = self.eb
        cb = m.ec1[eb]
        np.add.at(gx, cb, p[cb] * (m.enx[eb] * m.ea[eb])[:, None])
        np.add.at(gy, cb, p[cb] * (m.eny[eb] * m.ea[eb])[:, None])
        return gx / m.vol[:, None], gy / m.vol[:, None]


def cell2node_idw(m: RefMesh, f):
    """src/interpolation.f90:62-123 (inverse-distance weights, normalised per node)."""
    d = np.sqrt((m.xc[m.n2c] - m.xn[m.n2c_node]) ** 2 + (m.yc[m.n2c] - m.yn[m.n2c_node]) ** 2)
    w = 1.0 / d
    wsum = np.zeros(m.nnodes)
    np.add.at(wsum, m.n2c_node, w)
    out = np.zeros((m.nnodes,) + f.shape[1:])
    np.add.at(out, m.n2c_node, (w / wsum[m.n2c_node])[(...,) + (None,) * (f.ndim - 1)] * f[m.n2c])
    return out


class GradGGNB:
    """src/gradient_ggnb.f90:249-285 (grad_ggnb_exp): node values by IDW, then sum_edges n a (f1+f2)/2 / vol."""

    def __init__(self, m: RefMesh):
        self.m = m

    def __call__(self, p):
        m = self.m
        fv = cell2node_idw(m, p)
        fe = (fv[m.en1] + fv[m.en2]) / 2.0
        fx, fy = fe * (m.enx * m.ea)[:, None], fe * (m.eny * m.ea)[:, None]
        gx, gy = np.zeros_like(p), np.zeros_like(p)
        np.add.at(gx, m.ec1, fx); np.add.at(gy, m.ec1, fy)
        i = m.e_intr
        np.subtract.at(gx, m.ec2[i], fx[i]); np.subtract.at(gy, m.ec2[i], fy[i])
        return gx / m.vol[:, None], gy / m.vol[:, None]


class GradLSQ:
    """src/gradient_lsq.f90: weighted least squares over the face-neighbour ('fn') or node-neighbour ('nn') stencil."""

    def __init__(self, m: RefMesh, stencil: str, power: float):
        self.m = m
        nc = m.nc
        lists = []
        if stencil == "fn":
            tree, k8 = cKDTree(np.stack([m.xc, m.yc], 1)), min(8, nc)
            for ic in range(nc):
                nb = [int(j) for j in m.nghbre[ic, :m.nvrt[ic]] if j >= 0]
                izb = int(m.nvrt[ic]) - len(nb)
                if izb > 2:
                    raise RuntimeError("setup_fn: #s of edges on the boundary>2!")
                if izb > 0:   # :98-123: complete with the nearest centroids that are neither the cell nor a face neighbour
                    _, idx = tree.query([m.xc[ic], m.yc[ic]], k=k8)
                    for j in np.atleast_1d(idx):
                        if izb == 0:
                            break
                        if j != ic and int(j) not in nb:
                            nb.append(int(j)); izb -= 1
                lists.append(nb)
        else:
            for ic in range(nc):   # :230-270: unique cells around the cell's nodes, ascending, without the cell itself
                s = set()
                for v in m.cnode[ic, :m.nvrt[ic]]:
                    s.update(m.n2c[m.n2c_ptr[v]:m.n2c_ptr[v + 1]].tolist())
                s.discard(ic)
                lists.append(sorted(s))
        S = max(len(v) for v in lists)
        idx = np.zeros((nc, S), dtype=np.int64)
        msk = np.zeros((nc, S), dtype=bool)
        for ic, v in enumerate(lists):
            idx[ic, :len(v)] = v
            msk[ic, :len(v)] = True
            idx[ic, len(v):] = ic
        self.idx, self.msk = idx, msk
        dx, dy = m.xc[idx] - m.xc[:, None], m.yc[idx] - m.yc[:, None]
        dis = np.sqrt(dx ** 2 + dy ** 2)
        with np.errstate(divide="ignore"):
            w = np.where(msk & (dis > 0), 1.0 / np.where(dis > 0, dis, 1.0) ** power, 0.0)
        A = np.stack([w * dx, w * dy], axis=2)            # (nc, S, 2), padded rows are zero
        self.pinvw = np.linalg.pinv(A) * w[:, None, :]    # (nc, 2, S): grad = pinv(A) @ (w * dp)

    def __call__(self, p):
        dp = p[self.idx] - p[:, None, :]                  # (nc, S, 4)
        g = np.einsum("cds,csv->cdv", self.pinvw, dp)
        return g[:, 0, :], g[:, 1, :]

    def verify(self):
        """grad_lsq_verify (:490-529): f = 2x + y must give (2, 1)."""
        m = self.m
        f = (2.0 * m.xc + m.yc)[:, None]
        gx, gy = self(np.repeat(f, 4, axis=1))
        return float(max(np.abs(gx - 2.0).max(), np.abs(gy - 1.0).max()))


# ----------------------------------------------------------------------------------------------- physics
def roe_flux(g, L, R, nx, ny):
    """src/flux_invscid.f90:37-136, arrays of faces."""
    tx, ty = -ny, nx
    rL, uL, vL, pL = L.T
    rR, uR, vR, pR = R.T
    unL, unR = uL * nx + vL * ny, uR * nx + vR * ny
    utL, utR = uL * tx + vL * ty, uR * tx + vR * ty
    aL, aR = np.sqrt(g * pL / rL), np.sqrt(g * pR / rR)
    HL = aL * aL / (g - 1.0) + 0.5 * (uL * uL + vL * vL)
    HR = aR * aR / (g - 1.0) + 0.5 * (uR * uR + vR * vR)
    RT = np.sqrt(rR / rL)
    rho = RT * rL
    u, v, H = (uL + RT * uR) / (1.0 + RT), (vL + RT * vR) / (1.0 + RT), (HL + RT * HR) / (1.0 + RT)
    a = np.sqrt((g - 1.0) * (H - 0.5 * (u * u + v * v)))
    un, ut = u * nx + v * ny, u * tx + v * ty
    drho, dp, dun, dut = rR - rL, pR - pL, unR - unL, utR - utL
    LdU = [(dp - rho * a * dun) / (2.0 * a * a), rho * dut, drho - dp / (a * a), (dp + rho * a * dun) / (2.0 * a * a)]
    ws = [np.abs(un - a), np.abs(un), np.abs(un), np.abs(un + a)]
    d = 1.0 / 5.0
    ws[0] = np.where(ws[0] < d, 0.5 * (ws[0] * ws[0] / d + d), ws[0])
    ws[3] = np.where(ws[3] < d, 0.5 * (ws[3] * ws[3] / d + d), ws[3])
    tke = 0.5 * (u * u + v * v)
    one, zero = np.ones_like(u), np.zeros_like(u)
    Rv = [[one, zero, one, one],
          [u - a * nx, tx, u, u + a * nx],
          [v - a * ny, ty, v, v + a * ny],
          [H - un * a, ut, tke, H + un * a]]
    diss = [sum(ws[j] * LdU[j] * Rv[i][j] for j in range(4)) for i in range(4)]
    fL = [rL * unL, rL * unL * uL + pL * nx, rL * unL * vL + pL * ny, rL * unL * HL]
    fR = [rR * unR, rR * unR * uR + pR * nx, rR * unR * vR + pR * ny, rR * unR * HR]
    flux = np.stack([0.5 * (fL[i] + fR[i] - diss[i]) for i in range(4)], axis=1)
    return flux, 0.5 * (np.abs(un) + a)


def vortex_exact(cfg, t, x, y):
    """src/mms.f90:219-265."""
    g = cfg["gamma"]
    ri, ui, vi, p_i = cfg["vortex_inf"]
    K = cfg["vortex_kappa"]
    Ti = p_i / ri
    dx, dy = x - (cfg["vortex_pos"][0] + ui * t), y - (cfg["vortex_pos"][1] + vi * t)
    r = np.sqrt(dx ** 2 + dy ** 2)
    u = ui - K / (2.0 * PI) * dy * np.exp(0.5 * (1.0 - r ** 2))
    v = vi + K / (2.0 * PI) * dx * np.exp(0.5 * (1.0 - r ** 2))
    temp = Ti - (K / (2.0 * PI)) ** 2 * (g - 1.0) / (2.0 * g) * np.exp(1.0 - r ** 2)
    rho = temp ** (1.0 / (g - 1.0))
    return np.stack([rho, u, v, rho ** g], axis=-1)


MMS_C = [(1.12, 0.15, 3.12 * PI, 2.92 * PI), (1.32, 0.06, 2.09 * PI, 3.12 * PI),
         (1.18, 0.03, 2.15 * PI, 3.32 * PI), (1.62, 0.31, 3.79 * PI, 2.98 * PI)]   # src/mms.f90:80-101


def mms_exact(cfg, x, y, corrected=False):
    """src/mms.f90:124-213 -> (sol[...,4], rhs[...,4]); corrected: r*ux instead of the u*rx of :169."""
    g = cfg["gamma"]
    f, fx, fy = [], [], []
    for a0, as_, ax, ay in MMS_C:
        f.append(a0 + as_ * np.sin(ax * x + ay * y))
        fx.append(ax * as_ * np.cos(ax * x + ay * y))
        fy.append(ay * as_ * np.cos(ax * x + ay * y))
    (r, u, v, p), (rx, ux, vx, px), (ry, uy, vy, py) = f, fx, fy
    rH = g / (g - 1.0) * p + r * u * u / 2.0 + r * v * v / 2.0
    rHx = g / (g - 1.0) * px + rx * (u * u + v * v) / 2.0 + r * (u * ux + v * vx)
    rHy = g / (g - 1.0) * py + ry * (u * u + v * v) / 2.0 + r * (u * uy + v * vy)
    r0 = rx * u + (r * ux if corrected else u * rx) + ry * v + r * vy
    r1 = rx * u * u + 2.0 * r * u * ux + ry * u * v + r * uy * v + r * u * vy + px
    r2 = rx * u * v + r * ux * v + r * u * vx + ry * v * v + 2.0 * r * v * vy + py
    r3 = u * rHx + ux * rH + v * rHy + vy * rH
    return np.stack([r, u, v, p], axis=-1), np.stack([r0, r1, r2, r3], axis=-1)


def limiter_fn(kind, a, b, vol):
    """src/gradient_limiter.f90:103-134; kind 1 venk, 2 barth, 3 albada."""
    if kind == 1:
        eps2 = (5.0 * 2.0 * np.sqrt(vol / PI)) ** 3
        return ((a ** 2 + eps2) + 2.0 * b * a) / (a ** 2 + 2.0 * b ** 2 + a * b + eps2)
    if kind == 2:
        return np.minimum(1.0, a / b)
    eps2 = (0.3 * 2.0 * np.sqrt(vol / PI)) ** 3
    lim = ((b ** 2 + eps2) * a + (a ** 2 + eps2) * b) / (a ** 2 + b ** 2 + 2.0 * eps2)
    return lim / (b + eps2)


# ----------------------------------------------------------------------------------------------- solver
class RefSolver:
    """cfg keys: gamma dt cfl_user umuscl_cst lsq_pow grad_method(1 ggcb,2 ggnb,3 lsq) lsq_stencil('fn'|'nn')
    limiter(0..3) recon(1,2,3) rk_order ssprk steady lvortex ntstart pvar_inf vortex_pos vortex_kappa vortex_inf."""

    def __init__(self, mesh: RefMesh, cfg: dict):
        self.m, self.cfg = mesh, dict(cfg)
        c = self.cfg
        if c["recon"] != 3:
            c["umuscl_cst"] = 0.0                      # src/input.f90:248-254
        if c["ntstart"] == 0:
            c["lvortex"] = False                       # src/input.f90:140
        gm = c["grad_method"]
        self.lsq = None
        if gm == 3 or c["limiter"] > 0:
            self.lsq = GradLSQ(mesh, c["lsq_stencil"], c["lsq_pow"])
        self.grad = GradGGCB(mesh) if gm == 1 else GradGGNB(mesh) if gm == 2 else self.lsq
        dt = c["dt"]
        tab = {1: (3.60897, 2.04, 0.34206, 0.00897), 2: (0.11, 3.92, 1.86, 0.11), 3: (0.65, 2.7, 2.0, 0.65), 4: (1.0, 2.0, 2.0, 1.0)}
        self.rk_coef = list(tab[c["rk_order"]])
        self.h_rk = [dt / 2.0, dt / 2.0, dt, dt / 6.0]
        self.dts = self.dte = None
        if c["ssprk"]:
            assert c["rk_order"] == 2
            self.rk_coef = [1.0, 1.0, 1.0, 1.0]
            self.h_rk = [dt / 3.0, dt / 3.0, dt / 3.0, dt / 4.0]
            self.dts = [0.0, dt / 3.0, dt * 2.0 / 3.0, dt]
            self.dte = [dt / 3.0, dt * 2.0 / 3.0, dt, dt]
        self.cvar = np.zeros((mesh.nc, 4))
        self.dt_local = np.full(mesh.nc, dt)

    # -- src/data_solution.f90:72-106
    def cvar2pvar(self, q):
        g = self.cfg["gamma"]
        r, u, v = q[:, 0], q[:, 1] / q[:, 0], q[:, 2] / q[:, 0]
        return np.stack([r, u, v, (g - 1.0) * (q[:, 3] - 0.5 * r * (u ** 2 + v ** 2))], axis=1)

    def pvar2cvar(self, p):
        g = self.cfg["gamma"]
        return np.stack([p[:, 0], p[:, 0] * p[:, 1], p[:, 0] * p[:, 2],
                         p[:, 3] / (g - 1.0) + 0.5 * p[:, 0] * (p[:, 1] ** 2 + p[:, 2] ** 2)], axis=1)

    def initialize_solution(self):
        """src/initialize.f90:19-90 for ntstart <= 1."""
        c, m = self.cfg, self.m
        if c["ntstart"] == 1 and c["lvortex"]:
            p = vortex_exact(c, (c["ntstart"] - 1) * c["dt"], m.xc, m.yc)
        elif c["ntstart"] == 1:
            p = np.tile(np.asarray(c["pvar_inf"], dtype=np.float64), (m.nc, 1))
        else:
            p, _ = mms_exact(c, m.xc, m.yc)
        self.cvar = self.pvar2cvar(p)

    def limiter(self, p, gx, gy):
        """src/gradient_limiter.f90:19-97."""
        c, m = self.cfg, self.m
        if c["recon"] == 1:
            return np.zeros(m.nc)
        if c["limiter"] == 0:
            return np.ones(m.nc)
        idx, msk = self.lsq.idx, self.lsq.msk
        pn = p[idx]
        pmin = np.minimum(p, np.where(msk[:, :, None], pn, np.inf).min(axis=1))
        pmax = np.maximum(p, np.where(msk[:, :, None], pn, -np.inf).max(axis=1))
        phi = np.ones((m.nc, 4))
        for k in range(4):
            live = m.nvrt > k
            e = np.where(live, m.cedge[:, k], 0)
            pf = p + (m.ex[e] - m.xc)[:, None] * gx + (m.ey[e] - m.yc)[:, None] * gy
            diff = pf - p
            with np.errstate(divide="ignore", invalid="ignore"):
                fp = limiter_fn(c["limiter"], pmax - p, diff, m.vol[:, None])
                fm = limiter_fn(c["limiter"], pmin - p, diff, m.vol[:, None])
            f = np.where(diff > 0.0, fp, np.where(diff < 0.0, fm, 1.0))
            phi = np.where(live[:, None], np.minimum(phi, f), phi)
        return np.minimum(1.0, phi).min(axis=1)

    def compute_residual(self, time):
        """src/residual.f90:23-177 -> resid (= -R/vol), ws_nrml; keeps pvar, grad, phi."""
        c, m = self.cfg, self.m
        g, kap = c["gamma"], c["umuscl_cst"]
        p = self.cvar2pvar(self.cvar)
        if c["recon"] == 1:
            gx, gy = np.zeros_like(p), np.zeros_like(p)     # src/gradient.f90:49
        else:
            gx, gy = self.grad(p)
        phi = self.limiter(p, gx, gy)
        self.pvar, self.gx, self.gy, self.phi = p, gx, gy, phi
        res, ws = np.zeros_like(p), np.zeros(m.nc)
        e = m.e_intr
        cL, cR = m.ec1[e], m.ec2[e]
        gC = p[cR] - p[cL]
        gL = (m.ex[e] - m.xc[cL])[:, None] * gx[cL] + (m.ey[e] - m.yc[cL])[:, None] * gy[cL]
        gR = (m.ex[e] - m.xc[cR])[:, None] * gx[cR] + (m.ey[e] - m.yc[cR])[:, None] * gy[cR]
        pfL = p[cL] + phi[cL][:, None] * (kap / 2.0 * gC + (1.0 - kap) * gL)
        pfR = p[cR] + phi[cR][:, None] * (-kap / 2.0 * gC + (1.0 - kap) * gR)
        fl, wm = roe_flux(g, pfL, pfR, m.enx[e], m.eny[e])
        fa = fl * m.ea[e][:, None]
        np.add.at(res, cL, fa); np.subtract.at(res, cR, fa)
        np.add.at(ws, cL, wm * m.ea[e]); np.add.at(ws, cR, wm * m.ea[e])
        for bt, be, bc in zip(m.btype, m.bedge, m.bcell):
            if be.size == 0:
                continue
            nx, ny = m.enx[be], m.eny[be]
            pfL = p[bc] + phi[bc][:, None] * ((m.ex[be] - m.xc[bc])[:, None] * gx[bc] + (m.ey[be] - m.yc[bc])[:, None] * gy[bc])
            if bt == "freestream":
                pfR = np.tile(np.asarray(c["pvar_inf"], dtype=np.float64), (be.size, 1))
            elif bt == "slip_wall":
                un = pfL[:, 1] * nx + pfL[:, 2] * ny
                pfR = pfL.copy()
                pfR[:, 1] = pfL[:, 1] - 2.0 * un * nx
                pfR[:, 2] = pfL[:, 2] - 2.0 * un * ny
            elif bt == "dirichlet":
                pfR = vortex_exact(c, time, m.ex[be], m.ey[be]) if c["lvortex"] else mms_exact(c, m.ex[be], m.ey[be])[0]
            else:
                raise RuntimeError(f"Boundary condition={bt}  not implemented")
            fl, wm = roe_flux(g, pfL, pfR, nx, ny)
            np.add.at(res, bc, fl * m.ea[be][:, None])
            np.add.at(ws, bc, wm * m.ea[be])
        self.resid = -res / m.vol[:, None]
        self.ws_nrml = ws
        return self.resid

    def time_integration(self, t1, nsub):
        """src/runge_kutta.f90:94-437 -> (res_l2[nsub,4], vortex_err[nsub,14] | None, vortex_xy[nsub,2] | None)."""
        c, m = self.cfg, self.m
        dt = c["dt"]
        res_l2 = np.zeros((nsub, 4))
        ve = np.zeros((nsub, 14)) if c["lvortex"] else None
        vxy = np.zeros((nsub, 2)) if c["lvortex"] else None
        for istep in range(nsub):
            told = t1 + float(istep) * dt
            if c["ssprk"]:
                tstart = [told + s for s in self.dts]
                tend = told + self.dte[3]
            else:
                tstart = [told, told + 0.5 * dt, told + 0.5 * dt, told + dt]
                tend = told + dt
            f = np.zeros_like(self.cvar)
            q0 = self.cvar.copy()
            for rk in range(4):
                R = self.compute_residual(tstart[rk])
                if c["steady"] and rk == 0:
                    self.dt_local = c["cfl_user"] * m.vol / (0.5 * self.ws_nrml)      # :424-437
                if c["ssprk"]:
                    if c["steady"]:
                        cst = 1.0 / 4.0 if rk == 3 else 1.0 / 3.0
                        self.cvar = q0 + (self.dt_local * cst)[:, None] * (self.rk_coef[rk] * R + f)
                    else:
                        self.cvar = q0 + self.h_rk[rk] * (self.rk_coef[rk] * R + f)
                    f = f + R
                else:
                    f = f + self.rk_coef[rk] * R
                    if c["steady"]:
                        cst = 1.0 if rk == 2 else 1.0 / 6.0 if rk == 3 else 1.0 / 2.0
                        self.cvar = q0 + (self.dt_local * cst)[:, None] * (R if rk < 3 else f)
                    else:
                        self.cvar = q0 + self.h_rk[rk] * (R if rk < 3 else f)
            if c["lvortex"]:
                ve[istep], vxy[istep] = self.vortex_error(tend)
            res_l2[istep] = np.sqrt((np.abs(self.cvar - q0) ** 2).sum(axis=0) / float(m.nc))
        return res_l2, ve, vxy

    def vortex_error(self, time):
        """src/mms.f90:271-365: the 14 columns of log_vortex_err.plt and the centroid of maxloc(erho)."""
        c, m = self.cfg, self.m
        ci = m.cell_intr
        ex = self.pvar2cvar(vortex_exact(c, time, m.xc[ci], m.yc[ci]))
        d = np.abs(self.cvar[ci] - ex)
        n = float(ci.size)
        row = [time]
        for v in range(4):
            row += [d[:, v].max(), d[:, v].sum() / n, np.sqrt((d[:, v] ** 2).sum() / n)]
        row.append(np.sqrt((d ** 2).sum() / n))
        erho = np.zeros(m.nc)
        erho[ci] = d[:, 0]
        k = int(np.argmax(erho))
        return np.array(row), np.array([m.xc[k], m.yc[k]])

    def test_resid(self, corrected=False):
        """src/test.f90:481-519: L2 / Linf of resid + mms_source over the interior cells."""
        m = self.m
        R = self.compute_residual(0.0)
        ci = m.cell_intr
        _, src = mms_exact(self.cfg, m.xc[ci], m.yc[ci], corrected)
        e = R[ci] + src
        return np.sqrt((e ** 2).sum(axis=0) / float(ci.size)), np.abs(e).max(axis=0)
