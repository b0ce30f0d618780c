// fvs2d_main.cpp -- C++ host program over the C-ABI: the drop-in for `program fvs2d` (src/fvs2d.f90).
//
// Reads the reference's inputs from the working directory -- fvs2d.input (src/input.f90:84-120),
// fvs2d.vortex (src/mms.f90:54-61), <base>.grid / <base>.bc (src/grid_procs.f90:63-164), cont.cd/.s8 for a
// restart (src/initialize.f90:56-77) -- runs the save loop of src/fvs2d.f90:131-160 through
// fvs2d_gpu_time_integration, and writes the reference's outputs: log_res.plt (src/runge_kutta.f90:77-80,
// 169-184), log_vortex_err.plt / log_vortex_err_xy.plt (src/mms.f90:283-294,357-363), inst.cd + inst.s4|.s8
// (node-interpolated primitive variables, src/io.f90:60-144, src/interpolation.f90:62-123), save.cd + save.s8
// (src/io.f90:95-113,156-178; ios format of src/ios_unstrc.f90:141-290: text header + big-endian
// direct-access records), log_cp.plt / log_un.plt / log_clcd.plt when a wall boundary exists
// (src/io.f90:340-449), log.fvs2d (src/input.f90:283-409), log.grid, and in MMS mode (ntstart=0) the error_resid.plt
// row of test_resid (src/test.f90:481-519).  All numerics of the hot path happen in libfvs2d_gpu.so; this file is I/O only.
//
//   fvs2d_gpu.exe [device]     run (device = CUDA ordinal, default LOCAL_RANK / 0)
//   fvs2d_gpu.exe --check      inputs + mesh pre-processing only (fvs2d_host_build: no GPU needed): writes log.fvs2d and
//                              log.grid, prints the mesh counts and exits -- grid_data_verify of src/grid_procs.f90:800-878
// Start-up at large meshes: the .grid text is parsed with a single-pass strtod/strtol scanner over the
// whole file, and a binary image <base>.gridbin (magic, counts, node_xy, cell_node) is written next to it and used on
// later runs while it is newer than the text file (FVS2D_NO_GRIDBIN=1 disables both).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <sys/stat.h>
#include <string>
#include <vector>

#include "../../include/fvs2d_gpu.h"

namespace {

[[noreturn]] void stop(const std::string &msg) {
  std::fprintf(stdout, " %s\n", msg.c_str());
  std::exit(1);
}
void check(int rc) {
  if (rc) stop(fvs2d_gpu_last_error());
}

// ---- list-directed reads ----------------------------------------------------------------------
std::vector<std::string> tokens(const std::string &line, size_t n) {
  std::vector<std::string> out;
  std::string cur;
  for (char c : line) {
    if (c == ',' || c == ' ' || c == '\t' || c == '\r') {
      if (!cur.empty()) { out.push_back(cur); cur.clear(); }
      if (out.size() >= n) break;
    } else cur += c;
  }
  if (!cur.empty() && out.size() < n) out.push_back(cur);
  if (out.size() < n) stop("input: expected " + std::to_string(n) + " values on line: " + line);
  return out;
}
double freal(std::string t) {
  for (char &c : t) if (c == 'd' || c == 'D') c = 'e';
  return std::atof(t.c_str());
}
bool flogical(const std::string &t) {
  for (char c : t) { if (c == '.') continue; return c == 'T' || c == 't'; }
  return false;
}

// ---- Fortran edit descriptors -----------------------------------------------------------------
std::string fortran_e(double v, int w, int d) {  // Ew.d: 0.dddddE+ee
  char buf[64];
  if (v == 0.0 || !std::isfinite(v)) {
    std::snprintf(buf, sizeof buf, "0.%0*dE+00", d, 0);
  } else {
    int e = (int)std::floor(std::log10(std::fabs(v))) + 1;
    double m = std::fabs(v) / std::pow(10.0, e);
    double r = std::round(m * std::pow(10.0, d));
    if (r >= std::pow(10.0, d)) { r /= 10.0; e += 1; }
    if (r < std::pow(10.0, d - 1)) { r *= 10.0; e -= 1; }
    std::snprintf(buf, sizeof buf, "%s0.%0*.0fE%c%02d", v < 0 ? "-" : "", d, r, e < 0 ? '-' : '+', std::abs(e));
  }
  std::string s = buf;
  if ((int)s.size() < w) s = std::string(w - s.size(), ' ') + s;
  return s;
}

// ---- ios files (src/ios_unstrc.f90:141-290) ------------------------------------------------------
void write_be(std::ofstream &f, const double *a, size_t n, bool single) {
  std::vector<unsigned char> buf(n * (single ? 4 : 8));
  for (size_t i = 0; i < n; i++) {
    if (single) {
      float x = (float)a[i]; uint32_t u; std::memcpy(&u, &x, 4);
      for (int b = 0; b < 4; b++) buf[4 * i + b] = (unsigned char)(u >> (24 - 8 * b));
    } else {
      uint64_t u; std::memcpy(&u, &a[i], 8);
      for (int b = 0; b < 8; b++) buf[8 * i + b] = (unsigned char)(u >> (56 - 8 * b));
    }
  }
  f.write((const char *)buf.data(), buf.size());
}
void writecd(const std::string &base, int mnodes, int mcells, int mp, int mt, const std::vector<int> &itimes,
             const std::vector<std::string> &params, const std::vector<std::string> &info) {
  std::ofstream f(base + ".cd");
  char buf[256];
  std::snprintf(buf, sizeof buf, "     number of nodes = %d\n     number of cells = %d\n     number of parameters = %5d\n"
                "     number of timesteps  = %5d\n\n     Information about file :   (%3d  info lines )\n", mnodes, mcells, mp, mt, (int)info.size());
  f << buf;
  auto a72 = [&](const std::string &s) { std::string t = s.substr(0, 72); t.resize(72, ' '); f << "   " << t << "\n"; };
  for (auto &s : info) a72(s);
  f << "      Information about parameters :\n";
  for (auto &s : params) a72(s);
  f << "  Numbers of timesteps :\n";
  for (size_t i = 0; i < itimes.size(); i++) {
    std::snprintf(buf, sizeof buf, "  %10d", itimes[i]);
    f << buf;
    if ((i + 1) % 6 == 0 || i + 1 == itimes.size()) f << "\n";
  }
}

struct Input {
  std::string base;
  double rey, mach, aoa, gamma, dt, cfl, umuscl, lsq_pow;
  int ntimes, nsaves, ntstart, grad, limiter, recon, flux, rk_nstages, rk_order;
  bool steady, vortex, lw[4], s8, ssprk, lsq_nn;
  double vpos[2] = {0, 0}, vkap = 0, vinf[4] = {0, 0, 0, 0};
};

Input input_read() {  // src/input.f90:64-277
  std::ifstream f("fvs2d.input");
  if (!f) stop("cannot find \"fvs2d.input\" file!");
  std::vector<std::string> L;
  for (std::string s; std::getline(f, s);) L.push_back(s);
  if (L.size() < 29) stop("fvs2d.input: too short");
  Input in;
  in.base = L[3].substr(0, L[3].find('-'));
  in.base.erase(in.base.find_last_not_of(" \t\r") + 1);
  in.base.erase(0, in.base.find_first_not_of(" \t"));
  in.rey = freal(tokens(L[4], 1)[0]); in.mach = freal(tokens(L[5], 1)[0]); in.aoa = freal(tokens(L[6], 1)[0]);
  in.gamma = freal(tokens(L[7], 1)[0]); in.dt = freal(tokens(L[9], 1)[0]);
  in.ntimes = std::atoi(tokens(L[10], 1)[0].c_str()); in.nsaves = std::atoi(tokens(L[11], 1)[0].c_str());
  in.ntstart = std::atoi(tokens(L[12], 1)[0].c_str());
  auto t = tokens(L[13], 2); in.steady = flogical(t[0]); in.cfl = freal(t[1]);
  in.vortex = flogical(tokens(L[14], 1)[0]);
  t = tokens(L[16], 4); for (int i = 0; i < 4; i++) in.lw[i] = flogical(t[i]);
  std::string m = tokens(L[17], 1)[0];
  if (m == "s4" || m == "S4") in.s8 = false; else if (m == "s8" || m == "S8") in.s8 = true;
  else stop("format for regular output files must be either s4 or s8");
  t = tokens(L[21], 3); in.grad = std::atoi(t[0].c_str()); in.lsq_pow = freal(t[2]);
  std::string nb = t[1].substr(0, 2);
  in.lsq_nn = (nb == "nn" || nb == "NN");
  if (in.grad == 3 && !in.lsq_nn && nb != "fn" && nb != "FN") stop("check Least-Squares gradient scheme in input file");
  in.limiter = std::atoi(tokens(L[22], 1)[0].c_str());
  t = tokens(L[23], 2); in.recon = std::atoi(t[0].c_str()); in.umuscl = freal(t[1]);
  in.flux = std::atoi(tokens(L[24], 1)[0].c_str());
  in.rk_nstages = std::atoi(tokens(L[26], 1)[0].c_str()); in.rk_order = std::atoi(tokens(L[27], 1)[0].c_str());
  in.ssprk = flogical(tokens(L[28], 1)[0]);
  if (in.ntstart == 0) in.vortex = false;  // src/input.f90:140
  if (in.vortex) {                         // src/mms.f90:45-62
    std::ifstream v("fvs2d.vortex");
    if (!v) stop("cannot find \"fvs2d.vortex\" file!");
    std::vector<std::string> V;
    for (std::string s; std::getline(v, s);) V.push_back(s);
    if (V.size() < 6) stop("fvs2d.vortex: too short");
    auto p = tokens(V[0], 2); in.vpos[0] = freal(p[0]); in.vpos[1] = freal(p[1]);
    in.vkap = freal(tokens(V[1], 1)[0]);
    for (int i = 0; i < 4; i++) in.vinf[i] = freal(tokens(V[2 + i], 1)[0]);
  }
  return in;
}

struct Grid {
  int nnodes = 0, ntri = 0, nquad = 0;
  std::vector<double> xy;
  std::vector<int> cptr, cnode, bn, bt, bc;
};

// <base>.gridbin: "FVS2DGRD" | int32 version, nnodes, ntri, nquad | node_xy (2*nnodes f64) | cell_node (int32, 0-based)
const char kGridBinMagic[8] = {'F', 'V', 'S', '2', 'D', 'G', 'R', 'D'};
bool gridbin_read(const std::string &base, Grid &g) {
  struct stat st_txt, st_bin;
  if (stat((base + ".gridbin").c_str(), &st_bin) != 0) return false;
  if (stat((base + ".grid").c_str(), &st_txt) == 0 &&
      (st_txt.st_mtim.tv_sec > st_bin.st_mtim.tv_sec ||
       (st_txt.st_mtim.tv_sec == st_bin.st_mtim.tv_sec && st_txt.st_mtim.tv_nsec > st_bin.st_mtim.tv_nsec)))
    return false;  // the text file changed after the image was written
  FILE *f = std::fopen((base + ".gridbin").c_str(), "rb");
  if (!f) return false;
  char magic[8]; int32_t hdr[4];
  bool ok = std::fread(magic, 1, 8, f) == 8 && std::memcmp(magic, kGridBinMagic, 8) == 0 && std::fread(hdr, 4, 4, f) == 4 && hdr[0] == 1;
  if (ok) {
    g.nnodes = hdr[1]; g.ntri = hdr[2]; g.nquad = hdr[3];
    const size_t nn = 3 * (size_t)g.ntri + 4 * (size_t)g.nquad;
    g.xy.resize(2 * (size_t)g.nnodes); g.cnode.resize(nn);
    ok = std::fread(g.xy.data(), 8, g.xy.size(), f) == g.xy.size() && std::fread(g.cnode.data(), 4, nn, f) == nn;
  }
  std::fclose(f);
  return ok;
}
void gridbin_write(const std::string &base, const Grid &g) {
  FILE *f = std::fopen((base + ".gridbin").c_str(), "wb");
  if (!f) return;  // read-only directory: the image is an optimisation only
  const int32_t hdr[4] = {1, g.nnodes, g.ntri, g.nquad};
  std::fwrite(kGridBinMagic, 1, 8, f); std::fwrite(hdr, 4, 4, f);
  std::fwrite(g.xy.data(), 8, g.xy.size(), f); std::fwrite(g.cnode.data(), 4, g.cnode.size(), f);
  std::fclose(f);
}

Grid grid_read(const std::string &base) {  // src/grid_procs.f90:63-164
  Grid g;
  const bool use_bin = !std::getenv("FVS2D_NO_GRIDBIN");
  if (!(use_bin && gridbin_read(base, g))) {
    // whole file in memory, one pass: list-directed reads accept blanks / commas / newlines as separators and D exponents
    FILE *f = std::fopen((base + ".grid").c_str(), "rb");
    if (!f) stop("cannot find " + base + ".grid file!");
    std::fseek(f, 0, SEEK_END);
    const long sz = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    std::vector<char> buf((size_t)sz + 1);
    if (std::fread(buf.data(), 1, (size_t)sz, f) != (size_t)sz) stop("grid file: short read");
    std::fclose(f);
    buf[sz] = 0;
    char *p = buf.data();
    while (*p && *p != '\n') p++;  // line 1 is a comment
    auto skip = [&]() { while (*p == ' ' || *p == ',' || *p == '\n' || *p == '\r' || *p == '\t') p++; };
    auto next_int = [&](const char *what) -> int {
      skip();
      char *e; const long v = std::strtol(p, &e, 10);
      if (e == p) stop(std::string("grid ") + what);
      p = e;
      return (int)v;
    };
    auto next_real = [&](const char *what) -> double {
      skip();
      char *e; double v = std::strtod(p, &e);
      if (e == p) stop(std::string("grid ") + what);
      if (*e == 'd' || *e == 'D') {  // Fortran D exponent: strtod stopped at the mantissa
        char *e2; const long ex = std::strtol(e + 1, &e2, 10);
        if (e2 != e + 1) { v *= std::pow(10.0, (double)ex); e = e2; }
      }
      p = e;
      return v;
    };
    g.nnodes = next_int("header"); g.ntri = next_int("header"); g.nquad = next_int("header");
    if (g.nnodes <= 0 || g.ntri < 0 || g.nquad < 0) stop("grid header");
    while (*p && *p != '\n') p++;  // a list-directed read ignores the rest of the header record
    g.xy.resize(2 * (size_t)g.nnodes);
    for (size_t i = 0; i < g.xy.size(); i++) g.xy[i] = next_real("nodes");
    const size_t nn = 3 * (size_t)g.ntri + 4 * (size_t)g.nquad;
    g.cnode.resize(nn);
    for (size_t i = 0; i < nn; i++) g.cnode[i] = next_int("cells") - 1;
    if (use_bin) gridbin_write(base, g);
  }
  const int nc = g.ntri + g.nquad;
  g.cptr.resize(nc + 1); g.cptr[0] = 0;
  for (int i = 0; i < nc; i++) g.cptr[i + 1] = g.cptr[i] + (i < g.ntri ? 3 : 4);
  std::ifstream b(base + ".bc");
  if (!b) stop("cannot find " + base + ".bc file!");
  std::string s;
  std::getline(b, s);
  const int nb = std::atoi(tokens(s, 1)[0].c_str());
  for (int ib = 0; ib < nb; ib++) {
    std::getline(b, s);
    auto t = tokens(s, 2);
    g.bn.push_back(std::atoi(t[0].c_str()));
    const std::string ty = t[1];
    int code = ty == "freestream" ? FVS2D_BC_FREESTREAM : ty == "slip_wall" ? FVS2D_BC_SLIP_WALL : ty == "solid_wall" ? FVS2D_BC_SOLID_WALL
               : ty == "dirichlet" ? FVS2D_BC_DIRICHLET : 0;
    if (!code) stop("Boundary condition=" + ty + "  not implemented");
    g.bt.push_back(code);
  }
  for (int ib = 0; ib < nb; ib++)
    for (int i = 0; i < g.bn[ib]; i++) { std::getline(b, s); g.bc.push_back(std::atoi(tokens(s, 1)[0].c_str()) - 1); }
  return g;
}

// log.fvs2d: the echo of the parsed input (src/input.f90:283-409), same lines and edit descriptors.  aN right-justifies a
// shorter string in N columns (and keeps the leftmost N of a longer one).
std::string fa(const std::string &t, size_t w) { return t.size() >= w ? t.substr(0, w) : std::string(w - t.size(), ' ') + t; }
std::string ffix(double v, int w, int d) {
  char b[64]; std::snprintf(b, sizeof b, "%*.*f", w, d, v);
  std::string s = b;
  return (int)s.size() > w ? std::string(w, '*') : s;  // Fortran fills an overflowing field with asterisks
}
void write_log_input(const Input &in) {
  std::ofstream f("log.fvs2d");
  const std::string bar(139, '='), dash(139, '-');
  f << bar << "\n     FVM2D CODE                       \n" << bar << "\n";
  f << fa("grid file name: ", 35) << in.base << ".grid\n" << fa("bc file name: ", 35) << in.base << ".bc\n";
  f << fa("Reynolds number: ", 35) << fortran_e(in.rey, 16, 8) << "\n" << fa("Mach number: ", 35) << fortran_e(in.mach, 16, 8) << "\n"
    << fa("Flow angle (deg): ", 35) << fortran_e(in.aoa, 16, 8) << "\n" << dash << "\n";
  f << fa("Time-step: ", 35) << fortran_e(in.dt, 16, 8) << "\n" << fa("#s of total time-steps: ", 35) << in.ntimes << "\n"
    << fa("#s of output solution files: ", 35) << in.nsaves << "\n";
  if (in.nsaves > 0 && in.ntimes % in.nsaves == 0) f << fa("Interval to output solution files: ", 35) << in.ntimes / in.nsaves << "\n";
  else if (in.nsaves > 0)
    f << fa("Interval to output solution files: ", 35) << in.ntimes / in.nsaves + 1 << " & " << in.ntimes - (in.ntimes / in.nsaves + 1) * (in.nsaves - 1) << "\n";
  if (in.steady) {
    f << fa("Steady flow is computed: ", 35) << "local time-stepping is employed (input dt is ignored)\n";
    f << fa("Local dt is computed based on CFL=", 35) << ffix(in.cfl, 4, 2) << "\n";
  } else {
    f << fa("Starting time-step: ", 35) << in.ntstart << "\n" << fa("t_inital: ", 35) << fortran_e((double)(in.ntstart - 1) * in.dt, 16, 8) << "\n"
      << fa("t_final : ", 35) << fortran_e((double)(in.ntstart - 1) * in.dt + (double)in.ntimes * in.dt, 16, 8) << "\n";
  }
  const char *vort1 = " primative varialbes are initialized with isentropic vortex\n", *vort2 = " Code will read freestream parameters from fvs2d.vortex file\n";
  if (in.ntstart < 1) f << " primative varialbes are initialized with manufactured solution\n";
  else if (in.ntstart > 1) { f << " primative varialbes are initialized with continuation files\n"; if (in.vortex) f << vort1 << vort2; }
  else if (in.vortex) f << vort1 << vort2;
  else f << " primative varialbes are initialized with freestream values\n";
  f << dash << "\n";
  const bool any = in.lw[0] || in.lw[1] || in.lw[2] || in.lw[3];
  if (any) f << fa(in.s8 ? " Write out following variables in double precision:" : " Write out following variables in single precision:", 52) << "\n";
  const char *vn[4] = {" density", " u-velocity", " v-velcoity", " pressure"};
  for (int v = 0; v < 4; v++) if (in.lw[v]) f << fa(vn[v], 52) << "\n";
  f << " \n" << bar << "\n     Temporal & Spatial Discretization Schemes      \n" << bar << "\n";
  const std::string gm = fa(" Cell-center gradient method:", 38);
  if (in.grad == 1) f << gm << " Green-Gauss Cell-Base\n";
  else if (in.grad == 2) f << gm << " Green-Gauss Node-Base\n";
  else if (in.grad == 3) {
    const std::string st = in.lsq_nn ? "node" : "face";
    if (in.lsq_pow == 0.0) f << gm << " Unweigghted Least-Squeres based on " << st << " neighbor stencil\n";
    else f << gm << " Weigghted (1/d^" << ffix(in.lsq_pow, 3, 1) << ") Least-Squeres based on " << st << " neighbor stencil\n";
  }
  const char *lim[4] = {" not applied", " Venkatakrishnan", " Barth and Jespersen", " Van Albada"};
  if (in.limiter >= 0 && in.limiter <= 3) f << fa("Gradient limiter:", 38) << lim[in.limiter] << "\n";
  const std::string fr = fa(" Face reconstruction method:", 38);
  if (in.recon == 1) f << fr << " 1st-order upwind\n";
  else if (in.recon == 2) f << fr << " 2nd-order upwind\n";
  else if (in.recon == 3) f << fr << " UMUSCL with Kappa=" << ffix(in.umuscl, 8, 4) << "\n";
  if (in.flux == 1) f << fa(" Inviscid flux discretization scheme:", 38) << " Roe\n";
  f << dash << "\n" << fa(in.ssprk ? "Runge-Kutta SSP formualtion is employed" : "Runge-Kutta standard formualtion is employed", 50) << "\n";
  f << fa("#s of stages for Runge-Kutta time-integration: ", 52) << in.rk_nstages << "\n"
    << fa("Order of accuracy of Runge-Kutta time-integration: ", 52) << in.rk_order << "\n";
  f << dash << "\n Transient output files         \n";
}

}  // namespace

int main(int argc, char **argv) {
  const bool check_only = argc > 1 && std::string(argv[1]) == "--check";
  const int device = (argc > 1 && !check_only) ? std::atoi(argv[1]) : -1;
  Input in = input_read();
  write_log_input(in);
  Grid g = grid_read(in.base);
  const int nc = g.ntri + g.nquad;

  fvs2d_config c;
  std::memset(&c, 0, sizeof c);
  c.gamma = in.gamma; c.dt = in.dt; c.cfl_user = in.cfl;
  c.umuscl_cst = in.recon == 3 ? in.umuscl : 0.0;  // src/input.f90:248-254
  c.lsq_pow = in.lsq_pow; c.grad_method = in.grad; c.lsq_stencil = in.lsq_nn ? 1 : 0; c.limiter = in.limiter;
  c.recon = in.recon; c.flux = in.flux; c.rk_nstages = in.rk_nstages; c.rk_order = in.rk_order;
  c.ssprk = in.ssprk; c.steady = in.steady; c.lvortex = in.vortex; c.ntstart = in.ntstart;
  const double pi = std::acos(-1.0);
  c.pvar_inf[0] = 1.0;                                             // src/data_solution.f90:55-63
  c.pvar_inf[1] = in.mach * (in.aoa == 0.0 ? 1.0 : std::cos(in.aoa * pi / 180.0));
  c.pvar_inf[2] = in.mach * (in.aoa == 0.0 ? 0.0 : std::sin(in.aoa * pi / 180.0));
  c.pvar_inf[3] = 1.0 / in.gamma;
  c.vortex_pos[0] = in.vpos[0]; c.vortex_pos[1] = in.vpos[1]; c.vortex_kappa = in.vkap;
  for (int i = 0; i < 4; i++) c.vortex_inf[i] = in.vinf[i];
  const double mms[4][4] = {{1.12, 0.15, 3.12 * pi, 2.92 * pi}, {1.32, 0.06, 2.09 * pi, 3.12 * pi},  // src/mms.f90:80-101
                            {1.18, 0.03, 2.15 * pi, 3.32 * pi}, {1.62, 0.31, 3.79 * pi, 2.98 * pi}};
  std::memcpy(c.mms_c, mms, sizeof mms);
  c.ngpus = 1;

  if (check_only)  // host half of set_mesh only: connectivity, geometry, gradient operator, renumbering (no CUDA call)
    check(fvs2d_host_build(&c, 0, 1, g.nnodes, g.ntri, g.nquad, g.xy.data(), g.cptr.data(), g.cnode.data(), (int)g.bn.size(), g.bn.data(),
                           g.bt.data(), g.bc.data()));
  else {
    check(fvs2d_gpu_init(&c, device));
    check(fvs2d_gpu_set_mesh(g.nnodes, g.ntri, g.nquad, g.xy.data(), g.cptr.data(), g.cnode.data(), (int)g.bn.size(), g.bn.data(),
                             g.bt.data(), g.bc.data()));
  }
  int sizes[10]; double scal[6];
  check(fvs2d_gpu_sizes(sizes)); check(fvs2d_gpu_scalars(scal));
  {  // log.grid (src/grid_procs.f90:857-878)
    std::ofstream lg("log.grid");
    lg << "\n number of nodes: " << sizes[0] << "\n number of triangle cells: " << g.ntri << "\n number of quadrilateral cells: " << g.nquad
       << "\n number of total cells: " << sizes[1] << "\n number of total edges: " << sizes[2] << "\n number of total bondary edges: " << sizes[4]
       << "\n number of total bondary cells: " << sizes[6] << "\n\n";
    char b[256];
    std::snprintf(b, sizeof b, " Sum of the cell volumes via numerical cal: %.11E\n Sum of the cell volumes via Green theorem: %.11E\n\n"
                  " cell effective length, sqrt[sum(vol)/ncells]: %.11E\n cell effective length, sum[sqrt(vol)]/ncells: %.11E\n", scal[2], scal[3], scal[0], scal[1]);
    lg << b;
  }

  if (check_only) {
    std::printf(" nodes=%d cells=%d (tri=%d quad=%d) edges=%d (interior=%d boundary=%d) interior cells=%d boundary cells=%d\n", sizes[0], sizes[1],
                g.ntri, g.nquad, sizes[2], sizes[3], sizes[4], sizes[5], sizes[6]);
    std::printf(" sum(vol)=%.11E green=%.11E lsq_verify=%.3E\n o.k. (check only)\n", scal[2], scal[3], scal[4]);
    return 0;
  }

  // ---- initial condition or restart (src/initialize.f90:19-90)
  std::vector<double> cvar(4 * (size_t)nc);
  if (in.ntstart > 1) {
    std::ifstream f("cont.s8", std::ios::binary);
    if (!f) stop("cannot find cont.s8 (restart, ntstart>1)");
    std::vector<unsigned char> rec(8 * (size_t)nc);
    for (int iv = 0; iv < 4; iv++) {
      f.read((char *)rec.data(), rec.size());
      if (!f) stop("dimension between grid and cont files does not match!");
      for (int ic = 0; ic < nc; ic++) {
        uint64_t u = 0;
        for (int b = 0; b < 8; b++) u = (u << 8) | rec[8 * (size_t)ic + b];
        std::memcpy(&cvar[4 * (size_t)ic + iv], &u, 8);
      }
    }
    check(fvs2d_gpu_set_state(cvar.data()));
  } else {
    check(fvs2d_gpu_initialize_solution());
  }

  if (in.ntstart == 0) {  // as shipped the reference runs test_resid here and stops (src/fvs2d.f90:125-126, src/test.f90:481-519)
    double l2[4], li[4];
    check(fvs2d_gpu_test_resid(0, l2, li));
    const bool exists = (bool)std::ifstream("error_resid.plt");
    std::ofstream f("error_resid.plt", std::ios::app);
    if (!exists) f << "variables = \"h<sub>eff\" \"L<sub>2,rho\" \"L<sub>2,u\"   \"L<sub>2,v\"  \"L<sub>2,e\"  \"L<sub>inf,rho\" \"L<sub>inf,u\"   \"L<sub>inf,v\"  \"L<sub>inf,e\"    \n";
    f << fortran_e(scal[0], 16, 9) << " ";
    for (int v = 0; v < 4; v++) f << fortran_e(l2[v], 16, 9) << " ";
    for (int v = 0; v < 4; v++) f << fortran_e(li[v], 16, 9) << " ";
    f << "\n";
    std::printf("ok\n");
    fvs2d_gpu_finalize();
    return 0;
  }

  // ---- save-interval bookkeeping (src/input.f90:127-136)
  std::vector<int> nsub(in.nsaves);
  if (in.ntimes % in.nsaves == 0) std::fill(nsub.begin(), nsub.end(), in.ntimes / in.nsaves);
  else {
    for (int i = 0; i + 1 < in.nsaves; i++) nsub[i] = in.ntimes / in.nsaves + 1;
    nsub[in.nsaves - 1] = in.ntimes - (in.ntimes / in.nsaves + 1) * (in.nsaves - 1);
  }

  // ---- io_init (src/io.f90:53-116): inst.cd / save.cd
  const char *names[4] = {"rho", "u", "v", "p"};
  int mp = 0;
  for (int v = 0; v < 4; v++) mp += in.lw[v];
  std::vector<int> itimes(in.nsaves);
  itimes[0] = in.ntstart - 1 + nsub[0];
  for (int i = 1; i < in.nsaves; i++) itimes[i] = nsub[i] + itimes[i - 1];
  std::ofstream inst;
  int lw_sel[4];
  for (int v = 0; v < 4; v++) lw_sel[v] = in.lw[v];
  if (mp > 0) {
    // the header lists inf(1:mp) = the first mp of (rho,u,v,p) rather than the selected names (src/io.f90:72-87)
    std::vector<std::string> params(names, names + mp);
    writecd("inst", g.nnodes, nc, mp, in.nsaves, itimes, params, {});
    inst.open(in.s8 ? "inst.s8" : "inst.s4", std::ios::binary);
  }
  {
    char dch[16]; std::snprintf(dch, sizeof dch, "%6d", in.ntimes);
    std::string d = dch; d.erase(0, d.find_first_not_of(' '));
    // "#ncells and #nodes are replaced": record length = ncells (src/io.f90:95-113)
    writecd("save", nc, g.nnodes, 4, 1, {in.ntstart - 1 + in.ntimes}, {"rho", "rhou", "rhov", "rhoE"},
            {"number of time-step computed = " + d, "conservatve variables are saved in cell centers", "#ncells and #nodes are replaced", " "});
  }

  // ---- time loop (src/fvs2d.f90:131-160)
  std::ofstream res("log_res.plt");
  res << "variables = \"iteration\" \"|<greek>\\r</greek>|<sub>2</sub>\",  \"|<greek>\\r</greek>u|<sub>2</sub>\", "
         "\"|<greek>\\r</greek>v|<sub>2</sub>\", \"|<greek>\\r</greek>E|<sub>2</sub>\"\n";
  std::ofstream verr, vxy;
  if (in.vortex) {
    verr.open("log_vortex_err.plt");
    verr << "variables = \"t\", \n\"<greek>r</greek><sub>max</sub>\",  \"<greek>r</greek><sub>L1</sub>\", \"<greek>r</greek><sub>L2</sub>\",\n"
            "\"<greek>r</greek>u<sub>max</sub>\",  \"<greek>r</greek>u<sub>L1</sub>\", \"<greek>r</greek>u<sub>L2</sub>\",\n"
            "\"<greek>r</greek>v<sub>max</sub>\",  \"<greek>r</greek>v<sub>L1</sub>\", \"<greek>r</greek>v<sub>L2</sub>\",\n"
            "\"<greek>r</greek>E<sub>max</sub>\",  \"<greek>r</greek>E<sub>L1</sub>\", \"<greek>r</greek>E<sub>L2</sub>\",\n\"Q<sub>L2</sub>\"\n";
    vxy.open("log_vortex_err_xy.plt");
    vxy << "variables = \"t\", \"x\" \"y\"\n";
  }
  // ---- wall post-processing set-up (src/io.f90:340-362): only when a slip_wall / solid_wall boundary exists
  bool lcp = false;
  for (int t : g.bt) lcp = lcp || t == FVS2D_BC_SLIP_WALL || t == FVS2D_BC_SOLID_WALL;
  std::ofstream fcp, fun, fclcd;
  std::vector<int> b_edge, b_edge_ptr;
  std::vector<double> ea, enx, eny;
  if (lcp) {
    fcp.open("log_cp.plt"); fun.open("log_un.plt"); fclcd.open("log_clcd.plt");
    fun << "VARIABLES = \"time\" \"|V<sub>n</sub>|<sub><math>%</math></sub>\",   \"|V<sub>n</sub>|<sub>2</sub>\" , \"|V<sub>n</sub>|<sub>1</sub>\"\n";
    fclcd << "VARIABLES = \"time\" \"c<sub>l</sub>\",   \"c<sub>d</sub>\", \"c<sub>l1</sub>\",   \"c<sub>d1</sub>\"\n";
    auto geti = [&](const char *n, std::vector<int> &v) { v.resize(fvs2d_gpu_mesh_array(n, nullptr)); fvs2d_gpu_mesh_array(n, v.data()); };
    auto getd = [&](const char *n, std::vector<double> &v) { v.resize(fvs2d_gpu_mesh_array(n, nullptr)); fvs2d_gpu_mesh_array(n, v.data()); };
    geti("b_edge", b_edge); geti("b_edge_ptr", b_edge_ptr);
    getd("ea", ea); getd("enx", enx); getd("eny", eny);
  }
  // write_inst_cp_un (src/io.f90:340-449): wall pressure coefficient, normal-velocity norms and force coefficients.  The
  // library extrapolates p, u, v to the wall-edge centres with the unlimited cell gradients on the device
  // (fvs2d_gpu_wall_values: 4 doubles per wall edge come back instead of pvar + grad of every cell); the sums run
  // here in the reference's edge order.
  auto write_inst_cp_un = [&](double sol_time) {
    if (!lcp) return;
    const double p_inf = 1.0 / in.gamma, q2 = 2.0 / (in.mach * in.mach);
    const double ca_ = in.aoa == 0.0 ? 1.0 : std::cos(in.aoa * pi / 180.0), sa_ = in.aoa == 0.0 ? 0.0 : std::sin(in.aoa * pi / 180.0);
    for (size_t ib = 0; ib < g.bt.size(); ib++) {
      if (g.bt[ib] != FVS2D_BC_SLIP_WALL && g.bt[ib] != FVS2D_BC_SOLID_WALL) continue;
      const int e0 = b_edge_ptr[ib], e1 = b_edge_ptr[ib + 1];
      std::vector<double> wv(4 * (size_t)(e1 - e0));
      check(fvs2d_gpu_wall_values((int)ib, wv.data()));
      fcp << "TITLE     = \"cp\"\nVARIABLES = \"x\" \"cp_w\" \"cp_cell\"\nZONE I=" << (e1 - e0) << " J=1\n"
          << "STRANDID=1, SOLUTIONTIME=" << fortran_e(sol_time, 16, 8) << "\n";
      double un_max = 0, un_l2 = 0, un_l1 = 0, cn = 0, ca = 0, cn1 = 0, ca1 = 0;
      for (int i = e0; i < e1; i++) {
        const int ie = b_edge[i];
        const double *w = &wv[4 * (size_t)(i - e0)];  // x_f, p_w, p_cell, u_n
        const double cp = q2 * (w[1] - p_inf), cp1 = q2 * (w[2] - p_inf), un = w[3];
        fcp << fortran_e(w[0], 16, 8) << " " << fortran_e(cp, 16, 8) << " " << fortran_e(cp1, 16, 8) << " \n";
        cn = cn + cp * eny[ie] * ea[ie];   ca = ca + cp * enx[ie] * ea[ie];
        cn1 = cn1 + cp1 * eny[ie] * ea[ie]; ca1 = ca1 + cp1 * enx[ie] * ea[ie];
        un_l2 += un * un; un_l1 += std::fabs(un); un_max = std::max(un_max, std::fabs(un));
      }
      const double n = (double)(e1 - e0);
      fun << fortran_e(sol_time, 16, 8) << " " << fortran_e(un_max, 16, 8) << " " << fortran_e(std::sqrt(un_l2 / n), 16, 8) << " "
          << fortran_e(un_l1 / n, 16, 8) << " \n";
      fclcd << fortran_e(sol_time, 16, 8) << " " << fortran_e(cn * ca_ - ca * sa_, 16, 8) << " " << fortran_e(cn * sa_ + ca * ca_, 16, 8) << " "
            << fortran_e(cn1 * ca_ - ca1 * sa_, 16, 8) << " " << fortran_e(cn1 * sa_ + ca1 * ca_, 16, 8) << " \n";
    }
  };
  double t0 = (double)(in.ntstart - 1) * in.dt, ms_tot = 0, ms_grad = 0, ms_flux = 0;
  int it_tot = 0, icont = 0;
  // FVS2D_KERNEL_TIMERS=1: the reference's per-phase timers (gradient / flux + R-K; src/mainparam.f90:15-16) from a CUDA event
  // pair around every launch.  Off by default: it turns the CUDA-graph replay of the time step off (13-28 % on meshes of
  // 7 k - 65 k cells), so the production run prints the total only.
  const bool kernel_timers = std::getenv("FVS2D_KERNEL_TIMERS") != nullptr && std::atoi(std::getenv("FVS2D_KERNEL_TIMERS")) != 0;
  if (kernel_timers) fvs2d_gpu_set_option("timing", 1);
  for (int it = 0; it < in.nsaves; it++) {
    const int n = nsub[it];
    std::vector<double> r(4 * (size_t)n), ve(14 * (size_t)n), xy(2 * (size_t)n);
    check(fvs2d_gpu_time_integration(t0, n, r.data(), in.vortex ? ve.data() : nullptr, in.vortex ? xy.data() : nullptr));
    double ms[4]; long launches;
    fvs2d_gpu_last_timing(ms, &launches);
    ms_tot += ms[0]; ms_grad += ms[1]; ms_flux += ms[2];
    for (int s = 0; s < n; s++) {
      icont++;
      res << (icont + in.ntstart - 1) << " " << fortran_e(r[4 * s], 16, 8) << fortran_e(r[4 * s + 1], 16, 8) << fortran_e(r[4 * s + 2], 16, 8)
          << fortran_e(r[4 * s + 3], 16, 8) << "\n";
      if (in.vortex) {
        for (int k = 0; k < 14; k++) verr << fortran_e(ve[14 * (size_t)s + k], 16, 8) << " ";
        verr << "\n";
        char b[128];
        std::snprintf(b, sizeof b, " %24.15E %24.15E %24.15E\n", ve[14 * (size_t)s], xy[2 * s], xy[2 * s + 1]);
        vxy << b;
      }
    }
    it_tot += n;
    std::printf("%5d time-steps done \n", it_tot);
    std::fflush(stdout);
    t0 += in.dt * n;
    if (mp > 0) {  // write_inst_ios (src/io.f90:122-150): node-interpolated primitive variables, one record per variable;
      // interpolated on the device, only the mp node records come back
      std::vector<double> fv((size_t)mp * g.nnodes);
      check(fvs2d_gpu_interpolate_cell2node(lw_sel, fv.data()));
      write_be(inst, fv.data(), fv.size(), !in.s8);
    }
    write_inst_cp_un(t0);  // src/fvs2d.f90:157
  }
  // ---- write_save_ios (src/io.f90:156-178): 4 records of cvar at the cell centres, real*8, big-endian
  check(fvs2d_gpu_get_state(cvar.data()));
  {
    std::ofstream f("save.s8", std::ios::binary);
    std::vector<double> rec(nc);
    for (int v = 0; v < 4; v++) {
      for (int ic = 0; ic < nc; ic++) rec[ic] = cvar[4 * (size_t)ic + v];
      write_be(f, rec.data(), rec.size(), false);
    }
  }
  std::printf(" ------------------------------------------------------------------\n");
  if (kernel_timers)
    std::printf(" gpu-time(min): total=%7.3f, grad+limiter=%7.3f, flux+R-K=%7.3f   (%.3e cell-stage updates/s)\n\n", ms_tot / 6e4, ms_grad / 6e4,
                ms_flux / 6e4, (double)nc * 4.0 * it_tot / (ms_tot * 1e-3));
  else
    std::printf(" gpu-time(min): total=%7.3f   (%.3e cell-stage updates/s; FVS2D_KERNEL_TIMERS=1 adds the per-phase timers)\n\n", ms_tot / 6e4,
                (double)nc * 4.0 * it_tot / (ms_tot * 1e-3));
  fvs2d_gpu_finalize();
  std::printf(" o.k.\n");
  return 0;
}
