// api.cu -- context and C-ABI of libfvs2d_gpu.so (include/fvs2d_gpu.h).
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <dlfcn.h>
#include <nccl.h>

#include <omp.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/fvs2d_gpu.h"
#include "host_mesh.hpp"
#include "kernels.cuh"
#include "kernels_fused.cuh"
#include "layout.hpp"

using namespace fvs2d;

namespace {

// ---- NCCL, resolved lazily so that single-GPU use has no NCCL dependency ------------------------
struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool load(std::string &err) {
    if (lib) return true;
    const char *names[] = {"libnccl.so.2", "/opt/prime-rl/.venv/lib/python3.12/site-packages/nvidia/nccl/lib/libnccl.so.2", "libnccl.so"};
    for (const char *n : names) { lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
    if (!lib) { err = std::string("cannot load libnccl.so.2: ") + dlerror(); return false; }
#define SYM(f) f = (decltype(f))dlsym(lib, "nccl" #f); if (!f) { err = "libnccl lacks nccl" #f; return false; }
    SYM(GetUniqueId) SYM(CommInitRank) SYM(CommDestroy) SYM(Send) SYM(Recv) SYM(AllReduce) SYM(AllGather) SYM(GroupStart) SYM(GroupEnd) SYM(GetErrorString)
#undef SYM
    return true;
  }
};
NcclApi g_nccl;

struct FusedLaunch {  // one launch of k_stage_fused over a subset of the tiles
  FusedMeta meta{};
  int n_bnd = 0;            // leading boundary tiles of meta.tile_list (several ranks)
  std::vector<int> tiles;   // host copy of meta.tile_list
  const void *attr_key = nullptr;
  int per3 = 0, per2 = 0;
};

struct GlobalInfo {  // counts and sums of the WHOLE mesh (a partition-local build knows only its share until they are reduced)
  int nnodes = 0, ncells = 0, nedges = 0, nedges_intr = 0, nedges_bndr = 0, ncells_intr = 0, ncells_bndr = 0;
  double heff1 = 0, heff2 = 0, vol_sum = 0, vol_green = 0, xy_cell0[2] = {0, 0};
  bool partial_sums = false;
};

struct Ctx {
  bool inited = false, has_mesh = false, has_state = false;
  fvs2d_config cfg{};
  int device = 0;
  cudaStream_t st = nullptr;
  // communicator
  int rank = 0, nranks = 1;
  ncclComm_t comm = nullptr;
  // host products
  HostMesh mesh;          // the whole mesh, or (several ranks) this rank's submesh
  SubMesh sub;            // numbering of that submesh (its HostMesh is moved into `mesh`)
  GlobalInfo gi;
  GradOp grad;
  Layout L;
  // device
  std::vector<void *> allocs;
  size_t bytes = 0;
  DevMesh dm{};
  Phys phys{};
  int np = 0, nblocks = 0;
  int recon = RC_K0;
  double *q = nullptr, *f = nullptr, *pa = nullptr, *pb = nullptr, *g = nullptr, *phi = nullptr, *dtl = nullptr;
  double *resid = nullptr, *ws = nullptr, *bc = nullptr, *partial = nullptr, *vpartial = nullptr;
  int *vbest = nullptr, *vbest_loc = nullptr;
  double *stage_aos = nullptr;  // nc_global*8 doubles staging for AoS <-> SoA
  size_t stage_aos_len = 0;
  double *h_stage = nullptr;    // pinned host staging (several ranks: the caller's global arrays are permuted on the host)
  size_t h_stage_len = 0;
  double *logbuf = nullptr; int *logid = nullptr; size_t log_cap = 0;
  int *send_idx = nullptr; double *sendbuf = nullptr;
  PipeMeta pm{};
  int opt_tile = 2;      // pass-B kernel: 2 persistent smem pipeline (default), 0 direct gather
  bool tile_ok = false;
  int nsm = 148, nparts = 0;
  StepClock *clk = nullptr;            // device step clock
  int opt_graph = 1;                   // replay time steps as a CUDA graph (single GPU, timing off)
  cudaGraphExec_t graph_exec = nullptr;
  const void *graph_logbuf = nullptr;  // the graph is tied to these
  int graph_um = -1;
  int opt_ctas = 0;      // persistent pass-B CTAs per SM: 0 = as many as shared memory allows (max 3)
  int opt_pair = -1;     // small meshes on one GPU: two threads per cell in both passes (k_gradient2 / k_flux_rk2); -1 automatic
                         // (meshes that leave most thread slots of the machine empty with one thread per cell), 0 never, 1 always
  int opt_pdl = 1;       // programmatic dependent launch between the kernels of a time step (one GPU)
  bool pdl_on = false;   // set while fvs2d_gpu_time_integration issues / captures its step sequence
  int opt_overlap = 1;   // multi-GPU: overlap the halo exchanges with interior-tile work (second stream)
  int opt_fuse = -1;     // one kernel per stage (k_stage_fused) where it applies -- second-order upwind reconstruction without
                         // limiter, gradient tables that fit shared memory: -1 automatic (default), 0 never (two-pass path),
                         // 2 force a single launch per stage (no split by shared-memory need)
  int fz_state = 0;      // launch plan: 0 not prepared, 1 ready, -1 not applicable to this mesh / scheme
  int fz_tables = 0;     // device tables: 0 not uploaded, 1 uploaded, -1 this mesh has none
  FusedMeta fm{};        // table pointers shared by the launches of a stage
  FusedLaunch fl[2];     // the launch plan of a stage (build_fused_plan)
  int n_fl = 0;
  // in-kernel halo exchange of the fused path (several ranks): peers' arrays mapped with CUDA IPC
  bool p2p_ok = false;
  double *p_buf[2] = {nullptr, nullptr};    // the two primitive-state buffers (pa / pb swap every stage)
  double2 *peer_p[kMaxPeers][2] = {};       // the same two buffers of every peer
  int peer_np[kMaxPeers] = {}, peer_recv_begin[kMaxPeers] = {};
  unsigned *peer_flag[kMaxPeers] = {};
  unsigned *flags = nullptr, *done_ctr = nullptr;  // flags[r]: stages rank r has delivered to this rank
  int *p2p_timed_out = nullptr;                    // device flag: a wait for a peer gave up
  std::vector<void *> ipc_open;
  std::map<const void *, size_t> smem_attr;  // kernel -> dynamic shared memory it is configured for (per context: function
                                             // attributes are per device)
  const uint32_t *d_rs_word = nullptr;
  const int2 *d_rs_ent = nullptr;
  unsigned epoch_total = 0;  // Runge-Kutta stages run through the in-kernel exchange so far (all ranks agree)
  cudaStream_t sx = nullptr;  // exchange stream
  cudaEvent_t e_a = nullptr, e_g = nullptr, e_b = nullptr, e_p = nullptr;
  const int *d_tile_int = nullptr, *d_tile_bnd = nullptr;
  int n_int = 0, n_bnd = 0;
  bool bc_static_done = false;
  double rk_coef[4], h_rk[4], dts[4], dte[4];
  // output path, uploaded on first use: node -> cell lists with inverse-distance weights, boundary-edge tables
  std::vector<int> out_nodes;            // original ids of the nodes this rank interpolates to (the nodes of its owned cells)
  std::vector<int> be_ptr, be_pos;       // owned boundary edges per boundary; their positions in the caller's edge lists
  const int *d_n2c_ptr = nullptr, *d_n2c = nullptr;
  const double *d_idw = nullptr;
  double *d_fnode = nullptr;
  const int *d_be_cell = nullptr;
  const double2 *d_be_xy = nullptr, *d_be_nxy = nullptr;
  double *d_be_out = nullptr;
  // timing
  int opt_timing = 0;
  double last_ms[4] = {0, 0, 0, 0};
  long last_launches = 0;
  std::vector<cudaEvent_t> ev_pool;
  std::vector<std::pair<int, int>> ev_spans[3];
  size_t ev_used = 0;
  long ev_dropped = 0;   // spans not recorded because the event pool was in use (time_integration drains it every 256 steps)
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};
Ctx *C = nullptr;
std::string g_err;

int fail(const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return 1;
}
#define CUDA_OK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail("CUDA error %s at %s:%d (%s)", cudaGetErrorString(e_), __FILE__, __LINE__, #x); } while (0)
#define NCCL_OK(x) do { ncclResult_t r_ = (x); if (r_ != ncclSuccess) return fail("NCCL error %s at %s:%d", g_nccl.GetErrorString(r_), __FILE__, __LINE__); } while (0)
#define NEED(cond, msg) do { if (!(cond)) return fail("%s", msg); } while (0)

template <class T>
int dev_alloc(T *&p, size_t n) {
  p = nullptr;
  if (n == 0) n = 1;
  CUDA_OK(cudaMalloc((void **)&p, n * sizeof(T)));
  C->allocs.push_back(p);
  C->bytes += n * sizeof(T);
  return 0;
}
template <class T>
int dev_upload(const T *&p, const std::vector<T> &h) {
  T *d;
  if (dev_alloc(d, h.size())) return 1;
  if (!h.empty()) CUDA_OK(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  p = d;
  return 0;
}
template <class T>
void dev_free(T *&p, size_t bytes) {  // release one tracked allocation before free_device
  if (!p) return;
  cudaFree(p);
  C->allocs.erase(std::remove(C->allocs.begin(), C->allocs.end(), (void *)p), C->allocs.end());
  C->bytes -= std::min(C->bytes, bytes);
  p = nullptr;
}
inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// raise a kernel's dynamic shared-memory limit to `bytes` (never lowers it: several launch configurations share a kernel)
template <class K>
void ensure_smem_attr(K kernel, size_t bytes) {
  size_t &have = C->smem_attr[(const void *)kernel];
  if (have >= bytes) return;
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  have = bytes;
}

void close_p2p();
int all_ranks_agree(int mine, int &all);
void free_device() {
  if (C->p2p_ok) {  // peers have this rank's buffers mapped: everybody unmaps before anybody frees
    close_p2p();
    int dummy = 0;
    if (C->comm && C->st) all_ranks_agree(1, dummy);
  }
  for (void *p : C->allocs) cudaFree(p);
  C->allocs.clear();
  C->bytes = 0;
  C->stage_aos = nullptr; C->stage_aos_len = 0;
  if (C->h_stage) cudaFreeHost(C->h_stage);
  C->h_stage = nullptr; C->h_stage_len = 0;
  C->logbuf = nullptr; C->log_cap = 0;
  C->d_n2c_ptr = C->d_n2c = nullptr; C->d_idw = nullptr; C->d_fnode = nullptr;
  C->d_be_cell = nullptr; C->d_be_xy = C->d_be_nxy = nullptr; C->d_be_out = nullptr;
  C->fz_state = 0; C->n_fl = 0; C->fz_tables = 0;
  C->flags = nullptr; C->done_ctr = nullptr; C->p2p_timed_out = nullptr; C->d_rs_word = nullptr; C->d_rs_ent = nullptr;
}

// launch of a kernel of the time-step chain: with programmatic stream serialisation while a step sequence is being issued
// (C->pdl_on: one GPU, no per-launch timing), so that the kernel may be scheduled while its predecessor drains
// (pdl_entry() in the kernels); a plain launch otherwise
template <class... KArgs, class... Args>
void launch_chain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, Args &&...args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = C->st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = C->pdl_on ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// ---- event-based kernel timing (option "timing") ------------------------------------------------
struct Span {
  int bucket, a = -1;
  Span(int b) : bucket(b) {
    if (!C->opt_timing) return;
    if (C->ev_used + 2 > C->ev_pool.size()) { C->ev_dropped++; return; }
    a = (int)C->ev_used;
    C->ev_used += 2;
    cudaEventRecord(C->ev_pool[a], C->st);
  }
  ~Span() {
    if (a < 0) return;
    cudaEventRecord(C->ev_pool[a + 1], C->st);
    C->ev_spans[bucket].push_back({a, a + 1});
  }
};

int recon_mode(const fvs2d_config &c) {
  const double kap = c.recon == 3 ? c.umuscl_cst : 0.0;  // src/input.f90:248-254
  if (c.recon == 1) return RC_FIRST;
  if (kap == 0.0) return c.limiter > 0 ? RC_K0_PHI : RC_K0;
  return RC_GENERAL;
}

int rk_setup() {  // src/runge_kutta.f90:25-88
  const fvs2d_config &c = C->cfg;
  const double dt = c.dt;
  NEED(c.rk_nstages == 4, "only rk_nstages==4 is coded (src/runge_kutta.f90:36)");
  double a, b, cc, d;
  switch (c.rk_order) {
    case 1: a = 3.60897; b = 2.04; cc = 0.34206; d = 0.00897; break;
    case 2: a = 0.11; b = 3.92; cc = 1.86; d = 0.11; break;
    case 3: a = 0.65; b = 2.7; cc = 2.0; d = 0.65; break;
    case 4: a = 1.0; b = 2.0; cc = 2.0; d = 1.0; break;
    default: return fail("rk_order must be 1..4");
  }
  C->h_rk[0] = dt / 2.0; C->rk_coef[0] = a;
  C->h_rk[1] = dt / 2.0; C->rk_coef[1] = b;
  C->h_rk[2] = dt;       C->rk_coef[2] = cc;
  C->h_rk[3] = dt / 6.0; C->rk_coef[3] = d;
  for (int k = 0; k < 4; k++) C->dts[k] = C->dte[k] = 0;
  if (c.ssprk) {
    NEED(c.rk_order == 2, "SSPRK is coded only for 4 stages, order 2 (src/runge_kutta.f90:56)");
    for (int k = 0; k < 3; k++) { C->h_rk[k] = dt / 3.0; C->rk_coef[k] = 1.0; }
    C->h_rk[3] = dt / 4.0; C->rk_coef[3] = 1.0;
    C->dts[0] = 0.0;            C->dte[0] = dt / 3.0;
    C->dts[1] = dt / 3.0;       C->dte[1] = dt * 2.0 / 3.0;
    C->dts[2] = dt * 2.0 / 3.0; C->dte[2] = dt;
    C->dts[3] = dt;             C->dte[3] = dt;
  }
  return 0;
}

// ---- halo exchange: owned send-cells -> the peers' ghost runs, nv variables of SoA array a ---------
// nv = number of double variables; pair != 0: the array is pair-interleaved (nv/2 double2 variables)
struct HaloItem { double *a; int nv; int pair; };
int halo_exchange(const HaloItem *items, int nitems, cudaStream_t st = nullptr) {
  if (C->nranks == 1) return 0;
  if (!st) st = C->st;
  const Layout &L = C->L;
  const int nsend = L.send_ptr.empty() ? 0 : L.send_ptr.back();
  // pack: sendbuf layout [item][element variable][all send cells]
  size_t boff = 0;
  for (int it = 0; it < nitems; it++) {
    if (nsend) {
      if (items[it].pair)
        k_pack<double2><<<cdiv(nsend, 256), 256, 0, st>>>(nsend, items[it].nv / 2, C->np, C->send_idx, reinterpret_cast<const double2 *>(items[it].a),
                                                        reinterpret_cast<double2 *>(C->sendbuf + boff));
      else
        k_pack<double><<<cdiv(nsend, 256), 256, 0, st>>>(nsend, items[it].nv, C->np, C->send_idx, items[it].a, C->sendbuf + boff);
    }
    C->last_launches++;
    boff += (size_t)items[it].nv * nsend;
  }
  NCCL_OK(g_nccl.GroupStart());
  boff = 0;
  for (int it = 0; it < nitems; it++) {
    const int w = items[it].pair ? 2 : 1, nvar = items[it].nv / w;  // element width in doubles, element variables
    for (size_t pi = 0; pi < L.peers.size(); pi++) {
      const int peer = L.peers[pi];
      const int s0 = L.send_ptr[pi], sn = L.send_ptr[pi + 1] - s0;
      for (int v = 0; v < nvar; v++) {
        if (sn) NCCL_OK(g_nccl.Send(C->sendbuf + boff + ((size_t)v * nsend + s0) * w, (size_t)sn * w, ncclDouble, peer, C->comm, st));
        if (L.recv_count[pi])
          NCCL_OK(g_nccl.Recv(items[it].a + ((size_t)v * C->np + L.recv_begin[pi]) * w, (size_t)L.recv_count[pi] * w, ncclDouble, peer, C->comm, st));
      }
    }
    boff += (size_t)items[it].nv * nsend;
  }
  NCCL_OK(g_nccl.GroupEnd());
  return 0;
}

// pinned host buffer of at least `ndoubles` (grown, never shrunk; freed with the device arrays)
int ensure_host_stage(size_t ndoubles) {
  if (C->h_stage_len >= ndoubles) return 0;
  if (C->h_stage) cudaFreeHost(C->h_stage);
  C->h_stage = nullptr; C->h_stage_len = 0;
  CUDA_OK(cudaMallocHost((void **)&C->h_stage, std::max<size_t>(ndoubles, 1) * sizeof(double)));
  C->h_stage_len = ndoubles;
  return 0;
}

int ensure_stage(size_t ndoubles) {
  if (C->stage_aos_len >= ndoubles) return 0;
  if (C->stage_aos) {
    cudaFree(C->stage_aos);
    C->allocs.erase(std::remove(C->allocs.begin(), C->allocs.end(), (void *)C->stage_aos), C->allocs.end());
    C->bytes -= C->stage_aos_len * 8;
  }
  C->stage_aos_len = 0;
  if (dev_alloc(C->stage_aos, ndoubles)) return 1;
  C->stage_aos_len = ndoubles;
  return 0;
}

// device SoA (local numbering, owned cells) -> caller AoS (original numbering).  One rank: permuted on the device.  Several
// ranks: only the owned cells cross PCIe (local order), the host scatters them into the caller's global array.
int download_aos(const double *soa, int nvar, double *host_out, int pair = 0, int v0 = 0) {
  const int n_own = C->L.n_own;
  if (C->nranks == 1) {
    const size_t ng = (size_t)C->L.nc_global * nvar;
    if (ensure_stage(ng)) return 1;
    k_gather_out<<<cdiv(n_own, 256), 256, 0, C->st>>>(n_own, C->np, nvar, C->dm.orig_id, soa, C->stage_aos, pair, v0);
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemcpyAsync(host_out, C->stage_aos, ng * 8, cudaMemcpyDeviceToHost, C->st));
    CUDA_OK(cudaStreamSynchronize(C->st));
    return 0;
  }
  const size_t nl = (size_t)n_own * nvar;
  if (ensure_stage(nl)) return 1;
  k_gather_out<<<cdiv(n_own, 256), 256, 0, C->st>>>(n_own, C->np, nvar, nullptr, soa, C->stage_aos, pair, v0);
  CUDA_OK(cudaGetLastError());
  if (ensure_host_stage(nl)) return 1;
  const double *tmp = C->h_stage;
  CUDA_OK(cudaMemcpyAsync(C->h_stage, C->stage_aos, nl * 8, cudaMemcpyDeviceToHost, C->st));
  CUDA_OK(cudaStreamSynchronize(C->st));
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n_own; i++) {
    const size_t o = C->L.orig_id[i];
    for (int v = 0; v < nvar; v++) host_out[o * nvar + v] = tmp[(size_t)i * nvar + v];
  }
  return 0;
}

// conserved state of the LOCAL cells (owned + ghosts, library order, 4 doubles per cell) -> device q and primitive pa
int upload_local_state(const double *cv_loc) {
  const int n_loc = C->L.n_loc;
  const size_t n = (size_t)n_loc * 4;
  if (ensure_stage(n)) return 1;
  CUDA_OK(cudaMemcpyAsync(C->stage_aos, cv_loc, n * 8, cudaMemcpyHostToDevice, C->st));
  k_scatter_in<<<cdiv(n_loc, 256), 256, 0, C->st>>>(n_loc, C->np, 4, nullptr, C->stage_aos, C->q, 0);
  k_prim<<<cdiv(n_loc, 256), 256, 0, C->st>>>(n_loc, C->np, C->cfg.gamma, C->q, C->pa);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaStreamSynchronize(C->st));
  C->has_state = true;
  return 0;
}

// two threads per cell: meshes that would leave most of the machine's thread slots empty (one GPU, whole-mesh launches)
bool use_pair_kernels() {
  if (C->nranks != 1 || C->opt_pair == 0) return false;
  // measured (profiles/r2_small_meshes.md): 7 k cells 47.1 -> 38.9 us per step, 65 k cells 86.0 -> 76.4 us: automatic while one
  // thread per cell cannot fill the machine's thread slots (512 threads per SM at 128 registers)
  return C->opt_pair > 0 || C->L.n_own <= C->nsm * 512;   // B200: 75 776 cells
}

int launch_gradient(const double *p, const int *list = nullptr, int nlist = 0, bool force = false) {
  if (C->recon == RC_FIRST && !force) return 0;  // src/gradient.f90:49
  const int nb = list ? nlist : C->nblocks;
  if (nb == 0) return 0;
  Span sp(1);
  const bool lim = C->cfg.limiter > 0;
  if (!list && use_pair_kernels()) {
    const int nb2 = cdiv(2 * C->L.n_own, kBlock);
    if (C->L.g_form == 0) {
      if (lim) launch_chain(k_gradient2<0, true>, dim3(nb2), dim3(kBlock), 0, C->dm, C->cfg.limiter, p, C->g, C->phi);
      else launch_chain(k_gradient2<0, false>, dim3(nb2), dim3(kBlock), 0, C->dm, C->cfg.limiter, p, C->g, C->phi);
    } else {
      if (lim) launch_chain(k_gradient2<1, true>, dim3(nb2), dim3(kBlock), 0, C->dm, C->cfg.limiter, p, C->g, C->phi);
      else launch_chain(k_gradient2<1, false>, dim3(nb2), dim3(kBlock), 0, C->dm, C->cfg.limiter, p, C->g, C->phi);
    }
    C->last_launches++;
    return 0;
  }
  if (C->L.g_form == 0) {
    if (lim) launch_chain(k_gradient<0, true>, dim3(nb), dim3(kBlock), 0, C->dm, C->cfg.limiter, p, C->g, C->phi, list);
    else launch_chain(k_gradient<0, false>, dim3(nb), dim3(kBlock), 0, C->dm, C->cfg.limiter, p, C->g, C->phi, list);
  } else {
    if (lim) launch_chain(k_gradient<1, true>, dim3(nb), dim3(kBlock), 0, C->dm, C->cfg.limiter, p, C->g, C->phi, list);
    else launch_chain(k_gradient<1, false>, dim3(nb), dim3(kBlock), 0, C->dm, C->cfg.limiter, p, C->g, C->phi, list);
  }
  C->last_launches++;
  return 0;
}

struct TileSel { const int *list = nullptr; int n = 0, part_off = 0; };  // subset of tiles for the pipeline kernel
TileSel g_sel;

template <int UM, bool STEADY, int RC>
void launch_flux_one(const StageParams &S, const double *pin, double *pout) {
  const int nb = C->nblocks;
  if (!g_sel.list && use_pair_kernels()) {
    const int nb2 = cdiv(2 * C->L.n_own, kBlock);
    launch_chain(k_flux_rk2<UM, STEADY, RC>, dim3(nb2), dim3(kBlock), 0, C->dm, C->phys, S, pin, C->g, C->phi, C->bc, C->q, C->f, pout, C->dtl, C->resid,
                                                          C->ws, C->partial);
    C->nparts = nb2;
    return;
  }
  if (C->tile_ok && C->opt_tile == 2) {
    const size_t smem = kStages * pipe_stage_bytes<RC>(C->pm.S, C->pm.E) + 2 * kStages * sizeof(uint64_t);
    ensure_smem_attr(k_flux_pipe<UM, STEADY, RC>, smem);
    int per_sm = 0;  // persistent kernel: exactly as many CTAs as are co-resident
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_flux_pipe<UM, STEADY, RC>, kPipeThreads, smem);
    per_sm = std::max(1, per_sm);
    if (C->opt_ctas > 0) per_sm = std::min(per_sm, C->opt_ctas);
    if (getenv("FVS2D_DEBUG")) { static int once = 0; if (!once++) fprintf(stderr, "[fvs2d] k_flux_pipe: %zu B smem/CTA, %d CTAs/SM\n", smem, per_sm); }
    PipeMeta pm = C->pm;
    if (g_sel.list) { pm.tile_list = g_sel.list; pm.ntiles = g_sel.n; }
    const int grid = std::min(pm.ntiles, C->nsm * per_sm);
    if (grid > 0)
      launch_chain(k_flux_pipe<UM, STEADY, RC>, dim3(grid), dim3(kPipeThreads), smem, C->dm, pm, C->phys, S, pin, C->g, C->phi, C->bc, C->q, C->f, pout,
                                                                       C->dtl, C->resid, C->ws, C->partial + 4 * (size_t)g_sel.part_off);
    C->nparts = g_sel.part_off + grid;
  } else {
    launch_chain(k_flux_rk<UM, STEADY, RC>, dim3(nb), dim3(kBlock), 0, C->dm, C->phys, S, pin, C->g, C->phi, C->bc, C->q, C->f, pout, C->dtl, C->resid,
                                                        C->ws, C->partial);
    C->nparts = nb;
  }
}

template <int UM, bool STEADY>
void launch_flux_rc(const StageParams &S, const double *pin, double *pout) {
  switch (C->recon) {
    case RC_FIRST: launch_flux_one<UM, STEADY, RC_FIRST>(S, pin, pout); break;
    case RC_K0: launch_flux_one<UM, STEADY, RC_K0>(S, pin, pout); break;
    case RC_K0_PHI: launch_flux_one<UM, STEADY, RC_K0_PHI>(S, pin, pout); break;
    default: launch_flux_one<UM, STEADY, RC_GENERAL>(S, pin, pout); break;
  }
}

int launch_flux(int um, const StageParams &S, const double *pin, double *pout, const int *list = nullptr, int nlist = 0, int part_off = 0) {
  g_sel.list = list; g_sel.n = nlist; g_sel.part_off = part_off;
  if (list && nlist == 0) { C->nparts = part_off; return 0; }
  Span sp(2);
  const bool steady = C->cfg.steady != 0;
  if (um == UM_RESID) launch_flux_rc<UM_RESID, false>(S, pin, pout);
  else if (um == UM_RK) { if (steady) launch_flux_rc<UM_RK, true>(S, pin, pout); else launch_flux_rc<UM_RK, false>(S, pin, pout); }
  else { if (steady) launch_flux_rc<UM_SSPRK, true>(S, pin, pout); else launch_flux_rc<UM_SSPRK, false>(S, pin, pout); }
  C->last_launches++;
  return 0;
}

// ---- one kernel per stage (option "fuse"): tables on first use, then k_stage_fused instead of pass A + pass B
size_t fused_cta_bytes(const FusedMeta &g) { return kStages * fused_stage_bytes(g) + 2 * kStages * sizeof(uint64_t); }

// Pitches of a launch over `tiles`: the maxima over those tiles only, so that a launch over the triangle tiles of a mixed
// mesh needs less shared memory (three CTAs per SM) than one that also holds quadrilateral tiles (two).
FusedMeta fused_group_meta(const std::vector<int> &tiles) {
  const Layout &L = C->L;
  const int F0 = L.g_form == 0 ? 1 : 0;
  FusedMeta g = C->fm;  // pointers
  int n1m = 0, s2m = 0, em = 0, twm = 0, wm = 0, fwm = 0, hfm = 0;
  for (int t : tiles) {
    const int *th = &L.tile_hdr[8 * (size_t)t], *fh = &L.fz_hdr[8 * (size_t)t];
    n1m = std::max(n1m, th[3]); s2m = std::max(s2m, th[3] + fh[1]); em = std::max(em, th[1] + th[5]);
    twm = std::max(twm, (kBlock + th[3] + 7) & ~7); wm = std::max(wm, fh[3]); fwm = std::max(fwm, th[7]); hfm = std::max(hfm, fh[5]);
  }
  g.H1 = (n1m + 1) & ~1; g.HP = (s2m + 1) & ~1; g.E = (em + 1) & ~1; g.TW = twm; g.W = wm; g.CG = std::max(4, wm + F0);
  g.FW = fwm; g.HF = (hfm + 3) & ~3;
  g.XR = std::max(3 + wm + F0, 2 * fwm + (2 * g.HF + kBlock - 1) / kBlock);
  g.ntiles = (int)tiles.size();
  return g;
}

// The launch plan of a stage: at most two launches of k_stage_fused.
//   one rank      [tiles that fit three CTAs per SM] , [the rest]               (either may be empty)
//   several ranks [interior tiles that fit three CTAs per SM] , [boundary tiles (all of them, first) + the other interior tiles]
// Boundary tiles = tiles that read a ghost or hold a cell a peer needs (fz_tile_bnd); only the launch that holds them
// takes part in the in-kernel halo exchange, and the other one runs first, so the peers get a whole interior phase.
int build_fused_plan() {
  const Layout &L = C->L;
  const int nt = L.ntiles;
  int sm_smem = 0;
  CUDA_OK(cudaDeviceGetAttribute(&sm_smem, cudaDevAttrMaxSharedMemoryPerMultiprocessor, C->device));
  const size_t cap = (size_t)sm_smem / 3 - 1024;  // per CTA for three CTAs per SM
  std::vector<unsigned char> is_bnd(nt, 0);
  if (C->nranks > 1) for (int t : L.fz_tile_bnd) is_bnd[t] = 1;
  std::vector<size_t> need(nt);
  for (int t = 0; t < nt; t++) need[t] = fused_cta_bytes(fused_group_meta(std::vector<int>(1, t)));
  std::vector<int> ta, tb;
  double thr = (double)cap;
  for (int iter = 0; iter < 24; iter++, thr *= 0.985) {  // a group's pitches are maxima per array: tighten until group A fits
    ta.clear(); tb.clear();
    for (int t = 0; t < nt; t++) ((!is_bnd[t] && need[t] <= (size_t)thr) ? ta : tb).push_back(t);
    if (ta.empty() || fused_cta_bytes(fused_group_meta(ta)) <= cap) break;
  }
  if (!ta.empty() && fused_cta_bytes(fused_group_meta(ta)) > cap) { tb.insert(tb.end(), ta.begin(), ta.end()); ta.clear(); }
  if (C->opt_fuse == 2 || ta.size() * 8 < (size_t)nt) { tb.insert(tb.end(), ta.begin(), ta.end()); ta.clear(); }  // not worth a launch
  // group B: boundary tiles first
  std::stable_sort(tb.begin(), tb.end(), [&](int x, int y) { return is_bnd[x] > is_bnd[y] || (is_bnd[x] == is_bnd[y] && x < y); });
  int nb = 0;
  for (int t : tb) nb += is_bnd[t];
  C->n_fl = 0;
  for (const std::vector<int> *grp : {&ta, &tb}) {
    if (grp->empty()) continue;
    FusedLaunch &fl = C->fl[C->n_fl++];
    fl = FusedLaunch();
    fl.meta = fused_group_meta(*grp);
    if (fused_cta_bytes(fl.meta) > 227 * 1024) { C->n_fl = 0; return 0; }  // wide stencils: two-pass path
    if (dev_upload(fl.meta.tile_list, *grp)) return 1;
    fl.n_bnd = grp == &tb ? nb : 0;
    fl.tiles = *grp;
  }
  if (getenv("FVS2D_DEBUG"))
    for (int k = 0; k < C->n_fl; k++)
      fprintf(stderr, "[fvs2d] rank %d fused launch %d: %d tiles (%d boundary first) at %zu B/CTA\n", C->rank, k, C->fl[k].meta.ntiles,
              C->fl[k].n_bnd, fused_cta_bytes(C->fl[k].meta));
  return 0;
}

int upload_fused_tables();
int all_ranks_agree(int mine, int &all);

// several ranks: which peer ghost slots every cell of a boundary tile is stored to (HaloP2P::rs_word / rs_ent)
int build_remote_store_tables() {
  const Layout &L = C->L;
  C->d_rs_word = nullptr; C->d_rs_ent = nullptr;
  if (C->nranks == 1 || !C->p2p_ok || C->n_fl == 0) return 0;
  const FusedLaunch &fl = C->fl[C->n_fl - 1];
  std::vector<std::vector<int2>> ent(L.n_own);
  for (size_t k = 0; k < L.peers.size(); k++)
    for (int pos = L.send_ptr[k]; pos < L.send_ptr[k + 1]; pos++)
      ent[L.send_idx[pos]].push_back(make_int2((int)k, C->peer_recv_begin[k] + pos - L.send_ptr[k]));
  std::vector<uint32_t> word((size_t)std::max(1, fl.n_bnd) * kBlock, 0u);
  std::vector<int2> flat;
  for (int j = 0; j < fl.n_bnd; j++) {
    const int c0 = fl.tiles[j] * kBlock;
    for (int c = c0; c < std::min(L.n_own, c0 + kBlock); c++) {
      if (ent[c].empty()) continue;
      NEED(ent[c].size() <= 7 && flat.size() < (1u << 28), "in-kernel halo exchange: a cell is sent to more than 7 peers");
      word[(size_t)j * kBlock + (c - c0)] = (uint32_t)(flat.size() << 3) | (uint32_t)ent[c].size();
      flat.insert(flat.end(), ent[c].begin(), ent[c].end());
    }
  }
  // every sent cell must sit in a boundary tile of this launch
  size_t nsend = L.send_idx.size();
  NEED(flat.size() == nsend, "in-kernel halo exchange: a sent cell lies outside the boundary tiles");
  if (flat.empty()) flat.push_back(make_int2(0, 0));
  if (dev_upload(C->d_rs_word, word) || dev_upload(C->d_rs_ent, flat)) return 1;
  return 0;
}

int ensure_fused_local() {
  C->n_fl = 0;
  if ((C->nranks != 1 && (!C->L.deep || !C->p2p_ok)) || !C->tile_ok || C->recon != RC_K0) return 0;
  if (C->fz_tables == 0 && upload_fused_tables()) return 1;
  if (C->fz_tables != 1) return 0;
  if (build_fused_plan()) return 1;  // (again after a change of the "fuse" option: the tile lists are small)
  if (C->n_fl == 0) return 0;
  if (C->nranks > 1 && !C->L.peers.empty() && C->fl[C->n_fl - 1].n_bnd == 0) { C->n_fl = 0; return 0; }
  if (build_remote_store_tables()) return 1;
  C->fz_state = 1;
  return 0;
}

int ensure_fused() {
  if (C->fz_state) return 0;
  C->fz_state = -1;
  const int rc = ensure_fused_local();
  if (C->nranks > 1 && C->p2p_ok) {  // the peers wait for this rank's flags: every rank runs the fused path or none does
    int all = 0;
    if (all_ranks_agree(rc == 0 && C->fz_state == 1, all)) return 1;
    if (!all) { C->fz_state = -1; C->n_fl = 0; }
  }
  return rc;
}

int upload_fused_tables() {
  C->fz_tables = -1;
  const std::string err = build_fused_tables(C->L);
  if (!err.empty()) return fail("%s", err.c_str());
  const Layout &L = C->L;
  if (L.fz_built != 1 || !L.fz_v2) return 0;
  const int nt = L.ntiles;
  FusedMeta &fm = C->fm;
  std::vector<int> hdr4(16 * (size_t)nt, 0);
  for (int t = 0; t < nt; t++) {
    std::copy(&L.tile_hdr[8 * (size_t)t], &L.tile_hdr[8 * (size_t)t] + 8, &hdr4[16 * (size_t)t]);
    std::copy(&L.fz_hdr[8 * (size_t)t], &L.fz_hdr[8 * (size_t)t] + 6, &hdr4[16 * (size_t)t + 8]);
  }
  std::vector<double> gc;
  fused_coeff_rows(L, (size_t)C->np, gc);
  static_assert(sizeof(double2) == 2 * sizeof(double), "double2 layout");
  const int *dh;
  const double *dgc;
  if (dev_upload(dh, hdr4) || dev_upload(fm.h2_idx, L.fz_h2_idx) || dev_upload(fm.gslot, L.fz_gslot) || dev_upload(dgc, gc) ||
      dev_upload(fm.pack2, L.fz_pack2) || dev_upload(fm.hf, L.fz_hf)) return 1;
  fm.hdr = reinterpret_cast<const int4 *>(dh);
  fm.gc2 = reinterpret_cast<const double2 *>(dgc);
  fm.hc_idx = C->pm.hc_idx; fm.he_idx = C->pm.he_idx; fm.t_bf = C->pm.t_bf;
  fm.tile_list = nullptr;
  C->fz_tables = 1;
  return 0;
}

// one launch of the plan; P2P selects the kernel instance with the in-kernel halo exchange (only the launch that holds the
// boundary tiles needs it: the other instances carry none of its registers)
template <int UM, bool STEADY, int FORM, bool P2P>
int launch_fused_part(FusedLaunch &fl, int k, int part_off, const StageParams &S, const double *pin, double *pout) {
  auto k3 = k_stage_fused<UM, STEADY, FORM, 3, P2P>;
  auto k2 = k_stage_fused<UM, STEADY, FORM, 2, P2P>;  // more registers, for launches that only fit two CTAs per SM anyway
  const size_t smem = fused_cta_bytes(fl.meta);
  const void *key = (const void *)k3;
  ensure_smem_attr(k3, smem);
  ensure_smem_attr(k2, smem);
  if (fl.attr_key != key) {  // occupancy of this launch configuration, once per kernel instance
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fl.per3, k3, kPipeThreads, smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fl.per2, k2, kPipeThreads, smem);
    fl.attr_key = key;
    if (getenv("FVS2D_DEBUG")) fprintf(stderr, "[fvs2d] k_stage_fused launch %d: %zu B smem/CTA, CTAs/SM %d (128 regs) / %d%s\n", k, smem, fl.per3, fl.per2, P2P ? ", in-kernel halo exchange" : "");
  }
  const bool use3 = fl.per3 >= 3 && (C->opt_ctas == 0 || C->opt_ctas >= 3);
  int per_sm = std::max(1, use3 ? fl.per3 : fl.per2);
  if (C->opt_ctas > 0) per_sm = std::min(per_sm, C->opt_ctas);
  const int grid = std::min(fl.meta.ntiles, C->nsm * per_sm);
  HaloP2P hx{};
  if (P2P) {
    const int which = pout == C->p_buf[0] ? 0 : 1;
    hx.n_peers = (int)C->L.peers.size();
    hx.n_bnd = fl.n_bnd; hx.n_bnd_ctas = std::min(fl.n_bnd, grid);
    hx.stage = S.stage; hx.clk = C->clk;
    for (int q = 0; q < hx.n_peers; q++) {
      hx.peer_out[q] = C->peer_p[q][which];
      hx.peer_np[q] = C->peer_np[q];
      hx.peer_flag[q] = C->peer_flag[q];
      hx.my_flag[q] = C->flags + C->L.peers[q];
    }
    hx.rs_word = C->d_rs_word; hx.rs_ent = C->d_rs_ent; hx.done_ctr = C->done_ctr; hx.timed_out = C->p2p_timed_out;
  }
  if (grid > 0)
    launch_chain(use3 ? k3 : k2, dim3(grid), dim3(kPipeThreads), smem, C->dm, fl.meta, C->phys, S, pin, C->bc, C->q, C->f, pout, C->dtl,
                                                          C->partial + 4 * (size_t)part_off, hx);
  C->last_launches++;
  return grid;
}

template <int UM, bool STEADY>
void launch_fused_form(const StageParams &S, const double *pin, double *pout) {
  const bool gg = C->L.g_form == 0;
  int part_off = 0;
  for (int k = 0; k < C->n_fl; k++) {
    FusedLaunch &fl = C->fl[k];
    const bool p2p = C->nranks > 1 && fl.n_bnd > 0;
    if (p2p) part_off += gg ? launch_fused_part<UM, STEADY, 0, true>(fl, k, part_off, S, pin, pout) : launch_fused_part<UM, STEADY, 1, true>(fl, k, part_off, S, pin, pout);
    else part_off += gg ? launch_fused_part<UM, STEADY, 0, false>(fl, k, part_off, S, pin, pout) : launch_fused_part<UM, STEADY, 1, false>(fl, k, part_off, S, pin, pout);
  }
  C->nparts = part_off;
}

int launch_fused(int um, const StageParams &S, const double *pin, double *pout) {
  Span sp(2);
  const bool steady = C->cfg.steady != 0;
  if (um == UM_RK) { if (steady) launch_fused_form<UM_RK, true>(S, pin, pout); else launch_fused_form<UM_RK, false>(S, pin, pout); }
  else { if (steady) launch_fused_form<UM_SSPRK, true>(S, pin, pout); else launch_fused_form<UM_SSPRK, false>(S, pin, pout); }
  return 0;
}

int launch_bc(int stage) {
  if (C->L.nbf == 0) return 0;
  const bool time_dep = C->cfg.lvortex != 0;
  if (!time_dep && C->bc_static_done) return 0;
  Span sp(0);
  launch_chain(k_bc_state, dim3(cdiv(C->L.nbf, 128)), dim3(128), 0, C->dm, C->phys, C->clk, stage, C->bc);
  C->last_launches++;
  C->bc_static_done = true;
  return 0;
}

// ---- in-kernel halo exchange of the fused path: map the peers' state buffers and flag words (CUDA IPC) ----------
constexpr int kMaxRanks = 16;
struct P2PInfo {
  cudaIpcMemHandle_t h[3];   // primitive-state buffers 0 / 1, flag words
  long long off[3];          // offset of the array inside the exported allocation (small cudaMallocs are sub-allocated)
  int np, ok;
  int recv_begin[kMaxRanks], recv_count[kMaxRanks];  // by sender rank: where its cells land in this rank's numbering
};

void close_p2p() {
  for (void *b : C->ipc_open) cudaIpcCloseMemHandle(b);
  C->ipc_open.clear();
  C->p2p_ok = false;
}

int all_ranks_agree(int mine, int &all) {  // min over the communicator (also a barrier)
  int *d = nullptr;
  CUDA_OK(cudaMalloc(&d, sizeof(int)));
  CUDA_OK(cudaMemcpyAsync(d, &mine, sizeof(int), cudaMemcpyHostToDevice, C->st));
  NCCL_OK(g_nccl.AllReduce(d, d, 1, ncclInt, ncclMin, C->comm, C->st));
  CUDA_OK(cudaMemcpyAsync(&all, d, sizeof(int), cudaMemcpyDeviceToHost, C->st));
  CUDA_OK(cudaStreamSynchronize(C->st));
  cudaFree(d);
  return 0;
}

int setup_p2p() {
  close_p2p();
  C->epoch_total = 0;
  if (C->nranks == 1) return 0;
  const Layout &L = C->L;
  const int nr = C->nranks, npeer = (int)L.peers.size();
  if (dev_alloc(C->flags, (size_t)kMaxRanks) || dev_alloc(C->done_ctr, 1) || dev_alloc(C->p2p_timed_out, 1)) return 1;
  CUDA_OK(cudaMemset(C->p2p_timed_out, 0, sizeof(int)));
  CUDA_OK(cudaMemset(C->flags, 0, kMaxRanks * sizeof(unsigned)));
  CUDA_OK(cudaMemset(C->done_ctr, 0, sizeof(unsigned)));
  P2PInfo mine{};
  mine.ok = npeer <= kMaxPeers && nr <= kMaxRanks && !getenv("FVS2D_NO_P2P");
  mine.np = C->np;
  {
    typedef int (*range_fn)(unsigned long long *, size_t *, unsigned long long);  // cuMemGetAddressRange (driver API, no link dependency)
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn) { mine.ok = 0; cudaGetLastError(); }
    void *arr[3] = {C->p_buf[0], C->p_buf[1], C->flags};
    for (int k = 0; k < 3 && mine.ok; k++) {
      unsigned long long base = 0;
      size_t size = 0;
      if (((range_fn)fn)(&base, &size, (unsigned long long)arr[k]) != 0) { mine.ok = 0; break; }
      mine.off[k] = (long long)((unsigned long long)arr[k] - base);
      if (cudaIpcGetMemHandle(&mine.h[k], (void *)base) != cudaSuccess) { mine.ok = 0; cudaGetLastError(); }
    }
  }
  for (int k = 0; k < npeer && L.peers[k] < kMaxRanks; k++) { mine.recv_begin[L.peers[k]] = L.recv_begin[k]; mine.recv_count[L.peers[k]] = L.recv_count[k]; }
  std::vector<P2PInfo> all(nr);
  {
    char *d = nullptr;
    CUDA_OK(cudaMalloc(&d, sizeof(P2PInfo) * (size_t)(nr + 1)));
    CUDA_OK(cudaMemcpyAsync(d + sizeof(P2PInfo) * (size_t)nr, &mine, sizeof mine, cudaMemcpyHostToDevice, C->st));
    NCCL_OK(g_nccl.AllGather(d + sizeof(P2PInfo) * (size_t)nr, d, sizeof(P2PInfo), ncclChar, C->comm, C->st));
    CUDA_OK(cudaMemcpyAsync(all.data(), d, sizeof(P2PInfo) * (size_t)nr, cudaMemcpyDeviceToHost, C->st));
    CUDA_OK(cudaStreamSynchronize(C->st));
    cudaFree(d);
  }
  int ok = 1;
  for (int r = 0; r < nr; r++) ok = ok && all[r].ok;
  for (int k = 0; k < npeer && ok; k++) {
    const P2PInfo &pi = all[L.peers[k]];
    void *base[3] = {nullptr, nullptr, nullptr};
    for (int a = 0; a < 3 && ok; a++) {
      if (cudaIpcOpenMemHandle(&base[a], pi.h[a], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); break; }
      C->ipc_open.push_back(base[a]);
    }
    if (!ok) break;
    C->peer_p[k][0] = reinterpret_cast<double2 *>((char *)base[0] + pi.off[0]);
    C->peer_p[k][1] = reinterpret_cast<double2 *>((char *)base[1] + pi.off[1]);
    C->peer_flag[k] = reinterpret_cast<unsigned *>((char *)base[2] + pi.off[2]) + C->rank;
    C->peer_np[k] = pi.np;
    C->peer_recv_begin[k] = pi.recv_begin[C->rank];
    if (pi.recv_count[C->rank] != L.send_ptr[k + 1] - L.send_ptr[k]) return fail("halo plan mismatch between ranks %d and %d", C->rank, L.peers[k]);
  }
  int agreed = 0;
  if (all_ranks_agree(ok, agreed)) return 1;
  if (!agreed) { close_p2p(); return 0; }  // no peer access on this box: the two-pass path with NCCL send/recv
  C->p2p_ok = true;
  return 0;
}

// spins (one warp) until every peer has delivered stage `epoch`: the ghost slots are complete when this kernel retires
struct PeerList { int n; int r[kMaxPeers]; };
__global__ void k_wait_peers(const unsigned *__restrict__ flags, const PeerList pl, unsigned epoch, int *timed_out) {
  if ((int)threadIdx.x < pl.n) wait_flag(flags + pl.r[threadIdx.x], epoch, timed_out);
}

// one residual evaluation's worth of pass A (+ halo) for state p
int pass_a(const double *p) {
  if (launch_gradient(p)) return 1;
  if (C->nranks > 1 && C->recon != RC_FIRST) {
    Span sp(0);
    HaloItem it[2] = {{C->g, 8, 1}, {C->phi, 1, 0}};
    if (halo_exchange(it, C->cfg.limiter > 0 ? 2 : 1)) return 1;
  }
  return 0;
}

}  // namespace

// =================================================================================================
extern "C" {

const char *fvs2d_gpu_last_error(void) { return g_err.c_str(); }

int fvs2d_gpu_init(const fvs2d_config *cfg, int device) {
  NEED(cfg != nullptr, "fvs2d_gpu_init: null config");
  if (C) fvs2d_gpu_finalize();
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return fail("fvs2d_gpu_init: no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
  if (device < 0) {
    const char *lr = getenv("LOCAL_RANK");
    device = lr ? atoi(lr) % ndev : 0;
  }
  NEED(device < ndev, "fvs2d_gpu_init: device index out of range");
  C = new Ctx();
  C->cfg = *cfg;
  C->device = device;
  if (const char *ev = getenv("FVS2D_FUSE")) C->opt_fuse = atoi(ev);  // "fuse" option for hosts that cannot call set_option
  CUDA_OK(cudaSetDevice(device));
  CUDA_OK(cudaStreamCreateWithFlags(&C->st, cudaStreamNonBlocking));
  CUDA_OK(cudaDeviceGetAttribute(&C->nsm, cudaDevAttrMultiProcessorCount, device));
  CUDA_OK(cudaEventCreate(&C->ev0));
  CUDA_OK(cudaEventCreate(&C->ev1));
  // scheme validation: the stop conditions of src/input.f90:181-277 and the limiter/LSQ coupling
  NEED(cfg->grad_method >= 1 && cfg->grad_method <= 3, "check cell-center gradient scheme in input file");
  NEED(cfg->limiter >= 0 && cfg->limiter <= 3, "check gradient limiter scheme in input file");
  NEED(cfg->recon >= 1 && cfg->recon <= 3, "check face reconstruction scheme in input file");
  NEED(cfg->flux == 1, "check inviscid flux discretization scheme in input file");
  if (cfg->limiter > 0 && cfg->recon != 1)
    NEED(cfg->grad_method == 3, "gradient limiter needs the LSQ stencil (src/gradient_limiter.f90:54-58): use grad_method 3");
  if (rk_setup()) return 1;
  const double kap = cfg->recon == 3 ? cfg->umuscl_cst : 0.0;  // src/input.f90:248-254
  C->recon = recon_mode(*cfg);
  Phys &P = C->phys;
  P.gamma = cfg->gamma; P.kappa = kap; P.cfl = cfg->cfl_user;
  P.gm1 = cfg->gamma - 1.0; P.gog = cfg->gamma / (cfg->gamma - 1.0);
  {
    const double x = 2.0 / P.gm1, n = std::nearbyint(x);
    P.pow2n = (n >= 1.0 && n <= 64.0 && std::fabs(x - n) < 1e-12) ? (int)n : 0;
    P.pad = 0;
  }
  for (int i = 0; i < 4; i++) { P.pinf[i] = cfg->pvar_inf[i]; P.vinf[i] = cfg->vortex_inf[i]; }
  P.vpos[0] = cfg->vortex_pos[0]; P.vpos[1] = cfg->vortex_pos[1]; P.vkap = cfg->vortex_kappa;
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) P.mms[i][j] = cfg->mms_c[i][j];
  P.lvortex = cfg->lvortex; P.limiter = cfg->limiter;
  C->inited = true;
  return 0;
}

int fvs2d_gpu_comm_unique_id(char id[128]) {
  if (!g_nccl.load(g_err)) return 1;
  ncclUniqueId u;
  NCCL_OK(g_nccl.GetUniqueId(&u));
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  memcpy(id, &u, 128);
  return 0;
}

int fvs2d_gpu_comm_init(int rank, int nranks, const char id[128]) {
  NEED(C && C->inited, "fvs2d_gpu_comm_init: call fvs2d_gpu_init first");
  NEED(!C->has_mesh, "fvs2d_gpu_comm_init: must precede fvs2d_gpu_set_mesh");
  NEED(nranks >= 1 && rank >= 0 && rank < nranks, "fvs2d_gpu_comm_init: bad rank");
  if (nranks == 1) { C->rank = 0; C->nranks = 1; return 0; }
  if (!g_nccl.load(g_err)) return 1;
  ncclUniqueId u;
  memcpy(&u, id, 128);
  NCCL_OK(g_nccl.CommInitRank(&C->comm, nranks, u, rank));
  C->rank = rank; C->nranks = nranks;
  return 0;
}

}  // extern "C"

namespace {
int build_partition();
int device_upload();

// ---- Hilbert keys + stable radix sort on the device (set-up only; HilbertSorter of host_mesh.hpp) -------------------
// The same key arithmetic as the host loop of hilbert_order_raw (sums in node order, IEEE division, no contraction
// possible in (s - x0) * scale), CUB's LSD radix sort is stable like the host's: the permutation is identical.
__global__ void k_hilbert_keys(const int nc, const int *__restrict__ cptr, const int *__restrict__ cnode, const double *__restrict__ xn,
                               const double *__restrict__ yn, const int xs, const HilbertFrame f, uint64_t *__restrict__ key,
                               int *__restrict__ val) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nc) return;
  double sx = 0, sy = 0;
  const int s0 = cptr[i], s1 = cptr[i + 1];
  for (int s = s0; s < s1; s++) { const size_t v = (size_t)cnode[s] * xs; sx = __dadd_rn(sx, xn[v]); sy = __dadd_rn(sy, yn[v]); }
  const double nv = (double)(s1 - s0);
  sx = __ddiv_rn(sx, nv); sy = __ddiv_rn(sy, nv);
  const uint32_t ix = (uint32_t)__dmul_rn(__dsub_rn(sx, f.x0), f.scale_x), iy = (uint32_t)__dmul_rn(__dsub_rn(sy, f.y0), f.scale_y);
  key[i] = hilbert_d(ix, iy, f.bits);
  val[i] = i;
}

bool device_hilbert_sort(const HilbertFrame &f, int nc, const int *cptr, const int *cnode, const double *xn, const double *yn, int xs,
                         std::vector<int> &perm) {
  if (!C || !C->inited || getenv("FVS2D_HOST_SORT")) return false;
  const size_t nslots = (size_t)cptr[nc];
  int nn = 0;  // node count = largest id + 1 (only the entries the cells use are read)
#pragma omp parallel for schedule(static) reduction(max : nn)
  for (size_t s = 0; s < nslots; s++) nn = std::max(nn, cnode[s] + 1);
  int *d_cptr = nullptr, *d_cnode = nullptr, *d_val = nullptr, *d_val2 = nullptr;
  double *d_x = nullptr, *d_y = nullptr;
  uint64_t *d_key = nullptr, *d_key2 = nullptr;
  void *d_tmp = nullptr;
  bool ok = true;
  auto A = [&](void **p, size_t bytes) { if (ok && cudaMalloc(p, std::max<size_t>(bytes, 16)) != cudaSuccess) { ok = false; cudaGetLastError(); } };
  const size_t nxy = xs == 2 ? 2 * (size_t)nn : (size_t)nn;
  A((void **)&d_cptr, ((size_t)nc + 1) * 4); A((void **)&d_cnode, nslots * 4); A((void **)&d_x, nxy * 8);
  if (xs == 1) A((void **)&d_y, nxy * 8);
  A((void **)&d_key, (size_t)nc * 8); A((void **)&d_key2, (size_t)nc * 8); A((void **)&d_val, (size_t)nc * 4); A((void **)&d_val2, (size_t)nc * 4);
  size_t tmp_bytes = 0;
  if (ok) {
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_key, d_key2, d_val, d_val2, nc, 0, 2 * f.bits, C->st);
    A(&d_tmp, tmp_bytes);
  }
  if (ok) {
    ok = cudaMemcpyAsync(d_cptr, cptr, ((size_t)nc + 1) * 4, cudaMemcpyHostToDevice, C->st) == cudaSuccess &&
         cudaMemcpyAsync(d_cnode, cnode, nslots * 4, cudaMemcpyHostToDevice, C->st) == cudaSuccess &&
         cudaMemcpyAsync(d_x, xn, nxy * 8, cudaMemcpyHostToDevice, C->st) == cudaSuccess &&
         (xs == 2 || cudaMemcpyAsync(d_y, yn, nxy * 8, cudaMemcpyHostToDevice, C->st) == cudaSuccess);
  }
  if (ok) {
    k_hilbert_keys<<<cdiv(nc, 256), 256, 0, C->st>>>(nc, d_cptr, d_cnode, d_x, xs == 2 ? d_x + 1 : d_y, xs, f, d_key, d_val);
    ok = cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_key, d_key2, d_val, d_val2, nc, 0, 2 * f.bits, C->st) == cudaSuccess;
    perm.resize(nc);
    ok = ok && cudaMemcpyAsync(perm.data(), d_val2, (size_t)nc * 4, cudaMemcpyDeviceToHost, C->st) == cudaSuccess &&
         cudaStreamSynchronize(C->st) == cudaSuccess;
  }
  for (void *p : {(void *)d_cptr, (void *)d_cnode, (void *)d_x, (void *)d_y, (void *)d_key, (void *)d_key2, (void *)d_val, (void *)d_val2, d_tmp})
    if (p) cudaFree(p);
  if (!ok) cudaGetLastError();
  return ok;
}
const HilbertSorter g_device_sorter = device_hilbert_sort;

// host half of set_mesh: connectivity, geometry, gradient operator, renumbering, layout (no CUDA calls)
int host_build(int nnodes, int ntri, int nquad, const double *node_xy, const int *cell_ptr, const int *cell_node,
               int nb, const int *b_ncells, const int *b_type, const int *b_cell) {
  NEED(nnodes > 0 && ntri >= 0 && nquad >= 0 && ntri + nquad > 0, "fvs2d_gpu_set_mesh: empty mesh");
  NEED(node_xy && cell_ptr && cell_node, "fvs2d_gpu_set_mesh: null array");
  NEED(nb == 0 || (b_ncells && b_type && b_cell), "fvs2d_gpu_set_mesh: null boundary array");
  C->has_mesh = C->has_state = false;
  C->bc_static_done = false;
  {  // launchers such as torchrun export OMP_NUM_THREADS=1; the one-off pre-processing takes its share of the cores
    const int hw = (int)std::thread::hardware_concurrency();
    omp_set_num_threads(std::max(1, std::min(32, hw / std::max(1, C->nranks))));
  }
  HostMesh &m = C->mesh;
  m = HostMesh();
  const int nc = ntri + nquad;
  NEED(cell_ptr[0] == 0 && cell_ptr[nc] == 3 * ntri + 4 * nquad, "fvs2d_gpu_set_mesh: cell_ptr inconsistent with ncells_tri/ncells_quad");
  for (int ib = 0; ib < nb; ib++)
    NEED(b_type[ib] == FVS2D_BC_FREESTREAM || b_type[ib] == FVS2D_BC_SLIP_WALL || b_type[ib] == FVS2D_BC_DIRICHLET,
         "Boundary condition not implemented (src/residual.f90:206-216)");
  C->sub = SubMesh();
  GlobalInfo &gi = C->gi;
  gi = GlobalInfo();
  // Several ranks: partition-local pre-processing -- only this rank's cells and two rings around them are ever connected
  // (extract_submesh).  The least-squares stencil over face neighbours completes boundary stencils from the nearest
  // centroids of the WHOLE mesh (src/gradient_lsq.f90:98-123), so that scheme keeps the global build.
  const double t_begin = omp_get_wtime();
  auto lap = [&](const char *what) {
    if (getenv("FVS2D_DEBUG")) fprintf(stderr, "[fvs2d] rank %d host build: %-28s %.2f s since start\n", C->rank, what, omp_get_wtime() - t_begin);
  };
  const bool local_build = C->nranks > 1 && !(C->cfg.grad_method == 3 && C->cfg.lsq_stencil == 0) && !getenv("FVS2D_GLOBAL_BUILD");
  std::string err;
  if (local_build) {
    err = extract_submesh(nnodes, ntri, nquad, node_xy, cell_ptr, cell_node, nb, b_ncells, b_type, b_cell, C->rank, C->nranks, 2, C->sub,
                          &g_device_sorter);
    if (!err.empty()) return fail("%s", err.c_str());
    m = std::move(C->sub.m);
    C->sub.m = HostMesh();
    lap("submesh extracted");
  } else {
    m.nnodes = nnodes; m.ntri = ntri; m.nquad = nquad; m.ncells = nc;
    m.xn.resize(nnodes); m.yn.resize(nnodes);
    for (int i = 0; i < nnodes; i++) { m.xn[i] = node_xy[2 * (size_t)i]; m.yn[i] = node_xy[2 * (size_t)i + 1]; }
    m.cptr.assign(cell_ptr, cell_ptr + nc + 1);
    m.cnode.assign(cell_node, cell_node + m.cptr[nc]);
    m.nb = nb;
    m.b_ncells.assign(b_ncells, b_ncells + nb);
    m.b_type.assign(b_type, b_type + nb);
    size_t nbc = 0;
    for (int ib = 0; ib < nb; ib++) nbc += b_ncells[ib];
    m.b_cell.assign(b_cell, b_cell + nbc);
  }
  err = build_mesh(m);
  if (!err.empty()) return fail("%s", err.c_str());
  lap("connectivity + geometry");
  err = build_gradient(m, C->cfg.grad_method, C->cfg.lsq_stencil, C->cfg.lsq_pow, C->grad);
  if (!err.empty()) return fail("%s", err.c_str());
  lap("gradient stencils");
  if (!local_build) {
    gi.nnodes = m.nnodes; gi.ncells = m.ncells; gi.nedges = m.nedges; gi.nedges_intr = m.nedges_intr; gi.nedges_bndr = m.nedges_bndr;
    gi.ncells_intr = m.ncells_intr; gi.ncells_bndr = m.ncells_bndr;
    gi.heff1 = m.heff1; gi.heff2 = m.heff2; gi.vol_sum = m.vol_sum; gi.vol_green = m.vol_green;
    gi.xy_cell0[0] = m.xc[0]; gi.xy_cell0[1] = m.yc[0];
  } else {
    // this rank's share of the global counts and sums (an edge counts for the owner of its c1 cell; owned cells have no
    // cut faces); fvs2d_gpu_set_mesh adds the shares up over the communicator
    const SubMesh &sb = C->sub;
    gi.nnodes = nnodes; gi.ncells = nc;
    gi.xy_cell0[0] = sb.xy_cell0[0]; gi.xy_cell0[1] = sb.xy_cell0[1];
    auto owned = [&](int mc) { return sb.new_id[mc] >= sb.b0 && sb.new_id[mc] < sb.b1; };
    long long ne = 0, neb = 0, ci = 0, cb = 0;
    double vs = 0, vsq = 0, vg = 0;
    for (int ie = 0; ie < m.nedges; ie++) if (owned(m.ec1[ie])) { ne++; neb += m.ec2[ie] < 0; }
    for (int ic = 0; ic < m.ncells; ic++) {
      if (!owned(ic)) continue;
      bool intr = true;
      double v = 0;
      for (int sl = m.cptr[ic]; sl < m.cptr[ic + 1]; sl++) {
        intr = intr && m.nghbre[sl] >= 0;
        const int je = m.cedge[sl];
        const EdgeGeom eg = edge_geom(m, je);
        v += eg.nx * (m.ec1[je] == ic ? 1.0 : -1.0) * eg.x * eg.a;
      }
      (intr ? ci : cb)++;
      vs += m.vol[ic]; vsq += std::sqrt(m.vol[ic]); vg += v;
    }
    gi.nedges = (int)ne; gi.nedges_bndr = (int)neb; gi.nedges_intr = (int)(ne - neb); gi.ncells_intr = (int)ci; gi.ncells_bndr = (int)cb;
    gi.vol_sum = vs; gi.heff2 = vsq; gi.vol_green = vg;   // sums; heff1 / heff2 are formed after the reduction
    gi.partial_sums = true;
  }
  const int rc = build_partition();
  lap("renumbering + layout");
  return rc;
}

// renumbering + this rank's layout from the mesh and the gradient operator (again after fvs2d_gpu_set_lsq)
int build_partition() {
  HostMesh &m = C->mesh;
  std::string err;
  CellNumbering num;
  if (m.partial) {
    const SubMesh &sb = C->sub;
    num.nc_global = sb.nc_global;
    num.new_id = sb.new_id;
    num.orig = sb.orig.data();
    num.cuts = sb.cuts;
    num.order.resize(m.ncells);
    for (int i = 0; i < m.ncells; i++) num.order[i] = i;
    std::sort(num.order.begin(), num.order.end(), [&](int x, int y) { return sb.new_id[x] < sb.new_id[y]; });
  } else {
    num.nc_global = m.ncells;
    hilbert_order(m, num.order, &g_device_sorter);
    num.cuts = partition_cuts(num.order, m.ntri, C->nranks);
    num.new_id.resize(m.ncells);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < m.ncells; i++) num.new_id[num.order[i]] = i;
  }
  // several ranks + the fused stage kernel: one more ghost layer (the stencils of the face-neighbour ghosts); the
  // environment variable lets the CPU-side verification (fvs2d_host_build) ask for it
  // (decided before the tables exist: if they turn out not to fit, the two-pass path runs on the deeper layout)
  const bool deep = C->nranks > 1 && ((C->opt_fuse != 0 && C->recon == RC_K0 && !getenv("FVS2D_NO_P2P")) || getenv("FVS2D_DEEP_GHOSTS") != nullptr);
  err = build_layout(m, C->grad, num, C->rank, C->nranks, C->L, deep);
  if (!err.empty()) return fail("%s", err.c_str());
  C->has_mesh = true;
  return 0;
}
}  // namespace

extern "C" {

int fvs2d_host_build(const fvs2d_config *cfg, int rank, int nranks, int nnodes, int ntri, int nquad, const double *node_xy,
                     const int *cell_ptr, const int *cell_node, int nb, const int *b_ncells, const int *b_type, const int *b_cell) {
  NEED(cfg != nullptr, "fvs2d_host_build: null config");
  NEED(nranks >= 1 && rank >= 0 && rank < nranks, "fvs2d_host_build: bad rank");
  if (C) fvs2d_gpu_finalize();
  C = new Ctx();
  C->cfg = *cfg;
  C->recon = recon_mode(*cfg);
  if (const char *ev = getenv("FVS2D_FUSE")) C->opt_fuse = atoi(ev);
  C->rank = rank; C->nranks = nranks;
  return host_build(nnodes, ntri, nquad, node_xy, cell_ptr, cell_node, nb, b_ncells, b_type, b_cell);
}

int fvs2d_gpu_set_mesh(int nnodes, int ntri, int nquad, const double *node_xy, const int *cell_ptr, const int *cell_node,
                       int nb, const int *b_ncells, const int *b_type, const int *b_cell) {
  NEED(C && C->inited, "fvs2d_gpu_set_mesh: call fvs2d_gpu_init first");
  free_device();
  if (host_build(nnodes, ntri, nquad, node_xy, cell_ptr, cell_node, nb, b_ncells, b_type, b_cell)) return 1;
  return device_upload();
}

/* Replaces the least-squares operator built by fvs2d_gpu_set_mesh with the caller's (the reference's public `lsq` table,
 * src/gradient_lsq.f90:16-27: lsq(ic)%ncells, %cell, %w, %coef), so that the Fortran host's own coefficients -- kd-tree
 * tie-breaks included -- are used bit for bit.  CSR over the cells in the ORIGINAL numbering, 0-based cell ids. */
int fvs2d_gpu_set_lsq(const int *ptr, const int *cell, const double *w, const double *coef) {
  NEED(C && C->inited && C->has_mesh, "fvs2d_gpu_set_lsq: call fvs2d_gpu_init and fvs2d_gpu_set_mesh first");
  NEED(C->cfg.grad_method == 3, "fvs2d_gpu_set_lsq: the run does not use the least-squares gradient (grad_method 3)");
  NEED(ptr && cell && w && coef, "fvs2d_gpu_set_lsq: null array");
  const int nc = C->gi.ncells;   // the caller's table covers the whole mesh, original numbering
  NEED(ptr[0] == 0, "fvs2d_gpu_set_lsq: ptr[0] must be 0");
  for (int i = 0; i < nc; i++) NEED(ptr[i + 1] >= ptr[i] && ptr[i + 1] - ptr[i] <= kMaxStencil, "fvs2d_gpu_set_lsq: bad ptr");
  for (int64_t k = 0; k < ptr[nc]; k++) NEED(cell[k] >= 0 && cell[k] < nc, "fvs2d_gpu_set_lsq: cell id out of range");
  GradOp g;
  g.form = 1; g.method = 3; g.lsq_pow = C->cfg.lsq_pow;
  const HostMesh &m = C->mesh;
  const std::vector<int> &orig = C->sub.orig;   // empty: m is the whole mesh; else m-cell -> original id (ascending)
  g.ptr.assign((size_t)m.ncells + 1, 0);
  for (int mc = 0; mc < m.ncells; mc++) {
    const int o = orig.empty() ? mc : orig[mc];
    int n = 0;
    for (int k = ptr[o]; k < ptr[o + 1]; k++) n += orig.empty() || std::binary_search(orig.begin(), orig.end(), cell[k]);
    // (a submesh cell whose stencil leaves the submesh lies in its outermost ring: its gradient is never formed)
    g.ptr[mc + 1] = g.ptr[mc] + n;
  }
  const int64_t n = g.ptr[m.ncells];
  g.idx.resize(n); g.user_cx.resize(n); g.user_cy.resize(n);
#pragma omp parallel for schedule(static)
  for (int mc = 0; mc < m.ncells; mc++) {
    const int o = orig.empty() ? mc : orig[mc];
    int64_t e = g.ptr[mc];
    for (int k = ptr[o]; k < ptr[o + 1]; k++) {
      int j = cell[k];
      if (!orig.empty()) {
        const auto it = std::lower_bound(orig.begin(), orig.end(), j);
        if (it == orig.end() || *it != j) continue;
        j = (int)(it - orig.begin());
      }
      g.idx[e] = j;
      g.user_cx[e] = coef[2 * (size_t)k] * w[k];      // grad = sum coef(:,k) * (p_j - p_i) * w(k), src/gradient_lsq.f90:397
      g.user_cy[e] = coef[2 * (size_t)k + 1] * w[k];
      e++;
    }
  }
  C->grad = std::move(g);
  free_device();
  C->has_mesh = C->has_state = false;
  if (build_partition()) return 1;
  return device_upload();
}

}  // extern "C"

namespace {
int device_upload() {
  C->has_mesh = false;
  // ---- upload
  const Layout &L = C->L;
  DevMesh &d = C->dm;
  d.n_own = L.n_own; d.n_loc = L.n_loc; d.nbf = L.nbf;
  C->np = d.np = (L.n_loc + 31) / 32 * 32;
  C->nblocks = cdiv(L.n_own, kBlock);
  if (dev_upload(d.f_off, L.f_off) || dev_upload(d.f_nbr, L.f_nbr) || dev_upload(d.f_edge, L.f_edge)) return 1;
  {  // pair-interleaved geometry (16-byte gathers); per-cell arrays padded to the SoA pitch
    std::vector<double2> exy(L.nedges), enxy(L.nedges), xy(C->np, make_double2(0.0, 0.0));
    for (int e = 0; e < L.nedges; e++) { exy[e] = make_double2(L.ex[e], L.ey[e]); enxy[e] = make_double2(L.enx[e], L.eny[e]); }
    for (int i = 0; i < L.n_loc; i++) xy[i] = make_double2(L.xc[i], L.yc[i]);
    std::vector<double> vol(L.vol);
    vol.resize(C->np, 1.0);
    std::vector<double> ivol(vol.size());
    for (size_t i = 0; i < vol.size(); i++) ivol[i] = 1.0 / vol[i];
    if (dev_upload(d.exy, exy) || dev_upload(d.enxy, enxy) || dev_upload(d.ea, L.ea) || dev_upload(d.xy, xy) || dev_upload(d.vol, vol) ||
        dev_upload(d.ivol, ivol)) return 1;
  }
  C->tile_ok = L.tile_hc_max >= 0;
  if (C->tile_ok) {
    PipeMeta &pmeta = C->pm;
    const int *hdr; const uint32_t *tp;
    if (dev_upload(pmeta.hc_idx, L.tile_hc_idx) || dev_upload(pmeta.he_idx, L.tile_he_idx) || dev_upload(hdr, L.tile_hdr) ||
        dev_upload(tp, L.t_pack) || dev_upload(pmeta.t_bf, L.t_bf)) return 1;
    pmeta.hdr = reinterpret_cast<const int4 *>(hdr);
    pmeta.t_pack = tp;
    pmeta.S = (kBlock + L.tile_hc_max + 1) & ~1;
    pmeta.E = (L.tile_e_max + 1) & ~1;
    pmeta.ntiles = L.ntiles;
    if (kStages * pipe_stage_bytes<RC_GENERAL>(pmeta.S, pmeta.E) + 64 > 220 * 1024) C->tile_ok = false;  // pathological numbering: direct-gather kernel
  }
  if (dev_upload(d.g_off, L.g_off) || dev_upload(d.g_idx, L.g_idx) || dev_upload(d.g_cx, L.g_cx) || dev_upload(d.g_cy, L.g_cy)) return 1;
  if (dev_upload(d.c0x, L.c0x) || dev_upload(d.c0y, L.c0y)) return 1;
  if (dev_upload(d.bf_type, L.bf_type) || dev_upload(d.bf_edge, L.bf_edge)) return 1;
  if (dev_upload(d.is_intr, L.is_intr) || dev_upload(d.orig_id, L.orig_id)) return 1;
  const size_t np = C->np;
  if (dev_alloc(C->q, 4 * np) || dev_alloc(C->f, 4 * np) || dev_alloc(C->pa, 4 * np) || dev_alloc(C->pb, 4 * np)) return 1;
  C->p_buf[0] = C->pa; C->p_buf[1] = C->pb;
  if (dev_alloc(C->g, 8 * np) || dev_alloc(C->phi, np)) return 1;
  if (C->cfg.steady && dev_alloc(C->dtl, np)) return 1;
  if (dev_alloc(C->bc, 4 * (size_t)std::max(1, L.nbf))) return 1;
  if (dev_alloc(C->clk, 1)) return 1;
  if (C->graph_exec) { cudaGraphExecDestroy(C->graph_exec); C->graph_exec = nullptr; }
  if (dev_alloc(C->partial, (size_t)(2 * C->nblocks + 1) * 4) || dev_alloc(C->vpartial, (size_t)C->nblocks * 13) || dev_alloc(C->vbest, (size_t)C->nblocks) || dev_alloc(C->vbest_loc, (size_t)C->nblocks)) return 1;
  CUDA_OK(cudaMemset(C->q, 0, 4 * np * 8));
  CUDA_OK(cudaMemset(C->f, 0, 4 * np * 8));
  CUDA_OK(cudaMemset(C->pa, 0, 4 * np * 8));
  CUDA_OK(cudaMemset(C->pb, 0, 4 * np * 8));
  CUDA_OK(cudaMemset(C->g, 0, 8 * np * 8));
  {  // phi == 1 unless the limiter kernel overwrites it (src/gradient_limiter.f90:36-39)
    std::vector<double> ones(np, 1.0);
    CUDA_OK(cudaMemcpy(C->phi, ones.data(), np * 8, cudaMemcpyHostToDevice));
  }
  if (C->nranks > 1) {
    if (dev_upload(C->d_tile_int, L.tile_int) || dev_upload(C->d_tile_bnd, L.tile_bnd)) return 1;
    C->n_int = (int)L.tile_int.size(); C->n_bnd = (int)L.tile_bnd.size();
    if (!C->sx) {
      CUDA_OK(cudaStreamCreateWithFlags(&C->sx, cudaStreamNonBlocking));
      for (cudaEvent_t *e : {&C->e_a, &C->e_g, &C->e_b, &C->e_p}) CUDA_OK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    }
    const int nsend = L.send_ptr.empty() ? 0 : L.send_ptr.back();
    const int *si;
    if (dev_upload(si, L.send_idx)) return 1;
    C->send_idx = const_cast<int *>(si);
    if (dev_alloc(C->sendbuf, (size_t)std::max(1, nsend) * 9)) return 1;
  }
  if (setup_p2p()) return 1;
  if (C->gi.partial_sums && C->comm) {  // partition-local build: add the ranks' shares of the global counts and sums up
    GlobalInfo &gi = C->gi;
    double h[7] = {(double)gi.nedges, (double)gi.nedges_bndr, (double)gi.ncells_intr, (double)gi.ncells_bndr, gi.vol_sum, gi.heff2, gi.vol_green};
    double *d = nullptr;
    CUDA_OK(cudaMalloc(&d, sizeof h));
    CUDA_OK(cudaMemcpyAsync(d, h, sizeof h, cudaMemcpyHostToDevice, C->st));
    NCCL_OK(g_nccl.AllReduce(d, d, 7, ncclDouble, ncclSum, C->comm, C->st));
    CUDA_OK(cudaMemcpyAsync(h, d, sizeof h, cudaMemcpyDeviceToHost, C->st));
    CUDA_OK(cudaStreamSynchronize(C->st));
    cudaFree(d);
    gi.nedges = (int)h[0]; gi.nedges_bndr = (int)h[1]; gi.nedges_intr = gi.nedges - gi.nedges_bndr;
    gi.ncells_intr = (int)h[2]; gi.ncells_bndr = (int)h[3];
    gi.vol_sum = h[4]; gi.heff1 = std::sqrt(h[4] / (double)gi.ncells); gi.heff2 = h[5] / (double)gi.ncells; gi.vol_green = h[6];
    gi.partial_sums = false;
  }
  if (C->ev_pool.empty()) {  // once per context (a second set_mesh reuses them)
    C->ev_pool.assign(8192, nullptr);
    for (auto &e : C->ev_pool) CUDA_OK(cudaEventCreate(&e));
  }
  C->has_mesh = true;
  return 0;
}
}  // namespace

extern "C" {

int fvs2d_gpu_set_state(const double *cvar) {
  NEED(C && C->has_mesh, "fvs2d_gpu_set_state: no mesh");
  NEED(cvar != nullptr, "fvs2d_gpu_set_state: null cvar");
  if (C->nranks > 1) {  // only the entries of the cells this rank stores (owned + ghosts) are read and uploaded
    const int n_loc = C->L.n_loc;
    if (ensure_host_stage((size_t)n_loc * 4)) return 1;
    double *loc = C->h_stage;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n_loc; i++) {
      const size_t o = C->L.orig_id[i];
      for (int v = 0; v < 4; v++) loc[(size_t)i * 4 + v] = cvar[o * 4 + v];
    }
    return upload_local_state(loc);
  }
  const size_t ng = (size_t)C->L.nc_global * 4;
  if (ensure_stage(ng)) return 1;
  CUDA_OK(cudaMemcpyAsync(C->stage_aos, cvar, ng * 8, cudaMemcpyHostToDevice, C->st));
  k_scatter_in<<<cdiv(C->L.n_loc, 256), 256, 0, C->st>>>(C->L.n_loc, C->np, 4, C->dm.orig_id, C->stage_aos, C->q, 0);
  k_prim<<<cdiv(C->L.n_loc, 256), 256, 0, C->st>>>(C->L.n_loc, C->np, C->cfg.gamma, C->q, C->pa);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaStreamSynchronize(C->st));
  C->has_state = true;
  return 0;
}

int fvs2d_gpu_set_state_local(const double *cvar_own) {
  NEED(C && C->has_mesh, "fvs2d_gpu_set_state_local: no mesh");
  NEED(cvar_own != nullptr, "fvs2d_gpu_set_state_local: null cvar");
  const size_t n = (size_t)C->L.n_own * 4;
  if (ensure_stage(n)) return 1;
  CUDA_OK(cudaMemcpyAsync(C->stage_aos, cvar_own, n * 8, cudaMemcpyHostToDevice, C->st));
  k_scatter_in<<<cdiv(C->L.n_own, 256), 256, 0, C->st>>>(C->L.n_own, C->np, 4, nullptr, C->stage_aos, C->q, 0);
  k_prim<<<cdiv(C->L.n_own, 256), 256, 0, C->st>>>(C->L.n_own, C->np, C->cfg.gamma, C->q, C->pa);
  CUDA_OK(cudaGetLastError());
  if (C->nranks > 1) {  // ghost copies of the primitive state come from their owners
    HaloItem it{C->pa, 4, 1};
    if (halo_exchange(&it, 1)) return 1;
  }
  CUDA_OK(cudaStreamSynchronize(C->st));
  C->has_state = true;
  return 0;
}

int fvs2d_gpu_get_state_local(double *cvar_own) {
  NEED(C && C->has_state, "fvs2d_gpu_get_state_local: no state");
  NEED(cvar_own != nullptr, "fvs2d_gpu_get_state_local: null cvar");
  const size_t n = (size_t)C->L.n_own * 4;
  if (ensure_stage(n)) return 1;
  k_gather_out<<<cdiv(C->L.n_own, 256), 256, 0, C->st>>>(C->L.n_own, C->np, 4, nullptr, C->q, C->stage_aos, 0, 0);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaMemcpyAsync(cvar_own, C->stage_aos, n * 8, cudaMemcpyDeviceToHost, C->st));
  CUDA_OK(cudaStreamSynchronize(C->st));
  return 0;
}

int fvs2d_gpu_get_state(double *cvar) {
  NEED(C && C->has_state, "fvs2d_gpu_get_state: no state");
  NEED(cvar != nullptr, "fvs2d_gpu_get_state: null cvar");
  return download_aos(C->q, 4, cvar);
}

int fvs2d_gpu_initialize_solution(void) {
  NEED(C && C->has_mesh, "fvs2d_gpu_initialize_solution: no mesh");
  const fvs2d_config &c = C->cfg;
  NEED(c.ntstart <= 1, "fvs2d_gpu_initialize_solution: ntstart>1 is a restart; pass cont.s8's cvar to fvs2d_gpu_set_state");
  // evaluated for the cells this rank stores (owned + ghosts), in the library's order: no global array is formed
  const Layout &L = C->L;
  const int nc = L.n_loc;
  std::vector<double> cv(4 * (size_t)nc);
  const double pi = std::acos(-1.0), g = c.gamma;
  const double t0 = (double)(c.ntstart - 1) * c.dt;
#pragma omp parallel for schedule(static)
  for (int ic = 0; ic < nc; ic++) {
    double pv[4];
    const double x = L.xc[ic], y = L.yc[ic];
    if (c.ntstart == 1 && c.lvortex) {  // src/mms.f90:219-265
      const double ri = c.vortex_inf[0], ui = c.vortex_inf[1], vi = c.vortex_inf[2], p_i = c.vortex_inf[3];
      const double dx = x - (c.vortex_pos[0] + ui * t0), dy = y - (c.vortex_pos[1] + vi * t0);
      const double r = std::sqrt(dx * dx + dy * dy), kk = c.vortex_kappa / (2.0 * pi);
      pv[1] = ui - kk * dy * std::exp(0.5 * (1.0 - r * r));
      pv[2] = vi + kk * dx * std::exp(0.5 * (1.0 - r * r));
      const double temp = p_i / ri - kk * kk * (g - 1.0) / (2.0 * g) * std::exp(1.0 - r * r);
      pv[0] = std::pow(temp, 1.0 / (g - 1.0));
      pv[3] = std::pow(pv[0], g);
    } else if (c.ntstart == 1) {
      for (int v = 0; v < 4; v++) pv[v] = c.pvar_inf[v];
    } else {  // manufactured solution, src/mms.f90:137-146
      for (int v = 0; v < 4; v++) pv[v] = c.mms_c[v][0] + c.mms_c[v][1] * std::sin(c.mms_c[v][2] * x + c.mms_c[v][3] * y);
    }
    double *q = &cv[4 * (size_t)ic];  // pvar2cvar, src/data_solution.f90:92-106
    q[0] = pv[0]; q[1] = pv[0] * pv[1]; q[2] = pv[0] * pv[2];
    q[3] = pv[3] / (g - 1.0) + 0.5 * pv[0] * (pv[1] * pv[1] + pv[2] * pv[2]);
  }
  return upload_local_state(cv.data());
}

int fvs2d_gpu_compute_residual(double time, double *resid, double *ws_nrml) {
  NEED(C && C->has_state, "fvs2d_gpu_compute_residual: no state");
  const size_t np = C->np;
  if (!C->resid && (dev_alloc(C->resid, 4 * np) || dev_alloc(C->ws, np))) return 1;
  C->last_launches = 0;
  {
    StepClock hc{};
    hc.t1 = time;
    CUDA_OK(cudaMemcpyAsync(C->clk, &hc, sizeof hc, cudaMemcpyHostToDevice, C->st));
    CUDA_OK(cudaStreamSynchronize(C->st));  // hc is a stack variable
  }
  if (launch_bc(0)) return 1;
  if (pass_a(C->pa)) return 1;
  StageParams S{0, 0, 0.0, 0.0, C->cfg.dt};
  if (launch_flux(UM_RESID, S, C->pa, C->pb)) return 1;
  CUDA_OK(cudaGetLastError());
  if (resid && download_aos(C->resid, 4, resid)) return 1;
  if (ws_nrml && download_aos(C->ws, 1, ws_nrml)) return 1;
  CUDA_OK(cudaStreamSynchronize(C->st));
  return 0;
}

int fvs2d_gpu_get_aux(double *pvar, double *grad, double *phi_lim) {
  NEED(C && C->has_state, "fvs2d_gpu_get_aux: no state");
  // pvar, grad and phi_lim of ONE state, the current one (as if compute_residual had just been called on it): the fused
  // stage kernel never writes gradients to global memory, and the two-pass path leaves those of the last stage's input
  if ((grad || phi_lim) && launch_gradient(C->pa)) return 1;
  if (pvar && download_aos(C->pa, 4, pvar, 1, 0)) return 1;
  if (phi_lim && download_aos(C->phi, 1, phi_lim)) return 1;
  if (grad) {  // Fortran grad(ivar,ic,idim): two planes of (4,ncells); device g holds gx0-3, gy0-3 pair-interleaved
    const size_t plane = 4 * (size_t)C->L.nc_global;
    if (download_aos(C->g, 4, grad, 1, 0)) return 1;
    if (download_aos(C->g, 4, grad + plane, 1, 4)) return 1;
  }
  return 0;
}

// m-cell (id in C->mesh: the whole mesh or this rank's submesh) -> local id of the cells this rank stores, -1 otherwise
static std::vector<int> mcell_to_local() {
  const HostMesh &m = C->mesh;
  const std::vector<int> &orig = C->sub.orig;
  std::vector<int> m2l(m.ncells, -1);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < C->L.n_loc; i++) {
    const int o = C->L.orig_id[i];
    m2l[orig.empty() ? o : (int)(std::lower_bound(orig.begin(), orig.end(), o) - orig.begin())] = i;
  }
  return m2l;
}

int fvs2d_gpu_interpolate_cell2node(const int select[4], double *fnode) {
  NEED(C && C->has_state, "fvs2d_gpu_interpolate_cell2node: no state");
  NEED(select && fnode, "fvs2d_gpu_interpolate_cell2node: null argument");
  const HostMesh &m = C->mesh;
  int mask = 0, nsel = 0;
  for (int v = 0; v < 4; v++) if (select[v]) { mask |= 1 << v; nsel++; }
  if (nsel == 0) return 0;
  if (!C->d_n2c_ptr) {
    // first use: the nodes of this rank's OWNED cells, their node -> cell lists restricted to owned cells (local ids,
    // ascending original id) and the weights of cell2node_idw_setup (src/interpolation.f90:62-101).  The weight of a cell
    // is normalised by ALL cells around the node (their geometry is in the rank's submesh even where their state is not),
    // so on several ranks a node's value is the SUM of the ranks' shares.
    const std::vector<int> m2l = mcell_to_local();
    const int n_own = C->L.n_own;
    std::vector<unsigned char> used(m.nnodes, 0);
    for (int mc = 0; mc < m.ncells; mc++)
      if (m2l[mc] >= 0 && m2l[mc] < n_own)
        for (int sl = m.cptr[mc]; sl < m.cptr[mc + 1]; sl++) used[m.cnode[sl]] = 1;
    std::vector<int> &nodes = C->out_nodes;
    nodes.clear();
    for (int in = 0; in < m.nnodes; in++) if (used[in]) nodes.push_back(in);
    const int nl = (int)nodes.size();
    std::vector<int> ptr(nl + 1, 0);
    for (int k = 0; k < nl; k++) {
      int cnt = 0;
      for (int j = m.n2c_ptr[nodes[k]]; j < m.n2c_ptr[nodes[k] + 1]; j++) cnt += m2l[m.n2c[j]] >= 0 && m2l[m.n2c[j]] < n_own;
      ptr[k + 1] = ptr[k] + cnt;
    }
    std::vector<int> n2c(ptr[nl]);
    std::vector<double> idw(ptr[nl]);
#pragma omp parallel for schedule(static)
    for (int k = 0; k < nl; k++) {
      const int in = nodes[k];
      double idt = 0.0;
      for (int j = m.n2c_ptr[in]; j < m.n2c_ptr[in + 1]; j++) {
        const int ic = m.n2c[j];
        const double dx = m.xc[ic] - m.xn[in], dy = m.yc[ic] - m.yn[in];
        idt = idt + 1.0 / std::sqrt(dx * dx + dy * dy);
      }
      int e = ptr[k];
      for (int j = m.n2c_ptr[in]; j < m.n2c_ptr[in + 1]; j++) {
        const int ic = m.n2c[j], l = m2l[ic];
        if (l < 0 || l >= n_own) continue;
        const double dx = m.xc[ic] - m.xn[in], dy = m.yc[ic] - m.yn[in];
        n2c[e] = l;
        idw[e] = 1.0 / std::sqrt(dx * dx + dy * dy) / idt;
        e++;
      }
    }
    if (!C->sub.node_orig.empty()) for (int &v : nodes) v = C->sub.node_orig[v];  // original node ids of the caller's array
    if (dev_upload(C->d_n2c, n2c) || dev_upload(C->d_idw, idw) || dev_alloc(C->d_fnode, 4 * (size_t)std::max(1, nl))) return 1;
    if (dev_upload(C->d_n2c_ptr, ptr)) return 1;
  }
  const int nl = (int)C->out_nodes.size();
  const size_t nn = (size_t)C->gi.nnodes;
  // C->pa is the primitive state of the current cvar (cvar2pvar of write_inst_ios, src/io.f90:135)
  if (nl) k_cell2node<<<cdiv(nl, 256), 256, 0, C->st>>>(nl, C->np, mask, C->d_n2c_ptr, C->d_n2c, C->d_idw, C->pa, C->d_fnode);
  CUDA_OK(cudaGetLastError());
  if (C->nranks == 1 && (size_t)nl == nn) {  // every node belongs to some cell: the records are the caller's, in place
    CUDA_OK(cudaMemcpyAsync(fnode, C->d_fnode, (size_t)nsel * nn * 8, cudaMemcpyDeviceToHost, C->st));
    CUDA_OK(cudaStreamSynchronize(C->st));
    return 0;
  }
  std::vector<double> tmp((size_t)nsel * std::max(1, nl));
  CUDA_OK(cudaMemcpyAsync(tmp.data(), C->d_fnode, tmp.size() * 8, cudaMemcpyDeviceToHost, C->st));
  CUDA_OK(cudaStreamSynchronize(C->st));
  std::fill(fnode, fnode + (size_t)nsel * nn, 0.0);
  for (int k = 0; k < nsel; k++)
    for (int i = 0; i < nl; i++) fnode[(size_t)k * nn + C->out_nodes[i]] = tmp[(size_t)k * nl + i];
  return 0;
}

int fvs2d_gpu_wall_values(int ib, double *vals) {
  NEED(C && C->has_state, "fvs2d_gpu_wall_values: no state");
  NEED(vals != nullptr, "fvs2d_gpu_wall_values: null argument");
  const HostMesh &m = C->mesh;
  NEED(ib >= 0 && ib < m.nb, "fvs2d_gpu_wall_values: boundary index out of range");
  if (!C->d_be_cell) {
    // first use: the boundary edges whose cell this rank OWNS, per boundary in bndry(ib)%edge order: cell (edge%c1, local
    // id), geometry, and the position of the edge in the caller's list
    const std::vector<int> m2l = mcell_to_local();
    const bool part = !C->sub.orig.empty();
    std::vector<int> cell, &pos = C->be_pos, &ptr = C->be_ptr;
    std::vector<double2> exy, enxy;
    pos.clear();
    ptr.assign(m.nb + 1, 0);
    for (int b = 0; b < m.nb; b++) {
      for (int k = m.b_edge_ptr[b]; k < m.b_edge_ptr[b + 1]; k++) {
        const int je = m.b_edge[k], l = m2l[m.ec1[je]];
        if (l < 0 || l >= C->L.n_own) continue;
        if (part) {  // the caller's list position of the CELL: valid when every listed cell has one boundary edge, as
                     // the reference's own boundary loop assumes (src/residual.f90:112-125 indexes cell(i) and edge(i) alike)
          const int src = m.b_edge_src[k];
          NEED((k == m.b_edge_ptr[b] || m.b_edge_src[k - 1] != src) && (k + 1 == m.b_edge_ptr[b + 1] || m.b_edge_src[k + 1] != src),
               "fvs2d_gpu_wall_values on several ranks needs one boundary edge per listed boundary cell");
          pos.push_back(C->sub.b_pos[src]);
        } else {
          pos.push_back(k - m.b_edge_ptr[b]);
        }
        const EdgeGeom eg = edge_geom(m, je);
        cell.push_back(l);
        exy.push_back(make_double2(eg.x, eg.y));
        enxy.push_back(make_double2(eg.nx, eg.ny));
      }
      ptr[b + 1] = (int)cell.size();
    }
    if (dev_upload(C->d_be_cell, cell) || dev_upload(C->d_be_xy, exy) || dev_upload(C->d_be_nxy, enxy) ||
        dev_alloc(C->d_be_out, 4 * std::max<size_t>(1, cell.size()))) return 1;
  }
  const int e0 = C->be_ptr[ib], n = C->be_ptr[ib + 1] - e0;
  if (n == 0) return 0;
  // gradient_cellcntr_1var (src/gradient.f90:74-96): the unlimited gradient of the selected scheme, also for
  // first-order reconstruction, of the current primitive state
  if (launch_gradient(C->pa, nullptr, 0, true)) return 1;
  k_wall_values<<<cdiv(n, 128), 128, 0, C->st>>>(n, C->np, C->d_be_cell + e0, C->d_be_xy + e0, C->d_be_nxy + e0, C->dm.xy, C->pa, C->g,
                                                 C->d_be_out);
  CUDA_OK(cudaGetLastError());
  if (C->nranks == 1) {  // every edge of the boundary, in order
    CUDA_OK(cudaMemcpyAsync(vals, C->d_be_out, (size_t)n * 32, cudaMemcpyDeviceToHost, C->st));
    CUDA_OK(cudaStreamSynchronize(C->st));
    return 0;
  }
  std::vector<double> tmp(4 * (size_t)n);
  CUDA_OK(cudaMemcpyAsync(tmp.data(), C->d_be_out, tmp.size() * 8, cudaMemcpyDeviceToHost, C->st));
  CUDA_OK(cudaStreamSynchronize(C->st));
  for (int i = 0; i < n; i++)
    for (int v = 0; v < 4; v++) vals[4 * (size_t)C->be_pos[e0 + i] + v] = tmp[4 * (size_t)i + v];
  return 0;
}

int fvs2d_gpu_time_integration(double t1, int nsub, double *res_l2, double *vortex_err, double *vortex_err_xy) {
  NEED(C && C->has_state, "fvs2d_gpu_time_integration: no state");
  NEED(nsub >= 0, "fvs2d_gpu_time_integration: nsub < 0");
  const fvs2d_config &c = C->cfg;
  const int um = c.ssprk ? UM_SSPRK : UM_RK;
  const double dt = c.dt;
  const bool vort = c.lvortex != 0;
  // device log: per step 4 sums of squares, 13 vortex numbers + the centroid of the largest density error, 1 id
  const size_t per = 4 + 13 + 2;
  if (C->log_cap < (size_t)nsub) {
    // at least 4096 rows (0.5 MB): the step graph is tied to this buffer, so growing it from call to call would
    // force a re-capture inside the caller's time loop
    const size_t cap = std::max<size_t>((size_t)nsub, 4096);
    dev_free(C->logbuf, per * C->log_cap * sizeof(double));
    dev_free(C->logid, C->log_cap * sizeof(int));
    if (dev_alloc(C->logbuf, per * cap) || dev_alloc(C->logid, cap)) return 1;
    C->log_cap = cap;
  }
  C->last_launches = 0;
  C->ev_used = 0; C->ev_dropped = 0;
  for (auto &s : C->ev_spans) s.clear();
  double span_ms[3] = {0.0, 0.0, 0.0};
  // option "timing": the event pool holds ~300 steps' worth of spans; it is drained (one stream synchronisation) whenever it
  // runs low, so the per-phase times of a long call are complete
  auto drain_spans = [&]() -> int {
    CUDA_OK(cudaStreamSynchronize(C->st));
    if (C->sx) CUDA_OK(cudaStreamSynchronize(C->sx));
    for (int b = 0; b < 3; b++) {
      for (auto &pr : C->ev_spans[b]) { float t = 0; cudaEventElapsedTime(&t, C->ev_pool[pr.first], C->ev_pool[pr.second]); span_ms[b] += t; }
      C->ev_spans[b].clear();
    }
    C->ev_used = 0;
    return 0;
  };
  const bool overlap = C->nranks > 1 && C->opt_overlap && C->tile_ok && C->opt_tile == 2;
  bool p_pending = false;
  if (C->opt_fuse && ensure_fused()) return 1;
  const bool fused = C->opt_fuse != 0 && C->fz_state == 1 && C->opt_tile == 2 && !(use_pair_kernels() && C->opt_fuse < 0);
  {  // device step clock (src/runge_kutta.f90:135-145, 218-221): stage times are told + off[stage]
    StepClock hc{};
    hc.t1 = t1; hc.dt = dt; hc.istep = 0; hc.epoch0 = C->epoch_total;
    for (int rk = 0; rk < 4; rk++) hc.off[rk] = c.ssprk ? C->dts[rk] : (rk == 0 ? 0.0 : rk == 3 ? dt : 0.5 * dt);
    hc.off_end = c.ssprk ? C->dte[3] : dt;
    CUDA_OK(cudaMemcpyAsync(C->clk, &hc, sizeof hc, cudaMemcpyHostToDevice, C->st));
    CUDA_OK(cudaStreamSynchronize(C->st));  // hc is a stack variable
  }
  CUDA_OK(cudaEventRecord(C->ev0, C->st));
  // one time step = a fixed kernel sequence; nothing in it depends on the host-side step counter
  auto run_step = [&]() -> int {
    for (int rk = 0; rk < 4; rk++) {
      StageParams S;
      S.stage = rk; S.last = rk == 3; S.c = C->rk_coef[rk]; S.dt = dt;
      if (c.steady) S.h = c.ssprk ? (rk == 3 ? 1.0 / 4.0 : 1.0 / 3.0) : (rk == 2 ? 1.0 : rk == 3 ? 1.0 / 6.0 : 1.0 / 2.0);
      else S.h = C->h_rk[rk];
      if (launch_bc(rk)) return 1;
      if (fused) {
        // one kernel per stage; on several ranks it also delivers the new state into the peers' ghost slots (HaloP2P)
        if (launch_fused(um, S, C->pa, C->pb)) return 1;
      } else if (overlap) {
        // interior tiles never read a ghost: they run while the exchanges are in flight on the second stream
        const bool grad = C->recon != RC_FIRST;
        if (grad) {
          if (launch_gradient(C->pa, C->d_tile_int, C->n_int)) return 1;
          if (p_pending) CUDA_OK(cudaStreamWaitEvent(C->st, C->e_p, 0));   // ghost state of this stage has landed
          if (launch_gradient(C->pa, C->d_tile_bnd, C->n_bnd)) return 1;
          CUDA_OK(cudaEventRecord(C->e_a, C->st));
          CUDA_OK(cudaStreamWaitEvent(C->sx, C->e_a, 0));
          HaloItem it[2] = {{C->g, 8, 1}, {C->phi, 1, 0}};
          if (halo_exchange(it, c.limiter > 0 ? 2 : 1, C->sx)) return 1;
          CUDA_OK(cudaEventRecord(C->e_g, C->sx));
        }
        if (launch_flux(um, S, C->pa, C->pb, C->d_tile_int, C->n_int, 0)) return 1;
        const int parts_int = C->nparts;
        if (grad) CUDA_OK(cudaStreamWaitEvent(C->st, C->e_g, 0));          // ghost gradients have landed
        else if (p_pending) CUDA_OK(cudaStreamWaitEvent(C->st, C->e_p, 0));
        if (launch_flux(um, S, C->pa, C->pb, C->d_tile_bnd, C->n_bnd, parts_int)) return 1;
        CUDA_OK(cudaEventRecord(C->e_b, C->st));
        CUDA_OK(cudaStreamWaitEvent(C->sx, C->e_b, 0));
        HaloItem itp{C->pb, 4, 1};
        if (halo_exchange(&itp, 1, C->sx)) return 1;                       // overlaps the next stage's interior gradient
        CUDA_OK(cudaEventRecord(C->e_p, C->sx));
        p_pending = true;
      } else {
        if (pass_a(C->pa)) return 1;
        if (launch_flux(um, S, C->pa, C->pb)) return 1;
        if (C->nranks > 1) {
          Span sp(0);
          HaloItem it{C->pb, 4, 1};
          if (halo_exchange(&it, 1)) return 1;
        }
      }
      std::swap(C->pa, C->pb);
    }
    {
      Span sp(0);
      int vgrid = 0;
      if (vort) {
        vgrid = std::min(C->nblocks, C->nsm * kVortexCtas);
        launch_chain(k_vortex_err, dim3(vgrid), dim3(kBlock), 0, C->dm, C->phys, C->clk, C->q, C->vpartial, C->vbest, C->vbest_loc);
        C->last_launches++;
      }
      launch_chain(k_finish_step, dim3(1), dim3(kFinishThreads), 0, C->partial, C->nparts, C->vpartial, C->vbest, C->vbest_loc, C->dm.xy, vgrid, C->logbuf, (int)per,
                                                      C->logid, C->clk);
      C->last_launches++;
    }
    return 0;
  };
  // Small meshes are launch-bound (a 65 k-cell step is ~10 kernels of ~10 us): the first step runs eagerly (it also
  // configures the kernels), the remaining ones replay a CUDA graph captured from the same sequence.
  // (several ranks: only the fused path, whose halo exchange lives inside the stage kernel -- no NCCL call in the step)
  const bool use_graph = C->opt_graph && (C->nranks == 1 || fused) && !C->opt_timing && nsub >= 3;
  struct PdlScope { ~PdlScope() { C->pdl_on = false; } } pdl_scope;   // off again on every way out of the call
  C->pdl_on = C->opt_pdl && C->nranks == 1 && !C->opt_timing;
  int done = 0;
  if (nsub > 0) { if (run_step()) return 1; done = 1; }
  if (use_graph) {
    if (C->graph_exec && (C->graph_logbuf != C->logbuf || C->graph_um != (fused ? 256 * ((C->opt_fuse & 3) + 1) : 0) + (use_pair_kernels() ? 2048 : 0) + (C->pdl_on ? 4096 : 0) + um * 16 + C->opt_tile * 4 + C->opt_ctas)) {
      cudaGraphExecDestroy(C->graph_exec);
      C->graph_exec = nullptr;
    }
    if (!C->graph_exec) {
      const long l0 = C->last_launches;
      cudaGraph_t graph = nullptr;
      CUDA_OK(cudaStreamBeginCapture(C->st, cudaStreamCaptureModeThreadLocal));
      const int rc = run_step();
      const cudaError_t ce = cudaStreamEndCapture(C->st, &graph);
      C->last_launches = l0;
      if (rc || ce != cudaSuccess) { if (graph) cudaGraphDestroy(graph); return rc ? 1 : fail("CUDA graph capture failed: %s", cudaGetErrorString(ce)); }
      CUDA_OK(cudaGraphInstantiate(&C->graph_exec, graph, 0));
      cudaGraphDestroy(graph);
      C->graph_logbuf = C->logbuf;
      C->graph_um = (fused ? 256 * ((C->opt_fuse & 3) + 1) : 0) + (use_pair_kernels() ? 2048 : 0) + (C->pdl_on ? 4096 : 0) + um * 16 + C->opt_tile * 4 + C->opt_ctas;
    }
    const long per_step = C->last_launches;
    for (; done < nsub; done++) CUDA_OK(cudaGraphLaunch(C->graph_exec, C->st));
    C->last_launches = per_step * nsub;
  } else {
    for (; done < nsub; done++) {
      if (C->opt_timing && C->ev_used + 128 > C->ev_pool.size() && drain_spans()) return 1;
      if (run_step()) return 1;
    }
  }
  int p2p_timed_out = 0;
  if (p_pending) CUDA_OK(cudaStreamWaitEvent(C->st, C->e_p, 0));  // the last exchange belongs to this call
  if (fused && C->nranks > 1) {
    // the peers' last stage lands in this rank's ghost slots: complete before the call returns (compute_residual,
    // the output path and the next call's first gather read them)
    C->epoch_total += 4u * (unsigned)nsub;
    PeerList pl{};
    pl.n = (int)C->L.peers.size();
    for (int k = 0; k < pl.n; k++) pl.r[k] = C->L.peers[k];
    if (nsub > 0) k_wait_peers<<<1, 32, 0, C->st>>>(C->flags, pl, C->epoch_total, C->p2p_timed_out);
    CUDA_OK(cudaMemcpyAsync(&p2p_timed_out, C->p2p_timed_out, sizeof(int), cudaMemcpyDeviceToHost, C->st));
  }
  CUDA_OK(cudaEventRecord(C->ev1, C->st));
  CUDA_OK(cudaGetLastError());
  // ---- logs back to the host (the only device->host traffic of the call)
  std::vector<double> lg(per * (size_t)nsub);
  std::vector<int> ids(nsub);
  if (nsub > 0 && (res_l2 || vortex_err || vortex_err_xy)) {
    CUDA_OK(cudaMemcpyAsync(lg.data(), C->logbuf, lg.size() * 8, cudaMemcpyDeviceToHost, C->st));
    CUDA_OK(cudaMemcpyAsync(ids.data(), C->logid, ids.size() * 4, cudaMemcpyDeviceToHost, C->st));
  }
  CUDA_OK(cudaStreamSynchronize(C->st));
  NEED(!p2p_timed_out, "in-kernel halo exchange timed out after 20 s: a peer never delivered a stage -- the ranks do not make the "
                       "same sequence of fvs2d_gpu_time_integration calls, or a peer has failed");
  float ms = 0;
  CUDA_OK(cudaEventElapsedTime(&ms, C->ev0, C->ev1));
  C->last_ms[0] = ms;
  if (drain_spans()) return 1;
  for (int b = 0; b < 3; b++) C->last_ms[b == 0 ? 3 : b] = span_ms[b];
  if (nsub == 0 || !(res_l2 || vortex_err || vortex_err_xy)) return 0;

  // ---- combine across ranks (sums / maxima are tiny: 17 doubles per step)
  const int nr = C->nranks;
  std::vector<double> all;  // [rank][step][per]
  std::vector<int> allid;
  if (nr > 1) {
    double *dall = nullptr; int *dallid = nullptr;
    struct Scratch { double *&a; int *&b; ~Scratch() { cudaFree(a); cudaFree(b); } } scratch{dall, dallid};  // freed on every path
    CUDA_OK(cudaMalloc(&dall, lg.size() * 8 * nr));
    CUDA_OK(cudaMalloc(&dallid, ids.size() * 4 * nr));
    NCCL_OK(g_nccl.AllGather(C->logbuf, dall, lg.size(), ncclDouble, C->comm, C->st));
    NCCL_OK(g_nccl.AllGather(C->logid, dallid, ids.size(), ncclInt, C->comm, C->st));
    all.resize(lg.size() * nr); allid.resize(ids.size() * nr);
    CUDA_OK(cudaMemcpyAsync(all.data(), dall, all.size() * 8, cudaMemcpyDeviceToHost, C->st));
    CUDA_OK(cudaMemcpyAsync(allid.data(), dallid, allid.size() * 4, cudaMemcpyDeviceToHost, C->st));
    CUDA_OK(cudaStreamSynchronize(C->st));
  } else { all = lg; allid = ids; }
  const double ncg = (double)C->L.nc_global, nin = (double)C->gi.ncells_intr;
  for (int s = 0; s < nsub; s++) {
    double sum4[4] = {0, 0, 0, 0}, mx[4] = {0, 0, 0, 0}, s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0}, best = -1.0, bx = 0, by = 0;
    int bid = 0x7fffffff;
    for (int r = 0; r < nr; r++) {
      const double *a = &all[((size_t)r * nsub + s) * per];
      for (int v = 0; v < 4; v++) { sum4[v] += a[v]; mx[v] = std::max(mx[v], a[4 + v]); s1[v] += a[8 + v]; s2[v] += a[12 + v]; }
      const int id = allid[(size_t)r * nsub + s];
      if (a[16] > best || (a[16] == best && id < bid)) { best = a[16]; bid = id; bx = a[17]; by = a[18]; }
    }
    if (res_l2) for (int v = 0; v < 4; v++) res_l2[4 * s + v] = std::sqrt(sum4[v] / ncg);  // src/runge_kutta.f90:172-181
    if (vort && vortex_err) {  // the 14 columns of src/mms.f90:357-361
      double *o = &vortex_err[14 * (size_t)s];
      const double told = t1 + (double)s * dt;
      o[0] = c.ssprk ? told + C->dte[3] : told + dt;
      for (int v = 0; v < 4; v++) { o[1 + 3 * v] = mx[v]; o[2 + 3 * v] = s1[v] / nin; o[3 + 3 * v] = std::sqrt(s2[v] / nin); }
      o[13] = std::sqrt((s2[0] + s2[1] + s2[2] + s2[3]) / nin);
    }
    if (vort && vortex_err_xy) {
      // maxloc(erho) (src/mms.f90:363): first cell of the maximum; cell 1 when every error is zero
      const bool found = best > 0.0 && bid != 0x7fffffff;
      vortex_err_xy[2 * s] = found ? bx : C->gi.xy_cell0[0];
      vortex_err_xy[2 * s + 1] = found ? by : C->gi.xy_cell0[1];
    }
  }
  return 0;
}

int fvs2d_gpu_test_resid(int corrected, double l2[4], double linf[4]) {
  NEED(C && C->has_state, "fvs2d_gpu_test_resid: no state");
  NEED(!C->cfg.lvortex, "test_resid needs mms_source, which exists only when lvortex is false (src/mms.f90:45-66)");
  NEED(C->nranks == 1, "fvs2d_gpu_test_resid: single GPU only");
  const HostMesh &m = C->mesh;
  std::vector<double> r(4 * (size_t)m.ncells);
  if (fvs2d_gpu_compute_residual(0.0, r.data(), nullptr)) return 1;
  const fvs2d_config &c = C->cfg;
  const double g = c.gamma;
  for (int v = 0; v < 4; v++) { l2[v] = 0; linf[v] = 0; }
  for (int i = 0; i < m.ncells_intr; i++) {
    const int ic = m.cell_intr[i];
    const double x = m.xc[ic], y = m.yc[ic];
    // src/mms.f90:124-181
    double f0[4], fx[4], fy[4];
    for (int v = 0; v < 4; v++) {
      const double a0 = c.mms_c[v][0], as = c.mms_c[v][1], ax = c.mms_c[v][2], ay = c.mms_c[v][3];
      f0[v] = a0 + as * std::sin(ax * x + ay * y);
      fx[v] = ax * as * std::cos(ax * x + ay * y);
      fy[v] = ay * as * std::cos(ax * x + ay * y);
    }
    const double rr = f0[0], u = f0[1], vv = f0[2], p = f0[3];
    const double rx = fx[0], ux = fx[1], vx = fx[2], px = fx[3], ry = fy[0], uy = fy[1], vy = fy[2], py = fy[3];
    const double rH = g / (g - 1.0) * p + rr * u * u / 2.0 + rr * vv * vv / 2.0;
    const double rHx = g / (g - 1.0) * px + rx * (u * u + vv * vv) / 2.0 + rr * (u * ux + vv * vx);
    const double rHy = g / (g - 1.0) * py + ry * (u * u + vv * vv) / 2.0 + rr * (u * uy + vv * vy);
    double rhs[4];
    rhs[0] = corrected ? rx * u + rr * ux + ry * vv + rr * vy : rx * u + u * rx + ry * vv + rr * vy;  // :169 typo kept
    rhs[1] = rx * u * u + 2.0 * rr * u * ux + ry * u * vv + rr * uy * vv + rr * u * vy + px;
    rhs[2] = rx * u * vv + rr * ux * vv + rr * u * vx + ry * vv * vv + 2.0 * rr * vv * vy + py;
    rhs[3] = u * rHx + ux * rH + vv * rHy + vy * rH;
    for (int v = 0; v < 4; v++) {
      const double e = r[4 * (size_t)ic + v] + rhs[v];
      l2[v] += e * e;
      linf[v] = std::max(linf[v], std::fabs(e));
    }
  }
  for (int v = 0; v < 4; v++) l2[v] = std::sqrt(l2[v] / (double)m.ncells_intr);
  return 0;
}

int fvs2d_gpu_sizes(int s[10]) {
  NEED(C && C->has_mesh, "fvs2d_gpu_sizes: no mesh");
  const GlobalInfo &m = C->gi;
  s[0] = m.nnodes; s[1] = m.ncells; s[2] = m.nedges; s[3] = m.nedges_intr; s[4] = m.nedges_bndr;
  s[5] = m.ncells_intr; s[6] = m.ncells_bndr; s[7] = C->L.n_own; s[8] = C->L.n_loc; s[9] = C->L.nedges;
  return 0;
}

int fvs2d_gpu_scalars(double s[6]) {
  NEED(C && C->has_mesh, "fvs2d_gpu_scalars: no mesh");
  const GlobalInfo &m = C->gi;
  s[0] = m.heff1; s[1] = m.heff2; s[2] = m.vol_sum; s[3] = m.vol_green; s[4] = C->L.lsq_verify_err; s[5] = (double)C->bytes;
  return 0;
}

long fvs2d_gpu_mesh_array(const char *name, void *out) {
  if (!C || !C->has_mesh) { fail("fvs2d_gpu_mesh_array: no mesh"); return -1; }
  const HostMesh &m = C->mesh;
  const std::string n = name;
#define RET(nm, vec) if (n == nm) { if (out) memcpy(out, (vec).data(), (vec).size() * sizeof((vec)[0])); return (long)(vec).size(); }
  RET("xc", m.xc) RET("yc", m.yc) RET("vol", m.vol)
  if (n == "ex" || n == "ey" || n == "ea" || n == "enx" || n == "eny") {  // materialised on demand (verification only)
    if (out) for (int je = 0; je < m.nedges; je++) {
      const EdgeGeom eg = edge_geom(m, je);
      ((double *)out)[je] = n == "ex" ? eg.x : n == "ey" ? eg.y : n == "ea" ? eg.a : n == "enx" ? eg.nx : eg.ny;
    }
    return m.nedges;
  }
  if (n == "grad_cx" || n == "grad_cy" || n == "grad_c0x" || n == "grad_c0y") {  // the same for the gradient coefficients
    const GradOp &g = C->grad;
    const bool c0 = n == "grad_c0x" || n == "grad_c0y";
    if (c0 && g.form != 0) return 0;
    if (out) for (int ic = 0; ic < m.ncells; ic++) {
      double cx[kMaxStencil], cy[kMaxStencil], c0x, c0y;
      grad_cell_coeffs(m, g, ic, cx, cy, c0x, c0y);
      if (c0) ((double *)out)[ic] = n == "grad_c0x" ? c0x : c0y;
      else for (int64_t k = g.ptr[ic]; k < g.ptr[ic + 1]; k++) ((double *)out)[k] = n == "grad_cx" ? cx[k - g.ptr[ic]] : cy[k - g.ptr[ic]];
    }
    return c0 ? m.ncells : (long)g.ptr[m.ncells];
  }
  RET("en1", m.en1) RET("en2", m.en2) RET("ec1", m.ec1) RET("ec2", m.ec2) RET("cedge", m.cedge) RET("nghbre", m.nghbre)
  RET("cell_intr", m.cell_intr) RET("b_edge", m.b_edge) RET("b_edge_ptr", m.b_edge_ptr) RET("perm", C->L.perm)
  RET("loc2new", C->L.loc2new) RET("peers", C->L.peers) RET("send_ptr", C->L.send_ptr) RET("send_idx", C->L.send_idx)
  RET("recv_begin", C->L.recv_begin) RET("recv_count", C->L.recv_count)
  RET("f_off", C->L.f_off) RET("f_nbr", C->L.f_nbr) RET("f_edge", C->L.f_edge) RET("g_off", C->L.g_off) RET("g_idx", C->L.g_idx)
  RET("g_cx", C->L.g_cx) RET("g_cy", C->L.g_cy) RET("orig_id", C->L.orig_id) RET("bf_type", C->L.bf_type) RET("bf_edge", C->L.bf_edge)
  RET("lex", C->L.ex) RET("ley", C->L.ey) RET("is_intr", C->L.is_intr)
  RET("fz_tile_int", C->L.fz_tile_int) RET("fz_tile_bnd", C->L.fz_tile_bnd) RET("gh_ptr", C->L.gh_ptr) RET("gh_idx", C->L.gh_idx)
  RET("lea", C->L.ea) RET("lenx", C->L.enx) RET("leny", C->L.eny) RET("lxc", C->L.xc) RET("lyc", C->L.yc) RET("lvol", C->L.vol)
  RET("tile_es", C->L.tile_es) RET("tile_ne", C->L.tile_ne) RET("tile_hc_ptr", C->L.tile_hc_ptr) RET("tile_he_ptr", C->L.tile_he_ptr)
  RET("tile_hdr", C->L.tile_hdr) RET("t_pack", C->L.t_pack) RET("t_bf", C->L.t_bf)
  RET("tile_hc_idx", C->L.tile_hc_idx) RET("tile_he_idx", C->L.tile_he_idx) RET("f_pack", C->L.f_pack) RET("f_bf", C->L.f_bf)
  RET("grad_idx", C->grad.idx) RET("sub_orig", C->sub.orig) RET("sub_new_id", C->sub.new_id)
  if (n.rfind("fz_", 0) == 0) {  // tables of the fused stage kernel, built on first request (single rank)
    const std::string err = build_fused_tables(C->L);
    if (!err.empty()) { fail("%s", err.c_str()); return -1; }
    RET("fz_hdr", C->L.fz_hdr) RET("fz_h2_idx", C->L.fz_h2_idx) RET("fz_gslot", C->L.fz_gslot)
    RET("fz_pack2", C->L.fz_pack2) RET("fz_hf", C->L.fz_hf)
    if (n == "fz_gc") {  // coefficient rows at the pitch the device uses (cells padded to 32)
      std::vector<double> gc;
      fused_coeff_rows(C->L, (size_t)(C->L.n_loc + 31) / 32 * 32, gc);
      if (out) memcpy(out, gc.data(), gc.size() * 8);
      return (long)gc.size();
    }
    if (n == "fz_info") {
      const int info[8] = {C->L.fz_built, C->L.fz_w, C->L.fz_s2_max, C->L.fz_tw_max, C->L.fz_h2_max, C->L.fz_v2, C->L.fz_hf_max, 0};
      if (out) memcpy(out, info, sizeof info);
      return 8;
    }
  }
#undef RET
  if (n == "grad_ptr") {
    if (out) for (size_t i = 0; i < C->grad.ptr.size(); i++) ((int *)out)[i] = (int)C->grad.ptr[i];
    return (long)C->grad.ptr.size();
  }
  fail("fvs2d_gpu_mesh_array: unknown array '%s'", name);
  return -1;
}

int fvs2d_gpu_last_timing(double ms[4], long *launches) {
  NEED(C != nullptr, "fvs2d_gpu_last_timing: not initialised");
  for (int i = 0; i < 4; i++) ms[i] = C->last_ms[i];
  if (launches) *launches = C->last_launches;
  return 0;
}

int fvs2d_gpu_set_option(const char *key, int value) {
  NEED(C != nullptr, "fvs2d_gpu_set_option: not initialised");
  const std::string k = key ? key : "";
  if (k == "timing") { C->opt_timing = value; return 0; }
  if (k == "tile") { C->opt_tile = value; return 0; }
  if (k == "ctas") { C->opt_ctas = value; return 0; }
  if (k == "overlap") { C->opt_overlap = value; return 0; }
  if (k == "graph") { C->opt_graph = value; return 0; }
  if (k == "pair") { C->opt_pair = value; return 0; }
  if (k == "pdl") { C->opt_pdl = value; return 0; }
  if (k == "fuse") {
    if (value != C->opt_fuse) C->fz_state = 0;  // plan again under the new setting
    C->opt_fuse = value;
    return 0;
  }
  return fail("fvs2d_gpu_set_option: unknown option '%s'", key);
}

int fvs2d_gpu_finalize(void) {
  if (!C) return 0;
  if (C->inited) {
    cudaSetDevice(C->device);
    if (C->sx) cudaStreamSynchronize(C->sx);
    if (C->st) cudaStreamSynchronize(C->st);
    free_device();
  }
  for (auto &e : C->ev_pool) if (e) cudaEventDestroy(e);
  if (C->ev0) cudaEventDestroy(C->ev0);
  if (C->ev1) cudaEventDestroy(C->ev1);
  if (C->graph_exec) cudaGraphExecDestroy(C->graph_exec);
  if (C->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(C->comm);
  if (C->sx) { cudaStreamDestroy(C->sx); for (cudaEvent_t e : {C->e_a, C->e_g, C->e_b, C->e_p}) if (e) cudaEventDestroy(e); }
  if (C->st) cudaStreamDestroy(C->st);
  delete C;
  C = nullptr;
  return 0;
}

}  // extern "C"
