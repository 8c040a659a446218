// Host side of the B200 local-BA engine and its C-ABI (include/ppo_ba.h).
// One handle = one key-frame window resident in HBM + one CUDA stream.  The LM control loop
// (OptimizationAlgorithmLevenberg::solve, core/optimization_algorithm_levenberg.cpp:61-164) runs on
// the host and reads back three scalars per damped trial; everything else stays on the device.
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/ppo_ba.h"
#include "ppo_dense.h"
#include "ppo_kernels.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

using namespace ppo;

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      char buf_[512];                                                                              \
      snprintf(buf_, sizeof buf_, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      h->err = buf_;                                                                               \
      return PPO_E_CUDA;                                                                           \
    }                                                                                              \
  } while (0)

// ---- minimal NCCL binding (dlopen: the torch wheel already maps libnccl.so.2 into the process) -----
typedef struct ncclComm *ncclComm_t;
typedef enum { ncclSum_ = 0, ncclMax_ = 2 } ncclRedOp_t_;
typedef enum { ncclInt32_ = 2, ncclFloat64_ = 8 } ncclDataType_t_;
struct ncclUniqueId_ { char internal[128]; };
struct NcclApi {
  int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GetUniqueId)(ncclUniqueId_ *) = nullptr;
  int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId_, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  bool ok = false;
  void load() {
    if (ok) return;
    void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return;
    AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
    AllGather = (decltype(AllGather))dlsym(lib, "ncclAllGather");
    GetUniqueId = (decltype(GetUniqueId))dlsym(lib, "ncclGetUniqueId");
    CommInitRank = (decltype(CommInitRank))dlsym(lib, "ncclCommInitRank");
    CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
    GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
    ok = AllReduce && GetUniqueId && CommInitRank && CommDestroy;
  }
};
static NcclApi g_nccl;

struct ppo_ba_handle {
  ppo_ba_params P;
  int device = 0;
  cudaStream_t st = nullptr;
  cudaStream_t side[3] = {nullptr, nullptr, nullptr};  // plane / cuboid / point-cuboid edges are linearised next to the point edges
  cudaEvent_t ev_fork = nullptr, ev_join[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t evm[2] = {nullptr, nullptr};
  void *d_flush = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evp[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  std::string err;
  long long launches = 0;
  bool profiling = false;
  bool have_graph = false;
  std::vector<void *> allocs;
  DevGraph g;
  DevState sa, sb;  // current / trial (swapped on accept)
  DevState s0;      // estimates as given at set_graph (for reset)
  // host copies needed after set_graph
  int max_np = 0, ld = 0;
  int n_chunks = 0;
  int *d_kf_chunk_ptr = nullptr;
  double *d_chunk_part = nullptr;
  int *d_kf_iota_ptr = nullptr;  // 0, 1, 2, .. n_kf: "one chunk per key-frame" view of d_kf_part (sharded windows)
  double *d_kf_part = nullptr;   // n_kf x 27: per-key-frame sums of the chunk partials, summed across ranks
  unsigned *d_pair_keys = nullptr;            // sorted key-frame-pair keys of the Schur contributions
  unsigned long long *d_pair_vals = nullptr;  // (entry A << 32 | entry B)
  int n_pairs = 0;                            // upper bound (padding at the end of the sorted list)
  int *d_dup = nullptr;                       // set by k_pair_count: a landmark observed twice by one key-frame
  double *d_pair_bnd = nullptr;               // boundary records of the Schur pair list: 2 per 128-record chunk x 64 doubles (a C fragment)
  unsigned *d_pair_bnd_key = nullptr;
  int *d_pair_bnd_flag = nullptr;
  // cuboid-plane edges (constant residual): host side
  std::vector<int> cpe_cuboid, cpe_plane;
  std::vector<double> cpe_chi2, cpe_norm;
  std::vector<uint8_t> cpe_flags;
  std::vector<int32_t> outlier_idx;  // result of ppo_ba_point_edge_outliers (valid until its next call)
  int *d_cpe_cuboid = nullptr, *d_cpe_plane = nullptr;
  uint8_t *d_cpe_flags = nullptr;
  // partial-sum buffers
  double *d_chi_pt = nullptr, *d_chi_pl = nullptr, *d_chi_cb = nullptr, *d_chi_pc = nullptr, *d_scale_part = nullptr;
  int nb_lin = 0, nb_res = 0, nb_pl = 0, nb_cb = 0, nb_pc = 0, nb_bs = 0;
  Scalars *d_scal = nullptr, *h_scal = nullptr;
  int *d_not_spd = nullptr, *d_nout = nullptr;
  double *d_Winv = nullptr;
  double *d_S_bak = nullptr;   // PPO_SOLVER_6_3 (LinearSolverEigen semantics): the reduced system as it was before the Cholesky, for dense_ldlt_fallback
  void *d_dense_ws = nullptr;  // control block + tile version counters of the persistent factorisation
  int *h_dims = nullptr;
  // current mapping
  int n_p = 0, n_kf_free = 0, n_l = 0, n_active_edges = 0;
  // LM controller: state on the device (LmDev), inputs of a call (LmIn), host mirror of the final state
  LmDev *d_lm = nullptr;
  LmIn *d_lm_in = nullptr;
  LmDev *h_lm = nullptr;   // pinned
  LmIn *h_lm_in = nullptr; // pinned
  int *h_stop = nullptr;   // pinned + mapped: the caller's stop flag is forwarded here while a graph runs
  int *d_stop = nullptr;   // device view of h_stop
  // captured LM loops (one per (n_p, n_l) of a window): nested conditional WHILE nodes over the iteration / trial bodies
  struct LmGraph {
    int n_p, n_l, sm_cap;
    cudaGraphExec_t exec;
    cudaGraph_t graph;
    int nodes_iter, nodes_trial;
  };
  std::vector<LmGraph> lm_graphs;
  bool use_graph = true;
  int solve_sm_cap = 0;  // > 0: CTAs of the persistent factorisation (batch entry points: the windows of a batch share the SMs)
  long long host_syncs = 0;
  // sharding (multi-GPU, single window)
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  double *d_red = nullptr;  // 4 doubles for scalar allreduce
  long long collectives = 0;
  // distributed dense solve of a sharded window (ppo_dense.h): S | Winv | workspace live in ONE cudaMalloc'd arena that every
  // rank of the node maps through cudaIpc; the factorisation is spread over the ranks by tile column
  bool dist_solve = false;
  void *dist_arena = nullptr;
  size_t dist_arena_bytes = 0;
  void *dist_peer_base[DIST_MAX] = {};  // (other ranks: cudaIpcOpenMemHandle)
  DistPeers dist_peers;
  int dist_seq = 0;                     // solves so far (identical on every rank)
  unsigned *d_dist_ops = nullptr;       // this rank's worker queue for dist_ops_tc tile columns
  int dist_n_ops = 0, dist_ops_tc = -1;
  void dist_release() {
    for (int q = 0; q < DIST_MAX; q++) {
      if (dist_peer_base[q] && q != rank) cudaIpcCloseMemHandle(dist_peer_base[q]);
      dist_peer_base[q] = nullptr;
    }
    if (dist_arena) cudaFree(dist_arena);
    if (d_dist_ops) cudaFree(d_dist_ops);
    dist_arena = nullptr, d_dist_ops = nullptr, dist_arena_bytes = 0, dist_ops_tc = -1;
  }

  template <typename T>
  int dalloc(T **p, size_t n) {
    ppo_ba_handle *h = this;
    *p = nullptr;
    if (n == 0) n = 1;
    void *q = nullptr;
    CK(cudaMallocAsync(&q, n * sizeof(T), st));  // stream-ordered pool: re-used across windows without OS calls
    allocs.push_back(q);
    *p = (T *)q;
    return PPO_OK;
  }
  template <typename T>
  int upload(T **p, const std::vector<T> &v) {
    ppo_ba_handle *h = this;
    int rc = dalloc(p, v.size());
    if (rc) return rc;
    if (!v.empty()) {
      void *stage = nullptr;
      if ((rc = pinned(&stage, v.size() * sizeof(T)))) return rc;
      std::memcpy(stage, v.data(), v.size() * sizeof(T));
      CK(cudaMemcpyAsync(*p, stage, v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
    }
    return PPO_OK;
  }
  // Is the caller's array page-locked already (cudaMallocHost / cudaHostRegister)?  Then it is copied from where it lies: the caller's
  // arrays only have to stay valid until ppo_ba_set_graph returns, and set_graph drains the stream before it does.
  static bool is_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    return a.type == cudaMemoryTypeHost;
  }
  // caller's array -> (pinned staging unless it is pinned itself) -> device: one host pass at most, no intermediate std::vector
  template <typename T>
  int upload_raw(T **p, const T *src, size_t n) {
    ppo_ba_handle *h = this;
    int rc = dalloc(p, n);
    if (rc) return rc;
    if (n) {
      const void *from = src;
      if (n * sizeof(T) < 65536 || !is_pinned(src)) {
        void *stage = nullptr;
        if ((rc = pinned(&stage, n * sizeof(T)))) return rc;
        std::memcpy(stage, src, n * sizeof(T));
        from = stage;
      }
      CK(cudaMemcpyAsync(*p, from, n * sizeof(T), cudaMemcpyHostToDevice, st));
    }
    return PPO_OK;
  }
  // pinned host staging arena: every H2D / D2H copy of the C-ABI goes through page-locked memory
  char *hstage = nullptr;
  size_t hstage_cap = 0, hstage_off = 0;
  int pinned(void **out, size_t bytes) {
    ppo_ba_handle *h = this;
    bytes = (bytes + 255) & ~(size_t)255;
    if (hstage_off + bytes > hstage_cap) {
      // grow: the copies already enqueued read the old arena, so drain the stream before replacing it
      CK(cudaStreamSynchronize(st));
      if (hstage) cudaFreeHost(hstage);
      hstage_cap = std::max(hstage_cap * 2, hstage_off + bytes + (size_t)(64u << 20));
      hstage_off = 0;
      CK(cudaMallocHost((void **)&hstage, hstage_cap));
    }
    *out = hstage + hstage_off;
    hstage_off += bytes;
    return PPO_OK;
  }
  // A captured loop holds the window's device pointers and constants by value, so a new window (or new parameters) needs a new graph --
  // but not a new EXECUTABLE: cudaGraphExecUpdate accepts a re-captured graph of the same topology (new kernel arguments, launch
  // shapes and conditional handles; tools/ubench/cgraph_update_test.cu) in ~10 us where cudaGraphInstantiate takes ~0.6 ms.  The
  // executables of dropped loops are therefore kept (a few) and offered to build_lm_graph.
  std::vector<cudaGraphExec_t> spare_execs;
  void drop_lm_graphs(bool keep_execs = true) {
    for (auto &q : lm_graphs) {
      if (keep_execs) {
        if (spare_execs.size() >= 6) {  // (oldest out)
          cudaGraphExecDestroy(spare_execs.front());
          spare_execs.erase(spare_execs.begin());
        }
        spare_execs.push_back(q.exec);
      } else {
        cudaGraphExecDestroy(q.exec);
      }
      cudaGraphDestroy(q.graph);
    }
    lm_graphs.clear();
    if (!keep_execs) {
      for (auto &x : spare_execs) cudaGraphExecDestroy(x);
      spare_execs.clear();
    }
  }
  void free_graph() {
    drop_lm_graphs();  // they hold the device pointers of the window
    for (void *p : allocs) cudaFreeAsync(p, st);
    allocs.clear();
    have_graph = false;
  }
  // sharded window: landmarks (points, planes, their edges and the cuboid-plane edges of those planes) are rank-local; key-frames and
  // cuboids are replicated and the edges among them (camera-cuboid, point-cuboid) are accumulated by rank 0
  bool owner() const { return rank == 0; }
};

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// ---- batch variants: one host thread per window drives that window's LM loop on its own stream ------------------
template <typename F>
static int run_batch(int n, F &&one) {
  if (n < 0) return PPO_E_INVALID;
  std::vector<int> rc((size_t)std::max(n, 1), PPO_OK);
  std::vector<std::thread> th;
  th.reserve((size_t)std::max(n - 1, 0));
  for (int i = 1; i < n; i++) th.emplace_back([&rc, &one, i] { rc[i] = one(i); });
  if (n > 0) rc[0] = one(0);
  for (auto &t : th) t.join();
  for (int i = 0; i < n; i++)
    if (rc[i] != PPO_OK) return rc[i];
  return PPO_OK;
}

extern "C" {

void ppo_ba_default_params(ppo_ba_params *p) {
  std::memset(p, 0, sizeof *p);
  auto hd = [](double th) { return (double)(float)std::sqrt(th); };  // "const float th = sqrt(..)" in Optimizer.cc
  p->huber_mono = hd(5.991);
  p->huber_stereo = hd(7.815);
  p->huber_plane = hd(500.0);
  p->huber_vp_plane = hd(200.0);
  p->huber_bbox = hd(80.0);
  p->huber_corner = hd(10.0);
  p->huber_cuboid_plane = hd(500.0);
  p->chi2_mono = 5.991;
  p->chi2_stereo = 7.815;
  p->chi2_plane = 500.0;
  p->chi2_vp_plane = 200.0;
  p->norm_bbox = 80.0;
  p->norm_corner = 10.0;
  p->huber_se3 = 900.0;  // rk->setDelta(thHuberSE3), Optimizer.cc:1794; Parameters.cc:65
  p->norm_se3 = 900.0;
  p->lm_tau = 1e-5;
  p->lm_good_upper = 2. / 3.;
  p->lm_good_lower = 1. / 3.;
  p->lm_max_trials = 10;
  p->solver = PPO_SOLVER_DENSE_X;
  p->iters_round1 = 5;
  p->iters_round2 = 10;
  p->ptcu_max_outside_margin_ratio = 1.0;
  p->ptcu_prior_weight = 0.2;
}

int ppo_ba_create(const ppo_ba_params *params, int device, ppo_ba_handle **out) {
  if (!params || !out) return PPO_E_INVALID;
  int ndev = 0;
  if (device < 0) return PPO_E_INVALID;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device >= ndev) return PPO_E_NOGPU;
  ppo_ba_handle *h = new ppo_ba_handle();
  h->P = *params;
  h->device = device;
  std::memset(&h->g, 0, sizeof h->g);
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking) != cudaSuccess) {
    delete h;
    return PPO_E_CUDA;
  }
  for (auto &q : h->side) cudaStreamCreateWithFlags(&q, cudaStreamNonBlocking);
  cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
  for (auto &e : h->ev_join) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
  cudaEventCreate(&h->ev0);
  cudaEventCreate(&h->ev1);
  for (auto &e : h->evp) cudaEventCreate(&e);
  for (auto &e : h->evm) cudaEventCreate(&e);
  {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      unsigned long long keep = ~0ull;  // keep freed blocks cached in the pool
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  dense_setup_device(device);
  cudaMallocHost((void **)&h->h_scal, sizeof(Scalars));
  cudaMallocHost((void **)&h->h_dims, 8 * sizeof(int));
  cudaMallocHost((void **)&h->h_lm, sizeof(LmDev));
  cudaMallocHost((void **)&h->h_lm_in, sizeof(LmIn));
  cudaHostAlloc((void **)&h->h_stop, sizeof(int), cudaHostAllocMapped);
  if (h->h_stop) {
    *h->h_stop = 0;
    cudaHostGetDevicePointer((void **)&h->d_stop, h->h_stop, 0);
  }
  h->use_graph = std::getenv("PPO_BA_NO_GRAPH") == nullptr;
  *out = h;
  return PPO_OK;
}

void ppo_ba_destroy(ppo_ba_handle *h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->st);
  h->free_graph();
  h->drop_lm_graphs(false);  // ... and the spare executables
  cudaStreamSynchronize(h->st);
  h->dist_release();
  if (h->hstage) cudaFreeHost(h->hstage);
  cudaFreeHost(h->h_scal);
  cudaFreeHost(h->h_dims);
  cudaFreeHost(h->h_lm);
  cudaFreeHost(h->h_lm_in);
  if (h->h_stop) cudaFreeHost(h->h_stop);
  for (auto &q : h->side) {
    cudaStreamSynchronize(q);
    cudaStreamDestroy(q);
  }
  cudaEventDestroy(h->ev_fork);
  for (auto &e : h->ev_join) cudaEventDestroy(e);
  cudaEventDestroy(h->ev0);
  cudaEventDestroy(h->ev1);
  for (auto &e : h->evp) cudaEventDestroy(e);
  for (auto &e : h->evm) cudaEventDestroy(e);
  if (h->d_flush) cudaFree(h->d_flush);
  cudaStreamDestroy(h->st);
  delete h;
}

const char *ppo_ba_last_error(const ppo_ba_handle *h) { return h ? h->err.c_str() : "null handle"; }

int ppo_ba_set_params(ppo_ba_handle *h, const ppo_ba_params *params) {
  if (!h || !params) return PPO_E_INVALID;
  h->P = *params;
  if (h->have_graph) {  // constants the resident window carries (ppo_ba_set_graph copies them)
    DevGraph &g = h->g;
    g.huber_mono = h->P.huber_mono; g.huber_stereo = h->P.huber_stereo; g.huber_plane = h->P.huber_plane; g.huber_vp = h->P.huber_vp_plane;
    g.huber_bbox = h->P.huber_bbox; g.huber_corner = h->P.huber_corner; g.huber_se3 = h->P.huber_se3;
    g.ptcu_ratio = h->P.ptcu_max_outside_margin_ratio; g.ptcu_prior = h->P.ptcu_prior_weight;
    h->drop_lm_graphs();  // the captured kernels hold the old constants by value
    if (h->P.solver == PPO_SOLVER_6_3 && !h->d_S_bak && !h->dist_solve) {
      int rc = h->dalloc(&h->d_S_bak, dense_matrix_doubles(h->max_np));
      if (rc) return rc;
    }
  }
  return PPO_OK;
}

int ppo_ba_edge_count(const ppo_ba_handle *h, int kind) {
  switch (kind) {
    case PPO_EDGE_POINT: return h->g.n_pe;
    case PPO_EDGE_PLANE: return h->g.n_ple;
    case PPO_EDGE_CUBOID_CAM: return h->g.n_cbe;
    case PPO_EDGE_POINT_CUBOID: return h->g.n_pce;
    case PPO_EDGE_CUBOID_PLANE: return h->g.n_cpe;
  }
  return -1;
}

static int alloc_state(ppo_ba_handle *h, DevState *s) {
  int rc;
  if ((rc = h->dalloc(&s->kf_pose, 7 * (size_t)h->g.n_kf))) return rc;
  if ((rc = h->dalloc(&s->kf_Rt, 12 * (size_t)h->g.n_kf))) return rc;
  if ((rc = h->dalloc(&s->pt, 3 * (size_t)h->g.n_pt))) return rc;
  if ((rc = h->dalloc(&s->pl, 4 * (size_t)h->g.n_pl))) return rc;
  if ((rc = h->dalloc(&s->cu, 10 * (size_t)h->g.n_cu))) return rc;
  return PPO_OK;
}
static int copy_state(ppo_ba_handle *h, const DevState &dst, const DevState &src) {
  const DevGraph &g = h->g;
  CK(cudaMemcpyAsync(dst.kf_pose, src.kf_pose, 7 * sizeof(double) * g.n_kf, cudaMemcpyDeviceToDevice, h->st));
  CK(cudaMemcpyAsync(dst.kf_Rt, src.kf_Rt, 12 * sizeof(double) * g.n_kf, cudaMemcpyDeviceToDevice, h->st));
  CK(cudaMemcpyAsync(dst.pt, src.pt, 3 * sizeof(double) * g.n_pt, cudaMemcpyDeviceToDevice, h->st));
  CK(cudaMemcpyAsync(dst.pl, src.pl, 4 * sizeof(double) * g.n_pl, cudaMemcpyDeviceToDevice, h->st));
  CK(cudaMemcpyAsync(dst.cu, src.cu, 10 * sizeof(double) * g.n_cu, cudaMemcpyDeviceToDevice, h->st));
  return PPO_OK;
}

static void recompute_cpe(ppo_ba_handle *h, const double *meas, const double *info) {
  // EdgeCuboidPlane::computeError: _error = _measurement (G2O_Plane3D.h:470-473)
  for (size_t e = 0; e < h->cpe_chi2.size(); e++) {
    double c = 0, n = 0;
    for (int i = 0; i < 3; i++) c += meas[3 * e + i] * (info[3 * e + i] * meas[3 * e + i]), n += meas[3 * e + i] * meas[3 * e + i];
    h->cpe_chi2[e] = c;
    h->cpe_norm[e] = std::sqrt(n);
  }
}
static double cpe_chi_const(const ppo_ba_handle *h) {
  if (!h->owner()) return 0.0;
  double s = 0;
  for (size_t e = 0; e < h->cpe_chi2.size(); e++) {
    if (h->cpe_flags[e] & PPO_EF_LEVEL1) continue;
    double c = h->cpe_chi2[e];
    if (h->cpe_flags[e] & PPO_EF_ROBUST) {
      const double d = h->P.huber_cuboid_plane, dsqr = (double)(float)(d * d);  // float dsqr of RobustKernelHuber (robust_kernel_impl.h:84)
      if (c > dsqr) c = 2 * std::sqrt(c) * d - dsqr;
    }
    s += c;
  }
  return s;
}

// ---- peer-mapped arena of the distributed dense solve (collective: every rank of the sharded window calls set_graph) ----------
static int dist_setup(ppo_ba_handle *h) {
  auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t sS = al(dense_matrix_doubles(h->max_np) * 8), sW = al((size_t)dense_num_blocks(h->max_np) * DENSE_TILE * 8), sWs = al(dense_workspace_bytes(h->max_np));
  const size_t bytes = sS + sW + sWs;
  if (!g_nccl.AllGather) {
    h->err = "ncclAllGather not found in libnccl";
    return PPO_E_NCCL;
  }
  if (bytes != h->dist_arena_bytes) {  // (same decision on every rank: key-frames and cuboids are replicated, so max_np agrees)
    CK(cudaStreamSynchronize(h->st));
    h->dist_release();
    CK(cudaMalloc(&h->dist_arena, bytes));
    h->dist_arena_bytes = bytes;
    // exchange {cudaIpc handle, arena size} over the window's NCCL communicator
    struct Slot {
      cudaIpcMemHandle_t hd;
      unsigned long long bytes;
      unsigned long long pad;
    };
    static_assert(sizeof(Slot) == 80, "slot layout");
    std::vector<Slot> slots((size_t)h->world);
    std::memset(slots.data(), 0, sizeof(Slot) * slots.size());
    CK(cudaIpcGetMemHandle(&slots[h->rank].hd, h->dist_arena));
    slots[h->rank].bytes = bytes;
    Slot *d_slots = nullptr;
    CK(cudaMalloc((void **)&d_slots, sizeof(Slot) * slots.size()));
    CK(cudaMemcpyAsync(d_slots, slots.data(), sizeof(Slot) * slots.size(), cudaMemcpyHostToDevice, h->st));
    const int r = g_nccl.AllGather(d_slots + h->rank, d_slots, sizeof(Slot), /*ncclInt8*/ 0, h->comm, h->st);
    if (r != 0) {
      h->err = std::string("ncclAllGather: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error");
      cudaFree(d_slots);
      return PPO_E_NCCL;
    }
    h->collectives++;
    CK(cudaMemcpyAsync(slots.data(), d_slots, sizeof(Slot) * slots.size(), cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    cudaFree(d_slots);
    h->dist_peers.rank = h->rank, h->dist_peers.world = h->world;
    h->dist_peers.blk = std::getenv("PPO_DIST_BLOCK") ? std::max(1, std::atoi(std::getenv("PPO_DIST_BLOCK"))) : DIST_BLOCK_DEFAULT;
    h->dist_peers.fwd = std::getenv("PPO_DIST_FORWARD") ? (std::atoi(std::getenv("PPO_DIST_FORWARD")) != 0) : DIST_FORWARD_DEFAULT;
    for (int q = 0; q < h->world; q++) {
      if (slots[q].bytes != bytes) {
        h->err = "sharded window: the ranks disagree on the size of the reduced system";
        return PPO_E_INVALID;
      }
      void *base = h->dist_arena;
      if (q != h->rank) CK(cudaIpcOpenMemHandle(&base, slots[q].hd, cudaIpcMemLazyEnablePeerAccess));
      h->dist_peer_base[q] = base;
      char *b = reinterpret_cast<char *>(base);
      dense_dist_set_peer(&h->dist_peers, q, reinterpret_cast<double *>(b), reinterpret_cast<double *>(b + sS), b + sS + sW);
    }
    dense_workspace_init(reinterpret_cast<char *>(h->dist_arena) + sS + sW, h->max_np, h->st);
  }
  char *b = reinterpret_cast<char *>(h->dist_arena);
  h->g.S = reinterpret_cast<double *>(b);
  h->d_Winv = reinterpret_cast<double *>(b + sS);
  h->d_dense_ws = b + sS + sW;
  return PPO_OK;
}

static int set_graph_impl(ppo_ba_handle *h, const ppo_ba_graph *gi) {
  if (!h || !gi) return PPO_E_INVALID;
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->st));
  h->free_graph();
  h->hstage_off = 0;
  DevGraph &g = h->g;
  std::memset(&g, 0, sizeof g);
  // PPO_BA_TIMING=1: host-side phase times of this call on stderr (diagnostics only)
  static const bool timing = std::getenv("PPO_BA_TIMING") != nullptr;
  auto t_prev = std::chrono::steady_clock::now();
  auto tick = [&](const char *what) {
    if (!timing) return;
    const auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[ppo_ba_set_graph] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_prev).count());
    t_prev = now;
  };
  if (gi->n_kf <= 0 || gi->n_pt < 0 || gi->n_pl < 0 || gi->n_cu < 0) { h->err = "bad vertex counts"; return PPO_E_INVALID; }
  g.n_kf = gi->n_kf; g.n_pt = gi->n_pt; g.n_pl = gi->n_pl; g.n_cu = gi->n_cu;
  g.n_pe = gi->n_pe; g.n_ple = gi->n_ple; g.n_cbe = gi->n_cbe; g.n_pce = gi->n_pce; g.n_cpe = gi->n_cpe;
  if (g.n_pt > 0 && (!gi->pt_rowptr || gi->pt_rowptr[0] != 0 || gi->pt_rowptr[g.n_pt] != g.n_pe)) { h->err = "pt_rowptr inconsistent with n_pe"; return PPO_E_INVALID; }
  if (g.n_pt == 0 && g.n_pe != 0) { h->err = "point edges without points"; return PPO_E_INVALID; }
  for (int p = 0; p < g.n_pt; p++) if (gi->pt_rowptr[p + 1] < gi->pt_rowptr[p]) { h->err = "pt_rowptr not monotone"; return PPO_E_INVALID; }
  for (int e = 0; e < g.n_ple; e++)
    if (gi->ple_kf[e] < 0 || gi->ple_kf[e] >= g.n_kf || gi->ple_plane[e] < 0 || gi->ple_plane[e] >= g.n_pl || gi->ple_kind[e] > 2) { h->err = "plane edge out of range"; return PPO_E_INVALID; }
  for (int e = 0; e < g.n_cbe; e++)
    if (gi->cbe_kf[e] < 0 || gi->cbe_kf[e] >= g.n_kf || gi->cbe_cuboid[e] < 0 || gi->cbe_cuboid[e] >= g.n_cu || gi->cbe_kind[e] > PPO_CUBOID_SE3) { h->err = "cuboid edge out of range"; return PPO_E_INVALID; }
  for (int e = 0; e < g.n_pce; e++)
    if (gi->pce_cuboid[e] < 0 || gi->pce_cuboid[e] >= g.n_cu || gi->pce_rowptr[e + 1] < gi->pce_rowptr[e]) { h->err = "point-cuboid edge out of range"; return PPO_E_INVALID; }
  for (int e = 0; e < g.n_cpe; e++)
    if (gi->cpe_cuboid[e] < 0 || gi->cpe_cuboid[e] >= g.n_cu || gi->cpe_plane[e] < 0 || gi->cpe_plane[e] >= g.n_pl) { h->err = "cuboid-plane edge out of range"; return PPO_E_INVALID; }

  int rc;
  tick("validate");
#define UP(dst, vec) if ((rc = h->upload(&(dst), (vec)))) return rc
  // ---- vertices: constants ------------------------------------------------------------------------
  std::vector<uint8_t> kf_fixed(gi->kf_fixed, gi->kf_fixed + g.n_kf), pt_fixed(g.n_pt, 0), cu_flags(g.n_cu, 0);
  if (gi->pt_fixed) pt_fixed.assign(gi->pt_fixed, gi->pt_fixed + g.n_pt);
  if (g.n_cu) cu_flags.assign(gi->cu_flags, gi->cu_flags + g.n_cu);
  std::vector<float> kf_intr(gi->kf_intr, gi->kf_intr + 5 * (size_t)g.n_kf);
  { uint8_t *p; UP(p, kf_fixed); g.kf_fixed = p; UP(p, pt_fixed); g.pt_fixed = p; UP(p, cu_flags); g.cu_flags = p; }
  { float *p; UP(p, kf_intr); g.kf_intr = p; }
  // ---- estimates, normalised as the g2o vertex setters do -------------------------------------------
  std::vector<double> kf_pose(7 * (size_t)g.n_kf), pl(4 * (size_t)g.n_pl), cu(10 * (size_t)g.n_cu);
  auto norm_q = [](double *q) {  // SE3Quat::normalizeRotation
    double s = q[3] < 0 ? -1.0 : 1.0;
    double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    if (q[3] < 0) for (int i = 0; i < 4; i++) q[i] = -q[i];
    (void)s;
    for (int i = 0; i < 4; i++) q[i] /= n;
  };
  for (int i = 0; i < g.n_kf; i++) {
    for (int k = 0; k < 7; k++) kf_pose[7 * (size_t)i + k] = gi->kf_pose[7 * (size_t)i + k];
    norm_q(&kf_pose[7 * (size_t)i]);
  }
  for (int i = 0; i < g.n_pl; i++) {  // Plane3D::fromVector -> normalize
    double *c = &pl[4 * (size_t)i];
    for (int k = 0; k < 4; k++) c[k] = gi->pl_coef[4 * (size_t)i + k];
    double inv = 1. / std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
    for (int k = 0; k < 4; k++) c[k] = c[k] * inv;
    if (c[3] < 0.0) for (int k = 0; k < 4; k++) c[k] = -c[k];
  }
  for (int i = 0; i < g.n_cu; i++) {
    for (int k = 0; k < 10; k++) cu[10 * (size_t)i + k] = gi->cu_state[10 * (size_t)i + k];
    norm_q(&cu[10 * (size_t)i + 3]);
  }
  if ((rc = alloc_state(h, &h->sa)) || (rc = alloc_state(h, &h->sb)) || (rc = alloc_state(h, &h->s0))) return rc;
  auto stage_up = [&](double *dst, const double *src, size_t n) -> int {
    if (n == 0) return PPO_OK;
    const void *from = src;
    if (n * 8 < 65536 || !ppo_ba_handle::is_pinned(src)) {  // (the local vectors above are small and pageable; pt_xyz may be the caller's pinned array)
      void *stage = nullptr;
      int r = h->pinned(&stage, n * 8);
      if (r) return r;
      std::memcpy(stage, src, n * 8);
      from = stage;
    }
    CK(cudaMemcpyAsync(dst, from, n * 8, cudaMemcpyHostToDevice, h->st));
    return PPO_OK;
  };
  if ((rc = stage_up(h->s0.kf_pose, kf_pose.data(), kf_pose.size())) || (rc = stage_up(h->s0.pt, gi->pt_xyz, 3 * (size_t)g.n_pt)) ||
      (rc = stage_up(h->s0.pl, pl.data(), pl.size())) || (rc = stage_up(h->s0.cu, cu.data(), cu.size())))
    return rc;
  k_pose_cache<<<cdiv(g.n_kf, 128), 128, 0, h->st>>>(g.n_kf, h->s0.kf_pose, h->s0.kf_Rt);
  h->launches++;
  if ((rc = copy_state(h, h->sa, h->s0))) return rc;
  if ((rc = copy_state(h, h->sb, h->s0))) return rc;

  tick("vertices");
  // ---- point edges --------------------------------------------------------------------------------
  // The caller's arrays go to the device as they are (one memcpy into pinned staging each); the packed edge
  // records, the edge -> point map and the per-key-frame edge lists are built there.  The host only makes one
  // pass over the key-frame ids (range check + histogram, which sizes the per-key-frame chunk lists).
  std::vector<int> kf_cnt(g.n_kf + 1, 0);
  for (int e = 0; e < g.n_pe; e++) {
    const unsigned kf = (unsigned)gi->pe_kf[e];
    if (kf >= (unsigned)g.n_kf) { h->err = "pe_kf out of range"; return PPO_E_INVALID; }
    kf_cnt[kf + 1]++;
  }
  const int *rowptr = gi->pt_rowptr;
  {
    int *d_rowptr = nullptr, *d_pe_kf = nullptr, *d_pe_pt = nullptr, *d_iota = nullptr, *d_kfe = nullptr;
    unsigned *d_kf_sorted = nullptr;
    float *d_obs = nullptr, *d_is2 = nullptr;
    PointEdgeRec *d_rec = nullptr;
    const int one_zero[1] = {0};
    if ((rc = h->upload_raw(&d_rowptr, g.n_pt ? rowptr : one_zero, (size_t)g.n_pt + 1)) || (rc = h->upload_raw(&d_pe_kf, gi->pe_kf, (size_t)g.n_pe)) ||
        (rc = h->upload_raw(&d_obs, gi->pe_obs, 3 * (size_t)g.n_pe)) || (rc = h->upload_raw(&d_is2, gi->pe_invsigma2, (size_t)g.n_pe)) ||
        (rc = h->dalloc(&d_pe_pt, (size_t)g.n_pe)) || (rc = h->dalloc(&d_rec, (size_t)g.n_pe)) || (rc = h->dalloc(&d_iota, (size_t)g.n_pe)) ||
        (rc = h->dalloc(&d_kfe, (size_t)g.n_pe)) || (rc = h->dalloc(&d_kf_sorted, (size_t)g.n_pe)))
      return rc;
    g.pt_rowptr = d_rowptr; g.pe_pt = d_pe_pt; g.pe_rec = d_rec; g.pe_is2 = d_is2; g.kfe_edge = d_kfe;
    if (g.n_pe) {
      k_pack_point_edges<<<cdiv(g.n_pe, 256), 256, 0, h->st>>>(g.n_pe, d_pe_kf, d_obs, d_rec, d_iota);
      k_fill_pe_pt<<<cdiv(g.n_pt, 256), 256, 0, h->st>>>(g.n_pt, d_rowptr, d_pe_pt);
      // edges grouped by key-frame: stable radix sort of the edge ids by key-frame id (same order as a counting sort)
      int bits = 1;
      while ((1 << bits) < g.n_kf) bits++;
      size_t tmp_bytes = 0;
      CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (const unsigned *)d_pe_kf, d_kf_sorted, (const int *)d_iota, d_kfe, g.n_pe, 0, bits, h->st));
      char *tmp = nullptr;
      if ((rc = h->dalloc(&tmp, tmp_bytes))) return rc;
      CK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, (const unsigned *)d_pe_kf, d_kf_sorted, (const int *)d_iota, d_kfe, g.n_pe, 0, bits, h->st));
      h->launches += 4;
    }
  }
  tick("point edge uploads");
  // work units: consecutive points, <= 32 edges per warp (a point with > 32 edges is its own unit)
  std::vector<int> unit_pt0;
  unit_pt0.push_back(0);
  {
    int cnt = 0;
    for (int p = 0; p < g.n_pt; p++) {
      const int k = rowptr[p + 1] - rowptr[p];
      if (cnt > 0 && cnt + k > 32) {
        unit_pt0.push_back(p);
        cnt = 0;
      }
      cnt += k;
    }
    if (g.n_pt > 0) unit_pt0.push_back(g.n_pt);
  }
  g.n_units = (int)unit_pt0.size() - 1;
  { int *p; UP(p, unit_pt0); g.unit_pt0 = p; }
  {
    std::vector<int> unit_e0(unit_pt0.size());
    for (size_t u = 0; u < unit_pt0.size(); u++) unit_e0[u] = g.n_pt ? rowptr[unit_pt0[u]] : 0;
    int *p; UP(p, unit_e0); g.unit_e0 = p;
  }
  tick("work units");
  // per-key-frame chunks of <= POSE_THREADS edges of the sorted list
  for (int i = 0; i < g.n_kf; i++) kf_cnt[i + 1] += kf_cnt[i];
  std::vector<int> chunk_kf, chunk_b, chunk_e, kf_chunk_ptr(g.n_kf + 1, 0);
  for (int i = 0; i < g.n_kf; i++) {
    kf_chunk_ptr[i] = (int)chunk_kf.size();
    if (!kf_fixed[i])
      for (int b = kf_cnt[i]; b < kf_cnt[i + 1]; b += POSE_THREADS) {
        chunk_kf.push_back(i);
        chunk_b.push_back(b);
        chunk_e.push_back(std::min(b + POSE_THREADS, kf_cnt[i + 1]));
      }
  }
  kf_chunk_ptr[g.n_kf] = (int)chunk_kf.size();
  g.n_chunks = h->n_chunks = (int)chunk_kf.size();
  { int *p; UP(p, chunk_kf); g.chunk_kf = p; UP(p, chunk_b); g.chunk_begin = p; UP(p, chunk_e); g.chunk_end = p; }
  UP(h->d_kf_chunk_ptr, kf_chunk_ptr);
  if ((rc = h->dalloc(&h->d_chunk_part, 27 * (size_t)g.n_chunks))) return rc;
  if ((rc = h->dalloc(&h->d_kf_part, 27 * (size_t)g.n_kf))) return rc;
  {
    std::vector<int> iota((size_t)g.n_kf + 1);
    for (int i = 0; i <= g.n_kf; i++) iota[i] = i;
    UP(h->d_kf_iota_ptr, iota);
  }

  tick("edges by key-frame");
  // ---- plane edges: slots = unique (plane, key-frame) pairs, sorted by (plane, kf) ---------------------
  // bucket the edges by plane (counting sort), then sort / deduplicate the few key-frames of each plane
  g.n_lm = g.n_pl + g.n_pt;
  std::vector<int> ple_slot(g.n_ple), slot_kf, lm_rowptr(g.n_lm + 1, 0);
  {
    std::vector<int> first(g.n_pl + 1, 0), kfs(g.n_ple);
    for (int e = 0; e < g.n_ple; e++) first[gi->ple_plane[e] + 1]++;
    for (int p = 0; p < g.n_pl; p++) first[p + 1] += first[p];
    {
      std::vector<int> pos(first.begin(), first.end() - 1);
      for (int e = 0; e < g.n_ple; e++) kfs[pos[gi->ple_plane[e]]++] = gi->ple_kf[e];
    }
    std::vector<int> slot0(g.n_pl + 1, 0);
    slot_kf.reserve(g.n_ple);
    for (int p = 0; p < g.n_pl; p++) {
      std::sort(kfs.begin() + first[p], kfs.begin() + first[p + 1]);
      slot0[p] = (int)slot_kf.size();
      for (int i = first[p]; i < first[p + 1]; i++)
        if (i == first[p] || kfs[i] != kfs[i - 1]) slot_kf.push_back(kfs[i]);
      lm_rowptr[p + 1] = (int)slot_kf.size() - slot0[p];
    }
    slot0[g.n_pl] = (int)slot_kf.size();
    for (int e = 0; e < g.n_ple; e++) {
      const int p = gi->ple_plane[e];
      ple_slot[e] = (int)(std::lower_bound(slot_kf.begin() + slot0[p], slot_kf.begin() + slot0[p + 1], gi->ple_kf[e]) - slot_kf.begin());
    }
  }
  g.n_slots = (int)slot_kf.size();
  g.n_ent = g.n_slots + g.n_pe;
  for (int p = 0; p < g.n_pl; p++) lm_rowptr[p + 1] += lm_rowptr[p];
  for (int p = 0; p < g.n_pt; p++) lm_rowptr[g.n_pl + p + 1] = g.n_slots + rowptr[p + 1];
  {
    int *p;
    std::vector<int> v(gi->ple_plane, gi->ple_plane + g.n_ple); UP(p, v); g.ple_plane = p;
    v.assign(gi->ple_kf, gi->ple_kf + g.n_ple); UP(p, v); g.ple_kf = p;
    UP(p, ple_slot); g.ple_slot = p;
    UP(p, slot_kf); g.slot_kf = p;
    UP(p, lm_rowptr); g.lm_rowptr = p;
    uint8_t *q;
    std::vector<uint8_t> k(gi->ple_kind, gi->ple_kind + g.n_ple); UP(q, k); g.ple_kind = q;
    // measurements normalised like setMeasurement(Plane3D(v)) -> fromVector
    std::vector<double> meas(4 * (size_t)g.n_ple);
    for (int e = 0; e < g.n_ple; e++) {
      double *c = &meas[4 * (size_t)e];
      for (int i = 0; i < 4; i++) c[i] = gi->ple_meas[4 * (size_t)e + i];
      double inv = 1. / std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
      for (int i = 0; i < 4; i++) c[i] = c[i] * inv;
      if (c[3] < 0.0) for (int i = 0; i < 4; i++) c[i] = -c[i];
    }
    double *d;
    UP(d, meas); g.ple_meas = d;
    std::vector<double> info(gi->ple_info, gi->ple_info + 3 * (size_t)g.n_ple); UP(d, info); g.ple_info = d;
  }
  tick("plane edges");
  // ---- Schur contribution list: one record per pair of free-key-frame blocks of a landmark, sorted by key-frame pair ----
  {
    if ((unsigned long long)g.n_kf * (unsigned long long)g.n_kf >= 0xffffffffull) { h->err = "too many key-frames for 32-bit pair keys"; return PPO_E_INVALID; }
    // Upper bound of the list length from the block counts alone; pairs that involve a FIXED key-frame are not
    // emitted, their slots stay as end-of-list padding (key 0xffffffff sorts last, k_schur_pairs skips it).
    long long total = 0;
    for (int L = 0; L < g.n_lm; L++) {
      const long long k = lm_rowptr[L + 1] - lm_rowptr[L];
      total += k * (k + 1) / 2;
      if (total > 0x7fffffffll) { h->err = "Schur contribution list exceeds 2^31 entries"; return PPO_E_INVALID; }
    }
    h->n_pairs = (int)total;
    {
      const size_t nch = (size_t)cdiv((int)total, PAIR_CHUNK) + 1;
      if ((rc = h->dalloc(&h->d_pair_bnd, nch * 2 * 64)) || (rc = h->dalloc(&h->d_pair_bnd_key, nch * 2)) || (rc = h->dalloc(&h->d_pair_bnd_flag, nch))) return rc;
    }
    int *d_cnt = nullptr, *d_off = nullptr;
    unsigned *k_in = nullptr;
    unsigned long long *v_in = nullptr;
    if ((rc = h->dalloc(&d_cnt, (size_t)g.n_lm + 1)) || (rc = h->dalloc(&d_off, (size_t)g.n_lm + 1)) || (rc = h->dalloc(&h->d_dup, 1)) ||
        (rc = h->dalloc(&k_in, (size_t)total)) || (rc = h->dalloc(&v_in, (size_t)total)) || (rc = h->dalloc(&h->d_pair_keys, (size_t)total)) ||
        (rc = h->dalloc(&h->d_pair_vals, (size_t)total)))
      return rc;
    CK(cudaMemsetAsync(h->d_dup, 0, sizeof(int), h->st));
    if (total > 0) {
      CK(cudaMemsetAsync(k_in, 0xff, 4 * (size_t)total, h->st));
      CK(cudaMemsetAsync(d_cnt, 0, 4 * ((size_t)g.n_lm + 1), h->st));
      k_pair_count<<<cdiv(g.n_lm, 128), 128, 0, h->st>>>(g, d_cnt, h->d_dup);
      size_t scan_bytes = 0, sort_bytes = 0;
      int bits = 1;
      while (bits < 32 && (1ull << bits) <= (unsigned long long)g.n_kf * (unsigned long long)g.n_kf) bits++;  // padding > every key
      CK(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, d_cnt, d_off, g.n_lm + 1, h->st));
      CK(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, k_in, h->d_pair_keys, v_in, h->d_pair_vals, (int)total, 0, bits, h->st));
      char *tmp = nullptr;
      if ((rc = h->dalloc(&tmp, std::max(scan_bytes, sort_bytes)))) return rc;
      CK(cub::DeviceScan::ExclusiveSum(tmp, scan_bytes, d_cnt, d_off, g.n_lm + 1, h->st));
      k_gen_pairs<<<cdiv(g.n_lm, 4), 128, 0, h->st>>>(g, d_off, k_in, v_in);
      CK(cub::DeviceRadixSort::SortPairs(tmp, sort_bytes, k_in, h->d_pair_keys, v_in, h->d_pair_vals, (int)total, 0, bits, h->st));
      h->launches += 7;
    }
  }
  tick("Schur pair list");
  // ---- camera-cuboid / point-cuboid / cuboid-plane edges ---------------------------------------------
  {
    int *p; uint8_t *q; double *d;
    std::vector<int> v(gi->cbe_kf, gi->cbe_kf + g.n_cbe); UP(p, v); g.cbe_kf = p;
    v.assign(gi->cbe_cuboid, gi->cbe_cuboid + g.n_cbe); UP(p, v); g.cbe_cuboid = p;
    std::vector<uint8_t> k(gi->cbe_kind, gi->cbe_kind + g.n_cbe); UP(q, k); g.cbe_kind = q;
    std::vector<double> m(gi->cbe_meas, gi->cbe_meas + 16 * (size_t)g.n_cbe); UP(d, m); g.cbe_meas = d;
    m.assign(gi->cbe_info, gi->cbe_info + g.n_cbe); UP(d, m); g.cbe_info = d;
    v.assign(gi->pce_cuboid, gi->pce_cuboid + g.n_pce); UP(p, v); g.pce_cuboid = p;
    if (g.n_pce) v.assign(gi->pce_rowptr, gi->pce_rowptr + g.n_pce + 1); else v.assign(1, 0);
    const int npts = v.back();
    UP(p, v); g.pce_rowptr = p;
    if (npts) m.assign(gi->pce_pts, gi->pce_pts + 3 * (size_t)npts); else m.clear();
    UP(d, m); g.pce_pts = d;
  }
  h->cpe_cuboid.assign(gi->cpe_cuboid, gi->cpe_cuboid + g.n_cpe);
  h->cpe_plane.assign(gi->cpe_plane, gi->cpe_plane + g.n_cpe);
  h->cpe_chi2.assign(g.n_cpe, 0.0);
  h->cpe_norm.assign(g.n_cpe, 0.0);
  h->cpe_flags.assign(g.n_cpe, PPO_EF_ROBUST);
  if (g.n_cpe) recompute_cpe(h, gi->cpe_meas, gi->cpe_info);
  UP(h->d_cpe_cuboid, h->cpe_cuboid);
  UP(h->d_cpe_plane, h->cpe_plane);
  UP(h->d_cpe_flags, h->cpe_flags);

  {  // gather lists of the non-point edges (counting sorts; a few 10^4 entries): the assembly reduces per-edge contributions in list order
    auto csr = [&](int n_rows, int n, auto row_of, const int **ptr_out, const int **idx_out) -> int {
      std::vector<int> ptr((size_t)n_rows + 1, 0), idx((size_t)n);
      for (int e = 0; e < n; e++) ptr[(size_t)row_of(e) + 1]++;
      for (int r = 0; r < n_rows; r++) ptr[(size_t)r + 1] += ptr[r];
      std::vector<int> pos(ptr.begin(), ptr.end() - 1);
      for (int e = 0; e < n; e++) idx[pos[row_of(e)]++] = e;
      int *p = nullptr, *q = nullptr;
      int r2;
      if ((r2 = h->upload(&p, ptr)) || (r2 = h->upload(&q, idx))) return r2;
      *ptr_out = p, *idx_out = q;
      return PPO_OK;
    };
    if ((rc = csr(g.n_kf, g.n_ple, [&](int e) { return gi->ple_kf[e]; }, &g.kf_ple_ptr, &g.kf_ple_idx)) ||
        (rc = csr(g.n_kf, g.n_cbe, [&](int e) { return gi->cbe_kf[e]; }, &g.kf_cbe_ptr, &g.kf_cbe_idx)) ||
        (rc = csr(g.n_cu, g.n_cbe, [&](int e) { return gi->cbe_cuboid[e]; }, &g.cu_cbe_ptr, &g.cu_cbe_idx)) ||
        (rc = csr(g.n_cu, g.n_pce, [&](int e) { return gi->pce_cuboid[e]; }, &g.cu_pce_ptr, &g.cu_pce_idx)) ||
        (rc = csr(g.n_pl, g.n_ple, [&](int e) { return gi->ple_plane[e]; }, &g.pl_ple_ptr, &g.pl_ple_idx)) ||
        (rc = csr(g.n_slots, g.n_ple, [&](int e) { return ple_slot[e]; }, &g.slot_ple_ptr, &g.slot_ple_idx)))
      return rc;
  }
  tick("cuboid edges");
  // ---- flags, per-edge outputs, scratch -----------------------------------------------------------------
#define DA(ptr, n) if ((rc = h->dalloc(&(ptr), (n)))) return rc
  DA(g.pe_flags, (size_t)g.n_pe); DA(g.ple_flags, (size_t)g.n_ple); DA(g.cbe_flags, (size_t)g.n_cbe); DA(g.pce_flags, (size_t)g.n_pce);
  DA(g.pe_chi2, (size_t)g.n_pe); DA(g.ple_chi2, (size_t)g.n_ple); DA(g.cbe_chi2, (size_t)g.n_cbe); DA(g.cbe_norm, (size_t)g.n_cbe); DA(g.pce_chi2, (size_t)g.n_pce);
  DA(g.ple_J, 27 * (size_t)g.n_ple); DA(g.cbe_J, 240 * (size_t)g.n_cbe); DA(g.cbe_err, 16 * (size_t)g.n_cbe); DA(g.cbe_w, (size_t)g.n_cbe); DA(g.pce_J, 27 * (size_t)g.n_pce);
  DA(g.ple_part, 54 * (size_t)g.n_ple); DA(g.cbe_part, 81 * (size_t)g.n_cbe); DA(g.pce_part, 54 * (size_t)g.n_pce);
  DA(g.kf_act, (size_t)g.n_kf); DA(g.cu_act, (size_t)g.n_cu); DA(g.pl_act, (size_t)g.n_pl); DA(g.pt_act, (size_t)g.n_pt);
  DA(g.kf_idx, (size_t)g.n_kf); DA(g.cu_off, (size_t)g.n_cu); DA(g.ent_pidx, (size_t)g.n_ent); DA(g.dims, 8);
  int n_free = 0;
  for (int i = 0; i < g.n_kf; i++) n_free += !kf_fixed[i];
  h->max_np = 6 * n_free + 9 * g.n_cu;
  h->ld = dense_num_blocks(h->max_np);  // Tm: tile columns of the (tiled, packed lower) reduced system, see ppo_dense.h
  DA(g.Hpp_kf, 36 * (size_t)g.n_kf); DA(g.Hpp_cu, 81 * (size_t)g.n_cu); DA(g.Hpc, 54 * (size_t)g.n_cbe); DA(g.bp, (size_t)h->max_np);
  DA(g.Hll, 6 * (size_t)g.n_lm); DA(g.bl, 3 * (size_t)g.n_lm); DA(g.Hpl, 18 * (size_t)g.n_ent); DA(g.BD, 18 * (size_t)g.n_ent); DA(g.Zent, 3 * (size_t)g.n_ent); DA(g.Dinv, 6 * (size_t)g.n_lm);
  DA(g.xl, 3 * (size_t)g.n_lm); DA(g.xp, dense_x_doubles(h->max_np));
  h->dist_solve = h->world > 1 && h->world <= DIST_MAX && !(std::getenv("PPO_DIST_SOLVE") && std::atoi(std::getenv("PPO_DIST_SOLVE")) == 0);
  if (h->dist_solve) {  // S | Winv | workspace in the peer-mapped arena
    if ((rc = dist_setup(h))) return rc;
  } else {
    DA(g.S, dense_matrix_doubles(h->max_np));
  }
  {  // linearisation: a grid-stride loop over the work units, sized to fill the device once
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
    h->nb_lin = std::max(1, std::min(cdiv(g.n_units, LIN_WARPS), sms * LIN_CTAS_PER_SM));
  }
  h->nb_res = cdiv(g.n_pe, RES_THREADS);
  h->nb_pl = cdiv(g.n_ple, SMALL_THREADS); h->nb_cb = cdiv(g.n_cbe, SMALL_THREADS); h->nb_pc = cdiv(g.n_pce, SMALL_THREADS);
  h->nb_bs = cdiv(g.n_pl, BS_WARPS) + g.n_units;  // partial sums of k_backsub (planes) + k_backsub_points
  DA(h->d_chi_pt, (size_t)std::max(h->nb_lin, h->nb_res)); DA(h->d_chi_pl, (size_t)h->nb_pl); DA(h->d_chi_cb, (size_t)h->nb_cb); DA(h->d_chi_pc, (size_t)h->nb_pc);
  DA(h->d_scale_part, (size_t)h->nb_bs);
  h->d_S_bak = nullptr;
  if (!h->dist_solve && h->P.solver == PPO_SOLVER_6_3) DA(h->d_S_bak, dense_matrix_doubles(h->max_np));
  if (!h->dist_solve) {
    DA(h->d_Winv, (size_t)dense_num_blocks(h->max_np) * DENSE_TILE);
    char *ws = nullptr;
    DA(ws, dense_workspace_bytes(h->max_np));
    h->d_dense_ws = ws;
    dense_workspace_init(ws, h->max_np, h->st);
  }
  DA(h->d_scal, 1); DA(h->d_not_spd, 1); DA(h->d_nout, 4); DA(h->d_red, 4); DA(h->d_lm, 1); DA(h->d_lm_in, 1);
  CK(cudaMemsetAsync(h->d_lm, 0, sizeof(LmDev), h->st));
  CK(cudaMemsetAsync(g.pe_chi2, 0, 8 * (size_t)g.n_pe, h->st));
  CK(cudaMemsetAsync(g.ple_chi2, 0, 8 * (size_t)g.n_ple, h->st));
  CK(cudaMemsetAsync(g.cbe_chi2, 0, 8 * (size_t)g.n_cbe, h->st));
  CK(cudaMemsetAsync(g.cbe_norm, 0, 8 * (size_t)g.n_cbe, h->st));
  CK(cudaMemsetAsync(g.pce_chi2, 0, 8 * (size_t)g.n_pce, h->st));
  CK(cudaMemsetAsync(g.Hpl, 0, 8 * 18 * (size_t)g.n_ent, h->st));
  CK(cudaMemsetAsync(g.BD, 0, 8 * 18 * (size_t)g.n_ent, h->st));
  CK(cudaMemsetAsync(g.Zent, 0, 8 * 3 * (size_t)g.n_ent, h->st));
  CK(cudaMemsetAsync(g.xl, 0, 8 * 3 * (size_t)g.n_lm, h->st));
  CK(cudaMemsetAsync(g.pe_flags, PPO_EF_ROBUST, (size_t)g.n_pe, h->st));
  CK(cudaMemsetAsync(g.ple_flags, PPO_EF_ROBUST, (size_t)g.n_ple, h->st));
  CK(cudaMemsetAsync(g.cbe_flags, PPO_EF_ROBUST, (size_t)g.n_cbe, h->st));
  CK(cudaMemsetAsync(g.pce_flags, 0, (size_t)g.n_pce, h->st));
  g.huber_mono = h->P.huber_mono; g.huber_stereo = h->P.huber_stereo; g.huber_plane = h->P.huber_plane; g.huber_vp = h->P.huber_vp_plane;
  g.huber_bbox = h->P.huber_bbox; g.huber_corner = h->P.huber_corner; g.huber_se3 = h->P.huber_se3;
  g.ptcu_ratio = h->P.ptcu_max_outside_margin_ratio; g.ptcu_prior = h->P.ptcu_prior_weight;
  tick("scratch alloc + memsets");
  CK(cudaStreamSynchronize(h->st));  // host vectors go out of scope
  tick("drain stream");
  CK(cudaGetLastError());
  {
    int dup = 0;
    CK(cudaMemcpy(&dup, h->d_dup, sizeof(int), cudaMemcpyDeviceToHost));
    if (dup) {  // MapPoint::mObservations is a map keyed by KeyFrame*: one observation per key-frame
      h->err = "a landmark is observed twice by the same key-frame";
      return PPO_E_INVALID;
    }
  }
  h->have_graph = true;
  return PPO_OK;
#undef UP
#undef DA
}

int ppo_ba_set_graph(ppo_ba_handle *h, const ppo_ba_graph *gi) {
  const int rc = set_graph_impl(h, gi);
  // arrays of the caller that are page-locked are read by the copy engine where they lie: no copy may be in flight when the call returns,
  // whatever the outcome (the successful path has drained the stream already)
  if (rc != PPO_OK && h && h->st) cudaStreamSynchronize(h->st);
  return rc;
}

int ppo_ba_host_register(void *ptr, size_t bytes) {
  if (!ptr || bytes == 0) return PPO_E_INVALID;
  if (cudaHostRegister(ptr, bytes, cudaHostRegisterPortable) != cudaSuccess) {  // (page-locked for every device of the process)
    cudaGetLastError();
    return PPO_E_CUDA;
  }
  return PPO_OK;
}
int ppo_ba_host_unregister(void *ptr) {
  if (!ptr) return PPO_E_INVALID;
  if (cudaHostUnregister(ptr) != cudaSuccess) {
    cudaGetLastError();
    return PPO_E_CUDA;
  }
  return PPO_OK;
}

int ppo_ba_reset(ppo_ba_handle *h) {
  if (!h || !h->have_graph) return PPO_E_INVALID;
  CK(cudaSetDevice(h->device));
  DevGraph &g = h->g;
  int rc;
  if ((rc = copy_state(h, h->sa, h->s0))) return rc;
  CK(cudaMemsetAsync(g.pe_flags, PPO_EF_ROBUST, (size_t)g.n_pe, h->st));
  CK(cudaMemsetAsync(g.ple_flags, PPO_EF_ROBUST, (size_t)g.n_ple, h->st));
  CK(cudaMemsetAsync(g.cbe_flags, PPO_EF_ROBUST, (size_t)g.n_cbe, h->st));
  CK(cudaMemsetAsync(g.pce_flags, 0, (size_t)g.n_pce, h->st));
  std::fill(h->cpe_flags.begin(), h->cpe_flags.end(), (uint8_t)PPO_EF_ROBUST);
  if (g.n_cpe) CK(cudaMemcpyAsync(h->d_cpe_flags, h->cpe_flags.data(), g.n_cpe, cudaMemcpyHostToDevice, h->st));
  CK(cudaStreamSynchronize(h->st));
  return PPO_OK;
}

static int allreduce(ppo_ba_handle *h, void *buf, size_t n, int dtype, int op) {
  if (h->world <= 1 || n == 0) return PPO_OK;
  int r = g_nccl.AllReduce(buf, buf, n, dtype, op, h->comm, h->st);
  if (r != 0) {
    h->err = std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error");
    return PPO_E_NCCL;
  }
  h->collectives++;
  return PPO_OK;
}

// ---- SparseOptimizer::initializeOptimization(0) ------------------------------------------------------
static int init_mapping(ppo_ba_handle *h) {
  DevGraph &g = h->g;
  cudaStream_t st = h->st;
  CK(cudaMemsetAsync(g.kf_act, 0, 4 * (size_t)g.n_kf, st));
  CK(cudaMemsetAsync(g.cu_act, 0, 4 * (size_t)g.n_cu, st));
  CK(cudaMemsetAsync(g.pl_act, 0, 4 * (size_t)g.n_pl, st));
  CK(cudaMemsetAsync(g.pt_act, 0, 4 * (size_t)g.n_pt, st));
  const int nmax = std::max(std::max(g.n_pe, g.n_ple), std::max(std::max(g.n_cbe, g.n_pce), std::max(g.n_slots, g.n_pt)));
  if (nmax > 0) {
    k_mark_active<<<cdiv(nmax, 256), 256, 0, st>>>(g);
    h->launches++;
  }
  if (g.n_cpe) {
    k_mark_active_cpe<<<cdiv(g.n_cpe, 256), 256, 0, st>>>(g, h->d_cpe_cuboid, h->d_cpe_plane, h->d_cpe_flags);
    h->launches++;
  }
  if (h->world > 1) {  // a key-frame / cuboid is active if ANY rank holds an active edge on it (landmarks -- points, planes -- are rank-local)
    int rc;
    if ((rc = allreduce(h, g.kf_act, g.n_kf, ncclInt32_, ncclMax_)) || (rc = allreduce(h, g.cu_act, g.n_cu, ncclInt32_, ncclMax_))) return rc;
  }
  k_build_index<<<1, 32, 0, st>>>(g);
  h->launches++;
  if (nmax > 0) {
    k_entry_pidx<<<cdiv(nmax, 256), 256, 0, st>>>(g);
    h->launches++;
  }
  if (g.n_ple) {
    k_slot_pidx<<<cdiv(g.n_ple, 256), 256, 0, st>>>(g);
    h->launches++;
  }
  CK(cudaMemcpyAsync(h->h_dims, g.dims, 8 * sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  h->n_kf_free = h->h_dims[0];
  h->n_p = h->h_dims[1];
  h->n_l = h->h_dims[2] + h->h_dims[3];
  h->n_active_edges = h->h_dims[4];
  for (size_t e = 0; e < h->cpe_flags.size(); e++) h->n_active_edges += !(h->cpe_flags[e] & PPO_EF_LEVEL1);
  return PPO_OK;
}

// ---- computeActiveErrors + buildSystem at the current estimates -------------------------------------------
// Enqueues only (capturable in a CUDA graph): the scalars {chi2, max diagonal} stay in d_scal for the LM controller kernel.
static int linearize(ppo_ba_handle *h, bool only_points_kernel = false) {
  DevGraph &g = h->g;
  cudaStream_t st = h->st;
  const DevState &s = h->sa;
  if (!only_points_kernel) {  // point landmarks accumulate into zeroed Hll / bl (one add per value); everything else is written by k_combine
    CK(cudaMemsetAsync(g.Hll + 6 * (size_t)g.n_pl, 0, 8 * 6 * (size_t)g.n_pt, st));
    CK(cudaMemsetAsync(g.bl + 3 * (size_t)g.n_pl, 0, 8 * 3 * (size_t)g.n_pt, st));
  }
  if (h->profiling) cudaEventRecord(h->evp[0], st);
  // the plane / cuboid / point-cuboid edges (numeric Jacobians, few 10^4 threads, latency-bound) run on three side streams next
  // to the point kernels; every edge writes its own record, the sides meet in k_combine
  const bool fork = !only_points_kernel && (g.n_ple || g.n_cbe || g.n_pce);
  const bool use_side[3] = {fork && g.n_ple > 0, fork && g.n_cbe > 0, fork && g.n_pce > 0};
  cudaStream_t s_pl = use_side[0] ? h->side[0] : st, s_cb = use_side[1] ? h->side[1] : st, s_pc = use_side[2] ? h->side[2] : st;
  if (fork) {
    CK(cudaEventRecord(h->ev_fork, st));
    for (int q = 0; q < 3; q++)
      if (use_side[q]) CK(cudaStreamWaitEvent(h->side[q], h->ev_fork, 0));
  }
  if (g.n_units) {
    k_point_linearize<<<h->nb_lin, LIN_WARPS * 32, 0, st>>>(g, s, h->d_chi_pt);
    h->launches++;
  }
  if (only_points_kernel) return PPO_OK;
  if (g.n_chunks) {
    k_pose_accumulate<<<g.n_chunks, POSE_THREADS, 0, st>>>(g, s, h->d_chunk_part);
    h->launches++;
  }
  if (g.n_ple) {
    k_plane_jac<<<cdiv(g.n_ple * 9, 128), 128, 0, s_pl>>>(g, s);
    k_plane_edges<true><<<h->nb_pl, SMALL_THREADS, 0, s_pl>>>(g, s, h->d_chi_pl);
    h->launches += 2;
  }
  if (g.n_cbe) {
    k_cuboid_jac<<<cdiv(g.n_cbe * 15, 128), 128, 0, s_cb>>>(g, s);
    k_cuboid_edges<true><<<h->nb_cb, SMALL_THREADS, 0, s_cb>>>(g, s, h->d_chi_cb);
    k_cuboid_assemble<<<cdiv(g.n_cbe, CBA_EDGES), CBA_EDGES * 15, 0, s_cb>>>(g);
    h->launches += 3;
  }
  if (g.n_pce) {
    k_ptcu_jac<<<cdiv(g.n_pce * 9, 128), 128, 0, s_pc>>>(g, s);
    k_ptcu_edges<true><<<h->nb_pc, SMALL_THREADS, 0, s_pc>>>(g, s, h->d_chi_pc);
    h->launches += 2;
  }
  for (int q = 0; q < 3; q++)
    if (use_side[q]) {
      CK(cudaEventRecord(h->ev_join[q], h->side[q]));
      CK(cudaStreamWaitEvent(st, h->ev_join[q], 0));
    }
  {  // fixed-order assembly of the key-frame / cuboid / plane blocks from the per-edge records
    const int n = g.n_kf * 27 + g.n_cu * 54 + g.n_pl * 9 + g.n_slots * 18;
    if (h->world > 1) {
      // sharded window: the pose side of this rank's landmark edges (points: chunk partials, planes: edge records) is first summed per
      // key-frame into a compact array, added up across the ranks and fed back as ONE chunk per key-frame
      int rc;
      k_chunk_reduce<<<cdiv(g.n_kf * 27, 128), 128, 0, st>>>(g, h->d_kf_chunk_ptr, h->d_chunk_part, h->d_kf_part);
      h->launches++;
      if ((rc = allreduce(h, h->d_kf_part, 27 * (size_t)g.n_kf, ncclFloat64_, ncclSum_))) return rc;
      k_combine<<<cdiv(n, 128), 128, 0, st>>>(g, h->d_kf_iota_ptr, h->d_kf_part, 0);
    } else {
      k_combine<<<cdiv(n, 128), 128, 0, st>>>(g, h->d_kf_chunk_ptr, h->d_chunk_part, 1);
    }
    h->launches++;
  }
  if (h->profiling) cudaEventRecord(h->evp[1], st);
  const bool own = h->owner();  // replicated (non-point) edges count once: on rank 0
  k_scalars<<<1, SCAL_THREADS, 0, st>>>(g, h->d_scal, h->d_chi_pt, g.n_units ? h->nb_lin : 0, h->d_chi_pl, g.n_ple ? h->nb_pl : 0, h->d_chi_cb,
                              (own && g.n_cbe) ? h->nb_cb : 0, h->d_chi_pc, (own && g.n_pce) ? h->nb_pc : 0, h->d_lm, 1, nullptr, 0, 0, nullptr,
                              h->d_red);
  h->launches++;
  CK(cudaMemsetAsync(&h->d_scal->max_diag, 0, sizeof(double), st));
  k_max_diag<<<std::max(1, std::min(148, cdiv(3 * g.n_lm, 2048))), 256, 0, st>>>(g, h->d_scal);
  h->launches++;
  if (h->world > 1) {
    int rc;
    if ((rc = allreduce(h, h->d_red, 2, ncclFloat64_, ncclSum_))) return rc;
    k_scalars_from_red<<<1, 32, 0, st>>>(h->d_scal, h->d_red, 0);
    k_set_red_maxdiag<<<1, 32, 0, st>>>(h->d_scal, h->d_red);
    if ((rc = allreduce(h, h->d_red + 2, 1, ncclFloat64_, ncclMax_))) return rc;
    k_scalars_from_red<<<1, 32, 0, st>>>(h->d_scal, h->d_red, 1);
  }
  return PPO_OK;
}
// linearisation + scalars on the host (debug / timing entry points)
static int linearize_sync(ppo_ba_handle *h) {
  int rc = linearize(h);
  if (rc) return rc;
  CK(cudaMemcpyAsync(h->h_scal, h->d_scal, sizeof(Scalars), cudaMemcpyDeviceToHost, h->st));
  CK(cudaStreamSynchronize(h->st));
  CK(cudaGetLastError());
  return PPO_OK;
}
// writes the LM inputs of a call / an explicit damping (debug entry points) into the device-side controller state
static int set_device_lambda(ppo_ba_handle *h, double lambda) {
  CK(cudaMemcpyAsync(&h->d_lm->lambda, &lambda, sizeof(double), cudaMemcpyHostToDevice, h->st));
  return PPO_OK;
}

// ---- residuals at the trial estimates ---------------------------------------------------------------------
static void residual_kernels(ppo_ba_handle *h, const DevState &s) {
  DevGraph &g = h->g;
  cudaStream_t st = h->st;
  // four independent kernels (they read the state and write their own per-edge chi2 / partial sums): the three small
  // ones run on the side streams next to the point residuals
  const bool use_side[3] = {g.n_ple > 0, g.n_cbe > 0, g.n_pce > 0};
  const bool fork = g.n_pe > 0 && (use_side[0] || use_side[1] || use_side[2]);
  if (fork) {
    cudaEventRecord(h->ev_fork, st);
    for (int q = 0; q < 3; q++)
      if (use_side[q]) cudaStreamWaitEvent(h->side[q], h->ev_fork, 0);
  }
  cudaStream_t s_pl = fork ? h->side[0] : st, s_cb = fork ? h->side[1] : st, s_pc = fork ? h->side[2] : st;
  if (g.n_pe) { k_point_residual<<<h->nb_res, RES_THREADS, 0, st>>>(g, s, h->d_chi_pt); h->launches++; }
  if (g.n_ple) { k_plane_edges<false><<<h->nb_pl, SMALL_THREADS, 0, s_pl>>>(g, s, h->d_chi_pl); h->launches++; }
  if (g.n_cbe) { k_cuboid_edges<false><<<h->nb_cb, SMALL_THREADS, 0, s_cb>>>(g, s, h->d_chi_cb); h->launches++; }
  if (g.n_pce) { k_ptcu_edges<false><<<h->nb_pc, SMALL_THREADS, 0, s_pc>>>(g, s, h->d_chi_pc); h->launches++; }
  if (fork)
    for (int q = 0; q < 3; q++)
      if (use_side[q]) {
        cudaEventRecord(h->ev_join[q], h->side[q]);
        cudaStreamWaitEvent(st, h->ev_join[q], 0);
      }
}

// ---- setLambda + Schur complement (+ optional factorisation / back-substitution) ----------------------------
static int schur_system(ppo_ba_handle *h) {
  DevGraph &g = h->g;
  cudaStream_t st = h->st;
  const int n_p = h->n_p, ld = h->ld;  // ld = Tm (tile columns of the allocation)
  const int Tc = dense_num_blocks(n_p), grow = 64 * Tc;  // row of the reduced gradient: first row of tile row Tc
  const size_t s_used = dense_tile_index(ld, Tc, Tc) * (size_t)DENSE_TILE;  // tile columns 0 .. Tc-1 are one contiguous range
  CK(cudaMemsetAsync(g.S, 0, 8 * s_used, st));
  CK(cudaMemsetAsync(h->d_not_spd, 0, sizeof(int), st));
  const int own = h->owner() ? 1 : 0;
  if (g.n_pl) { k_schur_bd<<<cdiv(g.n_pl, BD_WARPS), BD_WARPS * 32, 0, st>>>(g, h->d_lm, n_p, ld, 1, g.n_pl); h->launches++; }
  if (g.n_units) { k_schur_bd_points<<<g.n_units, 32, 0, st>>>(g, h->d_lm); h->launches++; }
  if (h->n_pairs) {
    const int n_warps = cdiv(h->n_pairs, PAIR_CHUNK);
    // <U = 2 records in flight per warp, 1 accumulator chain, 8 CTAs per SM (32 registers)>: the kernel gathers 0.6 GB (config 2) out of an
    // L2-resident array and wants warps, not per-warp parallelism -- measured 150 us against 222 (<8,1,1>: 96 registers), 155 (<8,2,4>) and
    // 154 (<4,1,6>); config 4: 1.90 ms against 3.05 / 2.04 / 2.02
    k_schur_pairs<2, 1, 8><<<cdiv(n_warps, PAIR_WARPS), PAIR_WARPS * 32, 0, st>>>(g, h->d_pair_keys, h->d_pair_vals, h->n_pairs, ld, grow, h->d_pair_bnd,
                                                                                     h->d_pair_bnd_key, h->d_pair_bnd_flag);
    k_schur_pairs_fix<<<cdiv(n_warps, PAIR_WARPS), PAIR_WARPS * 32, 0, st>>>(g, n_warps, ld, grow, h->d_pair_bnd, h->d_pair_bnd_key, h->d_pair_bnd_flag);
    h->launches += 2;
  }
  const int n_comp = g.n_kf * 36 + g.n_cu * 81 + g.n_cbe * 54 + n_p;
  if (n_comp && own) { k_compose<<<cdiv(n_comp, 256), 256, 0, st>>>(g, h->d_lm, n_p, ld, grow); h->launches++; }
  // single large window sharded over ranks: sum the partial reduced systems (Hschur | bschur) over NVLink
  if (h->world > 1) {
    if (h->dist_solve) {  // every column is summed by its owner straight from the other ranks' memory (ppo_dense.cu: k_dist_reduce)
      dense_dist_reduce(h->dist_peers, n_p, h->max_np, h->d_dense_ws, ++h->dist_seq, st, &h->launches);
      return PPO_OK;
    }
    return allreduce(h, g.S, s_used, ncclFloat64_, ncclSum_);  // packed lower triangle + gradient row only
  }
  return PPO_OK;
}
// factorisation + both substitutions of the reduced system in g.S (single GPU, or spread over the ranks of a sharded window)
static int enqueue_dense_solve(ppo_ba_handle *h) {
  DevGraph &g = h->g;
  if (h->dist_solve) {
    const int Tc = dense_num_blocks(h->n_p);
    if (Tc != h->dist_ops_tc) {  // this rank's operation queue for the current number of tile columns
      std::vector<unsigned> ops;
      dense_dist_build_ops(Tc, h->rank, h->world, &ops, h->dist_peers.blk, h->dist_peers.fwd);
      CK(cudaStreamSynchronize(h->st));
      if (h->d_dist_ops) cudaFree(h->d_dist_ops);
      h->d_dist_ops = nullptr;
      CK(cudaMalloc((void **)&h->d_dist_ops, std::max<size_t>(4, ops.size() * sizeof(unsigned))));
      if (!ops.empty()) CK(cudaMemcpy(h->d_dist_ops, ops.data(), ops.size() * sizeof(unsigned), cudaMemcpyHostToDevice));
      h->dist_n_ops = (int)ops.size(), h->dist_ops_tc = Tc;
    }
    dense_cholesky_solve_dist(h->dist_peers, h->n_p, h->max_np, g.xp, h->d_dense_ws, h->d_not_spd, h->d_dist_ops, h->dist_n_ops, h->dist_seq, h->st, &h->launches);
  } else {
    // LinearSolverEigen (solvers/linear_solver_eigen.h:94-124) also solves indefinite systems: keep the system, redo it by LDL^T if the Cholesky fails
    const bool eigen_flavour = h->P.solver == PPO_SOLVER_6_3 && h->d_S_bak != nullptr;
    if (eigen_flavour) CK(cudaMemcpyAsync(h->d_S_bak, g.S, 8 * dense_used_doubles(h->n_p, h->max_np), cudaMemcpyDeviceToDevice, h->st));
    dense_cholesky_solve(g.S, h->n_p, h->max_np, g.xp, h->d_Winv, h->d_dense_ws, h->d_not_spd, h->st, &h->launches, h->solve_sm_cap);
    if (eigen_flavour) dense_ldlt_fallback(h->d_S_bak, h->n_p, h->max_np, g.xp, h->d_not_spd, h->st, &h->launches);
  }
  return PPO_OK;
}
static int solve_and_backsub(ppo_ba_handle *h) {
  DevGraph &g = h->g;
  int rc_solve = enqueue_dense_solve(h);
  if (rc_solve) return rc_solve;
  // planes: one warp per landmark; points: one lane per 6x3 block (work units of the linearisation); partial sums of the
  // LM scale go to disjoint ranges of d_scale_part
  const int nbp = cdiv(g.n_pl, BS_WARPS);
  const bool fork = g.n_pl > 0 && g.n_units > 0;  // the (few, long) plane landmarks next to the points
  cudaStream_t s_pl = fork ? h->side[0] : h->st;
  if (fork) {
    CK(cudaEventRecord(h->ev_fork, h->st));
    CK(cudaStreamWaitEvent(s_pl, h->ev_fork, 0));
  }
  if (g.n_pl) { k_backsub<<<nbp, BS_WARPS * 32, 0, s_pl>>>(g, h->d_lm, h->d_scale_part, 1, g.n_pl); h->launches++; }
  if (g.n_units) { k_backsub_points<<<g.n_units, 32, 0, h->st>>>(g, h->d_lm, h->d_scale_part + nbp); h->launches++; }
  if (fork) {
    CK(cudaEventRecord(h->ev_join[0], s_pl));
    CK(cudaStreamWaitEvent(h->st, h->ev_join[0], 0));
  }
  return PPO_OK;
}

// ---- one damped trial: setLambda + Schur + solve + back-substitution + update + computeActiveErrors + scalars (enqueue only) ----
static int enqueue_trial(ppo_ba_handle *h) {
  DevGraph &g = h->g;
  cudaStream_t st = h->st;
  int rc;
  if (h->profiling) cudaEventRecord(h->evp[2], st);
  if ((rc = schur_system(h))) return rc;
  if (h->profiling) cudaEventRecord(h->evp[3], st);
  if ((rc = solve_and_backsub(h))) return rc;
  if (h->profiling) cudaEventRecord(h->evp[4], st);
  const int nv = g.n_kf + g.n_cu + g.n_pl + g.n_pt;
  k_update<<<cdiv(nv, 128), 128, 0, st>>>(g, h->sa, h->sb);
  h->launches++;
  residual_kernels(h, h->sb);
  const bool own = h->owner();
  k_scalars<<<1, SCAL_THREADS, 0, st>>>(g, h->d_scal, h->d_chi_pt, g.n_pe ? h->nb_res : 0, h->d_chi_pl, g.n_ple ? h->nb_pl : 0, h->d_chi_cb,
                              (own && g.n_cbe) ? h->nb_cb : 0, h->d_chi_pc, (own && g.n_pce) ? h->nb_pc : 0, h->d_lm, 1, h->d_scale_part,
                              g.n_lm ? h->nb_bs : 0, own ? h->n_p : 0, h->d_not_spd, h->d_red);
  h->launches++;
  if (h->world > 1) {  // {chi2, scale, "some rank met a non-positive pivot", "some rank saw the stop flag"}
    k_stop_to_red<<<1, 32, 0, st>>>(h->d_stop, h->d_red);
    if ((rc = allreduce(h, h->d_red, 4, ncclFloat64_, ncclSum_))) return rc;
    k_scalars_from_red<<<1, 32, 0, st>>>(h->d_scal, h->d_red, 2);
    k_stop_from_red<<<1, 32, 0, st>>>(h->d_stop, h->d_red);
  }
  if (h->profiling) cudaEventRecord(h->evp[5], st);
  return PPO_OK;
}
static void enqueue_decide(ppo_ba_handle *h, int graph, cudaGraphConditionalHandle h_trial, cudaGraphConditionalHandle h_iter) {
  DevGraph &g = h->g;
  k_lm_decide<<<1, 32, 0, h->st>>>(h->d_lm, h->d_scal, h->d_stop, graph, h_trial, h_iter);
  const size_t n = 3 * (size_t)g.n_pt + 19 * (size_t)g.n_kf;
  k_accept<<<(int)std::max<size_t>(1, std::min<size_t>(592, (n + 255) / 256)), 256, 0, h->st>>>(g, h->sa, h->sb, h->d_lm);
  h->launches += 2;
}

// The whole LM loop of one optimize() as ONE graph:  k_lm_begin -> WHILE(iteration){ linearise, k_lm_iter_begin, WHILE(trial){ trial,
// k_lm_decide, k_accept } }.  Bodies are stream-captured from the same enqueue functions the host-driven path uses.
static int build_lm_graph(ppo_ba_handle *h, ppo_ba_handle::LmGraph *out) {
  cudaStream_t st = h->st;
  cudaGraph_t G = nullptr, tmp = nullptr;
  CK(cudaGraphCreate(&G, 0));
  cudaGraphConditionalHandle h_iter, h_trial;
  CK(cudaGraphConditionalHandleCreate(&h_iter, G, 0, 0));
  auto add_while = [&](cudaGraph_t parent, cudaGraphConditionalHandle hd, cudaGraph_t *body) -> int {
    cudaStreamCaptureStatus cs;
    const cudaGraphNode_t *deps = nullptr;
    size_t ndeps = 0;
    CK(cudaStreamGetCaptureInfo(st, &cs, nullptr, nullptr, &deps, &ndeps));
    cudaGraphNodeParams p = {};
    p.type = cudaGraphNodeTypeConditional;
    p.conditional.handle = hd;
    p.conditional.type = cudaGraphCondTypeWhile;
    p.conditional.size = 1;
    cudaGraphNode_t node;
    CK(cudaGraphAddNode(&node, parent, deps, ndeps, &p));
    *body = p.conditional.phGraph_out[0];
    CK(cudaStreamUpdateCaptureDependencies(st, &node, 1, cudaStreamSetCaptureDependencies));
    return PPO_OK;
  };
  const long long l0 = h->launches;
  int rc;
  cudaGraph_t body_iter = nullptr, body_trial = nullptr;
  CK(cudaStreamBeginCaptureToGraph(st, G, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
  k_lm_begin<<<1, 32, 0, st>>>(h->d_lm, h->d_lm_in, 1, h_iter);
  rc = add_while(G, h_iter, &body_iter);
  cudaStreamEndCapture(st, &tmp);
  if (rc) return rc;
  CK(cudaGraphConditionalHandleCreate(&h_trial, body_iter, 0, 0));
  CK(cudaStreamBeginCaptureToGraph(st, body_iter, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
  rc = linearize(h);
  k_lm_iter_begin<<<1, 32, 0, st>>>(h->d_lm, h->d_scal, 1, h_trial);
  h->launches++;
  if (!rc) rc = add_while(body_iter, h_trial, &body_trial);
  cudaStreamEndCapture(st, &tmp);
  if (rc) return rc;
  const long long l1 = h->launches;
  CK(cudaStreamBeginCaptureToGraph(st, body_trial, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
  rc = enqueue_trial(h);
  enqueue_decide(h, 1, h_trial, h_iter);
  cudaStreamEndCapture(st, &tmp);
  if (rc) return rc;
  const long long l2 = h->launches;
  h->launches = l0;  // nothing ran yet: launches are counted per replay
  out->graph = G;
  out->nodes_iter = (int)(l1 - l0), out->nodes_trial = (int)(l2 - l1);
  static const bool timing = std::getenv("PPO_BA_TIMING") != nullptr;
  const auto ti0 = std::chrono::steady_clock::now();
  out->exec = nullptr;
  // same topology as an earlier loop of this handle: update its executable in place.  An executable of another topology (a window
  // with other edge kinds) stays in the list for a later window of its kind: it is only ever launched after a successful update,
  // which rewrites every node.
  for (size_t q = h->spare_execs.size(); q-- > 0 && !out->exec;) {
    cudaGraphExecUpdateResultInfo info;
    if (cudaGraphExecUpdate(h->spare_execs[q], G, &info) == cudaSuccess) {
      out->exec = h->spare_execs[q];
      h->spare_execs.erase(h->spare_execs.begin() + (long)q);
    } else {
      cudaGetLastError();
    }
  }
  const bool updated = out->exec != nullptr;
  if (!out->exec) CK(cudaGraphInstantiate(&out->exec, G, 0));
  if (timing)
    std::fprintf(stderr, "[build_lm_graph] %s %8.3f ms\n", updated ? "exec update" : "instantiate", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - ti0).count());
  return PPO_OK;
}

// ---- SparseOptimizer::optimize + OptimizationAlgorithmLevenberg::solve ------------------------------------------
int ppo_ba_optimize(ppo_ba_handle *h, int iters, const volatile unsigned char *stop, ppo_ba_stats *stats) {
  if (!h || !h->have_graph) return PPO_E_INVALID;
  CK(cudaSetDevice(h->device));
  if (stats) std::memset(stats, 0, sizeof *stats);
  cudaStream_t st = h->st;
  CK(cudaEventRecord(h->ev0, st));
  int rc = init_mapping(h);  // (host sync 1: the sizes of the reduced system decide the launch shapes)
  if (rc) return rc;
  h->host_syncs++;
  bool stop_at_entry = stop && *stop;
  {  // all ranks of a sharded window must take the same branch: emptiness of the UNION of the shards
    int n_l_all = h->n_l;
    if (h->world > 1) {  // ... and the stop flag at entry: all ranks run the same number of iterations
      int both[2] = {n_l_all, (stop && *stop) ? 1 : 0};
      int *d = reinterpret_cast<int *>(h->d_red + 2);
      CK(cudaMemcpyAsync(d, both, 2 * sizeof(int), cudaMemcpyHostToDevice, st));
      if ((rc = allreduce(h, d, 2, ncclInt32_, ncclSum_))) return rc;
      CK(cudaMemcpyAsync(both, d, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      n_l_all = both[0];
      stop_at_entry = both[1] > 0;
    }
    if (h->n_p + n_l_all == 0) return PPO_E_EMPTY;
  }
  const ppo_ba_params &P = h->P;
  LmIn &in = *h->h_lm_in;
  in.iters = stop_at_entry ? 0 : iters;  // SparseOptimizer::optimize tests terminate() at the loop head
  in.max_trials = P.lm_max_trials, in.tau = P.lm_tau, in.good_upper = P.lm_good_upper, in.good_lower = P.lm_good_lower;
  in.chi_const = cpe_chi_const(h);
  CK(cudaMemcpyAsync(h->d_lm_in, h->h_lm_in, sizeof(LmIn), cudaMemcpyHostToDevice, st));
  *h->h_stop = stop_at_entry ? 1 : 0;
  const bool graph = h->use_graph && h->world == 1 && !h->profiling;
  float ms;
  if (graph) {
    ppo_ba_handle::LmGraph *G = nullptr;
    for (auto &q : h->lm_graphs)
      if (q.n_p == h->n_p && q.n_l == h->n_l && q.sm_cap == h->solve_sm_cap) G = &q;
    if (!G) {
      ppo_ba_handle::LmGraph q;
      q.n_p = h->n_p, q.n_l = h->n_l, q.sm_cap = h->solve_sm_cap;
      static const bool timing = std::getenv("PPO_BA_TIMING") != nullptr;
      const auto tb0 = std::chrono::steady_clock::now();
      if ((rc = build_lm_graph(h, &q))) return rc;
      if (timing) std::fprintf(stderr, "[ppo_ba_optimize] build_lm_graph %8.3f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tb0).count());
      h->lm_graphs.push_back(q);
      G = &h->lm_graphs.back();
    }
    CK(cudaGraphLaunch(G->exec, st));
    CK(cudaMemcpyAsync(h->h_lm, h->d_lm, sizeof(LmDev), cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(h->ev1, st));
    if (stop) {  // forward the caller's flag to the device while the loop runs (polled by k_lm_decide, like g2o between trials)
      while (cudaEventQuery(h->ev1) == cudaErrorNotReady)
        if (*stop) *(volatile int *)h->h_stop = 1;
    }
    CK(cudaEventSynchronize(h->ev1));  // (host sync 2)
    h->host_syncs++;
    CK(cudaGetLastError());
    h->launches += 1 + (long long)h->h_lm->done * G->nodes_iter + (long long)h->h_lm->total_trials * G->nodes_trial;
  } else {
    // host-driven loop over the same kernels (profiling with per-phase events; sharded windows, whose collectives sit between
    // the kernels): the host only mirrors the two loop flags of the device-side controller
    k_lm_begin<<<1, 32, 0, st>>>(h->d_lm, h->d_lm_in, 0, 0);
    h->launches++;
    int *flags = reinterpret_cast<int *>(h->h_scal);  // pinned scratch: {trial_continue, iter_continue}
    bool iter_continue = in.iters > 0;
    while (iter_continue) {
      if ((rc = linearize(h))) return rc;
      k_lm_iter_begin<<<1, 32, 0, st>>>(h->d_lm, h->d_scal, 0, 0);
      h->launches++;
      bool first = true;
      bool trial_continue = true;
      while (trial_continue) {
        if (stop && *stop) *(volatile int *)h->h_stop = 1;
        if ((rc = enqueue_trial(h))) return rc;
        enqueue_decide(h, 0, 0, 0);
        CK(cudaMemcpyAsync(flags, &h->d_lm->trial_continue, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        h->host_syncs++;
        CK(cudaGetLastError());
        if (h->profiling && stats) {
          if (first) {
            cudaEventElapsedTime(&ms, h->evp[0], h->evp[1]);
            stats->ms_linearize += ms;
          }
          cudaEventElapsedTime(&ms, h->evp[2], h->evp[3]); stats->ms_schur += ms;
          cudaEventElapsedTime(&ms, h->evp[3], h->evp[4]); stats->ms_solve += ms;
          cudaEventElapsedTime(&ms, h->evp[4], h->evp[5]); stats->ms_update += ms;
        }
        first = false;
        trial_continue = flags[0] != 0;
        iter_continue = flags[1] != 0;
      }
    }
    CK(cudaMemcpyAsync(h->h_lm, h->d_lm, sizeof(LmDev), cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(h->ev1, st));
    CK(cudaEventSynchronize(h->ev1));
    h->host_syncs++;
  }
  if (stats) {
    const LmDev &lm = *h->h_lm;
    cudaEventElapsedTime(&ms, h->ev0, h->ev1);
    stats->ms_total = ms;
    stats->iterations = lm.done;
    stats->terminated = lm.term ? lm.term : ((lm.stop_seen || (stop && *stop)) ? 2 : 0);
    stats->n_pose_dim = h->n_p;
    stats->n_landmarks = h->n_l;
    stats->n_active_edges = h->n_active_edges;
    stats->total_trials = lm.total_trials;
    stats->chi2_initial = lm.chi2_initial;
    stats->chi2_final = lm.done ? lm.currentChi : 0.0;
    for (int i = 0; i < lm.done && i < PPO_TRACE_MAX; i++) stats->trace[i] = lm.trace[i];
  }
  return PPO_OK;
}

// ---- batch variants (run_batch above): one host thread per window drives that window's LM loop on its own stream --------
// The dense solve of one window is a dependency chain that cannot fill the device; a batch gives every window an equal share of the
// SMs for its persistent factorisation, so that the chains of the windows on one device run side by side.
static void batch_share_sms(ppo_ba_handle **h, int n) {
  int per_dev[64] = {};
  for (int i = 0; i < n; i++)
    if (h[i] && h[i]->device >= 0 && h[i]->device < 64) per_dev[h[i]->device]++;
  for (int i = 0; i < n; i++) {
    if (!h[i]) continue;
    const int k = (h[i]->device >= 0 && h[i]->device < 64) ? per_dev[h[i]->device] : 1;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h[i]->device);
    // (not every window is in its solve at the same moment: PPO_BATCH_SM_OVERSUB lets the shares add up to more than the device)
    static const double oversub = std::getenv("PPO_BATCH_SM_OVERSUB") ? std::atof(std::getenv("PPO_BATCH_SM_OVERSUB")) : 1.0;
    h[i]->solve_sm_cap = k > 1 ? std::max(4, std::min(sms, (int)(oversub * sms / k))) : 0;
  }
}
int ppo_ba_optimize_batch(ppo_ba_handle **h, int n, int iters, const volatile unsigned char *stop, ppo_ba_stats *stats) {
  if (!h || (n > 0 && !stats)) return PPO_E_INVALID;
  batch_share_sms(h, n);
  const int rc = run_batch(n, [=](int i) { return ppo_ba_optimize(h[i], iters, stop, &stats[i]); });
  for (int i = 0; i < n; i++)
    if (h[i]) h[i]->solve_sm_cap = 0;
  return rc;
}
int ppo_ba_local_ba_batch(ppo_ba_handle **h, int n, const volatile unsigned char *stop, ppo_ba_result *res) {
  if (!h || (n > 0 && !res)) return PPO_E_INVALID;
  batch_share_sms(h, n);
  const int rc = run_batch(n, [=](int i) { return ppo_ba_local_ba(h[i], stop, &res[i]); });
  for (int i = 0; i < n; i++)
    if (h[i]) h[i]->solve_sm_cap = 0;
  return rc;
}

int ppo_ba_recompute_edge_errors(ppo_ba_handle *h, int kind) {
  if (!h || !h->have_graph || kind != PPO_EDGE_POINT) return PPO_E_INVALID;
  CK(cudaSetDevice(h->device));
  if (h->g.n_pe) {
    k_point_error_level1<<<cdiv(h->g.n_pe, 256), 256, 0, h->st>>>(h->g, h->sa);
    h->launches++;
    CK(cudaGetLastError());
  }
  return PPO_OK;
}

int ppo_ba_edge_chi2(ppo_ba_handle *h, int kind, double *chi2, unsigned char *depth_positive, double *err_norm) {
  if (!h || !h->have_graph) return PPO_E_INVALID;
  CK(cudaSetDevice(h->device));
  DevGraph &g = h->g;
  const int n = ppo_ba_edge_count(h, kind);
  if (n < 0) return PPO_E_INVALID;
  if (n == 0) return PPO_OK;
  CK(cudaStreamSynchronize(h->st));
  const double *src = nullptr;
  switch (kind) {
    case PPO_EDGE_POINT: src = g.pe_chi2; break;
    case PPO_EDGE_PLANE: src = g.ple_chi2; break;
    case PPO_EDGE_CUBOID_CAM: src = g.cbe_chi2; break;
    case PPO_EDGE_POINT_CUBOID: src = g.pce_chi2; break;
    default: break;
  }
  if (kind == PPO_EDGE_CUBOID_PLANE) {
    if (chi2) std::copy(h->cpe_chi2.begin(), h->cpe_chi2.end(), chi2);
    if (err_norm) std::copy(h->cpe_norm.begin(), h->cpe_norm.end(), err_norm);
    if (depth_positive) std::memset(depth_positive, 1, n);
    return PPO_OK;
  }
  if (chi2) CK(cudaMemcpy(chi2, src, 8 * (size_t)n, cudaMemcpyDeviceToHost));
  if (err_norm) {
    if (kind == PPO_EDGE_CUBOID_CAM) CK(cudaMemcpy(err_norm, g.cbe_norm, 8 * (size_t)n, cudaMemcpyDeviceToHost));
    else std::fill(err_norm, err_norm + n, std::numeric_limits<double>::quiet_NaN());  // only defined for cuboid edges
  }
  if (depth_positive) {
    if (kind == PPO_EDGE_POINT || kind == PPO_EDGE_PLANE) {
      unsigned char *d = nullptr;
      CK(cudaMallocAsync((void **)&d, n, h->st));  // stream-ordered pool: no OS call in the steady state
      k_depth_flags<<<cdiv(n, 256), 256, 0, h->st>>>(g, h->sa, kind, d);
      h->launches++;
      cudaError_t e = cudaMemcpyAsync(depth_positive, d, n, cudaMemcpyDeviceToHost, h->st);
      cudaFreeAsync(d, h->st);
      cudaStreamSynchronize(h->st);
      CK(e);
    } else {
      std::memset(depth_positive, 1, n);
    }
  }
  return PPO_OK;
}

int ppo_ba_point_edge_outliers(ppo_ba_handle *h, double chi2_mono, double chi2_stereo, const int32_t **idx, int32_t *n) {
  if (!h || !h->have_graph || !idx || !n) return PPO_E_INVALID;
  CK(cudaSetDevice(h->device));
  DevGraph &g = h->g;
  *idx = nullptr, *n = 0;
  h->outlier_idx.clear();
  if (g.n_pe == 0) return PPO_OK;
  int *d_idx = nullptr, *d_cnt = nullptr;
  CK(cudaMallocAsync((void **)&d_idx, sizeof(int) * ((size_t)g.n_pe + 1), h->st));  // stream-ordered pool: no OS call in the steady state
  d_cnt = d_idx + g.n_pe;
  cudaError_t e = cudaMemsetAsync(d_cnt, 0, sizeof(int), h->st);
  k_point_edge_outliers<<<cdiv(g.n_pe, 256), 256, 0, h->st>>>(g, h->sa, chi2_mono, chi2_stereo, d_idx, d_cnt);
  h->launches++;
  int cnt = 0;
  if (e == cudaSuccess) e = cudaMemcpyAsync(&cnt, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, h->st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->st);
  if (e == cudaSuccess && cnt > 0) {
    h->outlier_idx.resize((size_t)cnt);
    e = cudaMemcpyAsync(h->outlier_idx.data(), d_idx, sizeof(int) * (size_t)cnt, cudaMemcpyDeviceToHost, h->st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->st);
  }
  cudaFreeAsync(d_idx, h->st);
  CK(e);
  CK(cudaGetLastError());
  std::sort(h->outlier_idx.begin(), h->outlier_idx.end());
  *idx = h->outlier_idx.data(), *n = (int32_t)h->outlier_idx.size();
  return PPO_OK;
}

static uint8_t *flags_ptr(ppo_ba_handle *h, int kind) {
  switch (kind) {
    case PPO_EDGE_POINT: return h->g.pe_flags;
    case PPO_EDGE_PLANE: return h->g.ple_flags;
    case PPO_EDGE_CUBOID_CAM: return h->g.cbe_flags;
    case PPO_EDGE_POINT_CUBOID: return h->g.pce_flags;
    case PPO_EDGE_CUBOID_PLANE: return h->d_cpe_flags;
  }
  return nullptr;
}
int ppo_ba_set_edge_flags(ppo_ba_handle *h, int kind, const unsigned char *flags) {
  if (!h || !h->have_graph || !flags) return PPO_E_INVALID;
  CK(cudaSetDevice(h->device));
  const int n = ppo_ba_edge_count(h, kind);
  if (n < 0) return PPO_E_INVALID;
  if (kind == PPO_EDGE_CUBOID_PLANE) h->cpe_flags.assign(flags, flags + n);
  if (n) CK(cudaMemcpy(flags_ptr(h, kind), flags, n, cudaMemcpyHostToDevice));
  return PPO_OK;
}
int ppo_ba_get_edge_flags(ppo_ba_handle *h, int kind, unsigned char *flags) {
  if (!h || !h->have_graph || !flags) return PPO_E_INVALID;
  CK(cudaSetDevice(h->device));
  const int n = ppo_ba_edge_count(h, kind);
  if (n < 0) return PPO_E_INVALID;
  CK(cudaStreamSynchronize(h->st));
  if (n) CK(cudaMemcpy(flags, flags_ptr(h, kind), n, cudaMemcpyDeviceToHost));
  return PPO_OK;
}

int ppo_ba_outlier_pass(ppo_ba_handle *h, int32_t n_out[3]) {
  if (!h || !h->have_graph) return PPO_E_INVALID;
  CK(cudaSetDevice(h->device));
  DevGraph &g = h->g;
  const ppo_ba_params &P = h->P;
  CK(cudaMemsetAsync(h->d_nout, 0, 4 * sizeof(int), h->st));
  const int nmax = std::max(g.n_pe, std::max(g.n_ple, g.n_cbe));
  if (nmax) {
    k_outlier_pass<<<cdiv(nmax, 256), 256, 0, h->st>>>(g, h->sa, P.chi2_mono, P.chi2_stereo, P.chi2_plane, P.chi2_vp_plane, P.norm_bbox,
                                                     P.norm_corner, P.norm_se3, h->d_nout);
    h->launches++;
  }
  int out[4];
  CK(cudaMemcpyAsync(out, h->d_nout, sizeof out, cudaMemcpyDeviceToHost, h->st));
  CK(cudaStreamSynchronize(h->st));
  if (n_out) n_out[0] = out[0], n_out[1] = out[1], n_out[2] = out[2];
  return PPO_OK;
}

int ppo_ba_local_ba(ppo_ba_handle *h, const volatile unsigned char *stop, ppo_ba_result *res) {
  if (!h || !res) return PPO_E_INVALID;
  std::memset(res, 0, sizeof *res);
  if (stop && *stop) {  // Optimizer.cc:2723-2725
    res->skipped = 1;
    return PPO_OK;
  }
  int rc = ppo_ba_optimize(h, h->P.iters_round1, stop, &res->round1);
  if (rc == PPO_E_EMPTY) rc = PPO_OK;  // g2o only logs "0 vertices to optimize" (core/sparse_optimizer.cpp:356-359); the caller's outlier pass, erase lists and write-back still run
  if (rc != PPO_OK) return rc;
  if (!(stop && *stop)) {
    int32_t n_out[3];
    if ((rc = ppo_ba_outlier_pass(h, n_out))) return rc;
    res->n_outlier_point_edges = n_out[0];
    res->n_outlier_plane_edges = n_out[1];
    res->n_outlier_cuboid_edges = n_out[2];
    rc = ppo_ba_optimize(h, h->P.iters_round2, stop, &res->round2);
    if (rc == PPO_E_EMPTY) rc = PPO_OK;
  }
  return rc;
}

int ppo_ba_get_state(ppo_ba_handle *h, ppo_ba_state *out) {
  if (!h || !h->have_graph || !out) return PPO_E_INVALID;
  CK(cudaSetDevice(h->device));
  DevGraph &g = h->g;
  const size_t nb[4] = {56 * (size_t)g.n_kf, 24 * (size_t)g.n_pt, 32 * (size_t)g.n_pl, 80 * (size_t)g.n_cu};
  const double *src[4] = {h->sa.kf_pose, h->sa.pt, h->sa.pl, h->sa.cu};
  double *dst[4] = {out->kf_pose, out->pt_xyz, out->pl_coef, out->cu_state};
  // ONE reservation for all requested members: growing the arena frees the old block, so pointers into it taken before a
  // second reservation would dangle (the uploads of set_graph are drained by now, so the arena is re-used from its start)
  CK(cudaStreamSynchronize(h->st));
  size_t total = 0, off[4] = {0, 0, 0, 0};
  for (int k = 0; k < 4; k++)
    if (dst[k] && nb[k]) {
      off[k] = total;
      total += (nb[k] + 255) & ~(size_t)255;
    }
  if (total == 0) return PPO_OK;
  h->hstage_off = 0;
  void *base = nullptr;
  int rc;
  if ((rc = h->pinned(&base, total))) return rc;
  for (int k = 0; k < 4; k++)
    if (dst[k] && nb[k]) CK(cudaMemcpyAsync((char *)base + off[k], src[k], nb[k], cudaMemcpyDeviceToHost, h->st));
  CK(cudaStreamSynchronize(h->st));
  for (int k = 0; k < 4; k++)
    if (dst[k] && nb[k]) std::memcpy(dst[k], (char *)base + off[k], nb[k]);
  h->hstage_off = 0;
  return PPO_OK;
}

int ppo_ba_set_profiling(ppo_ba_handle *h, int enable) {
  if (!h) return PPO_E_INVALID;
  h->profiling = enable != 0;
  return PPO_OK;
}
long long ppo_ba_launch_count(const ppo_ba_handle *h) { return h ? h->launches : 0; }
long long ppo_ba_host_sync_count(const ppo_ba_handle *h) { return h ? h->host_syncs : 0; }
int ppo_ba_set_graph_mode(ppo_ba_handle *h, int enable) {
  if (!h) return PPO_E_INVALID;
  h->use_graph = enable != 0;
  return PPO_OK;
}

static const size_t FLUSH_BYTES = 256ull << 20;
int ppo_ba_flush_l2(ppo_ba_handle *h) {
  if (!h) return PPO_E_INVALID;
  CK(cudaSetDevice(h->device));
  if (!h->d_flush) CK(cudaMalloc(&h->d_flush, FLUSH_BYTES));
  CK(cudaMemsetAsync(h->d_flush, 0, FLUSH_BYTES, h->st));
  return PPO_OK;
}
int ppo_ba_mark(ppo_ba_handle *h, int which) {
  if (!h || which < 0 || which > 1) return PPO_E_INVALID;
  CK(cudaSetDevice(h->device));
  CK(cudaEventRecord(h->evm[which], h->st));
  return PPO_OK;
}
int ppo_ba_elapsed_ms(ppo_ba_handle *h, double *ms) {
  if (!h || !ms) return PPO_E_INVALID;
  CK(cudaSetDevice(h->device));
  CK(cudaEventSynchronize(h->evm[1]));
  float f = 0;
  CK(cudaEventElapsedTime(&f, h->evm[0], h->evm[1]));
  *ms = f;
  return PPO_OK;
}

int ppo_ba_time_assembly(ppo_ba_handle *h, int reps, double *ms_mean, double *algo_bytes) {
  if (!h || !h->have_graph || reps <= 0) return PPO_E_INVALID;
  CK(cudaSetDevice(h->device));
  DevGraph &g = h->g;
  int rc = init_mapping(h);
  if (rc) return rc;
  // warm-up
  for (int i = 0; i < 3; i++) {
    CK(cudaMemsetAsync(g.Hll, 0, 8 * 6 * (size_t)g.n_lm, h->st));
    CK(cudaMemsetAsync(g.bl, 0, 8 * 3 * (size_t)g.n_lm, h->st));
    if ((rc = linearize(h, true))) return rc;
  }
  double total = 0;
  for (int i = 0; i < reps; i++) {
    CK(cudaMemsetAsync(g.Hll, 0, 8 * 6 * (size_t)g.n_lm, h->st));
    CK(cudaMemsetAsync(g.bl, 0, 8 * 3 * (size_t)g.n_lm, h->st));
    if ((rc = ppo_ba_flush_l2(h))) return rc;  // cold L2 for every timed launch
    CK(cudaEventRecord(h->ev0, h->st));
    k_point_linearize<<<h->nb_lin, LIN_WARPS * 32, 0, h->st>>>(g, h->sa, h->d_chi_pt);
    h->launches++;
    CK(cudaEventRecord(h->ev1, h->st));
    CK(cudaEventSynchronize(h->ev1));
    float ms;
    CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    total += ms;
  }
  if (ms_mean) *ms_mean = total / reps;
  // algorithmic bytes of one launch: per edge 16 B record + 4 B inv_sigma2 + 4 B point index
  // + 1 B flags read, 144 B Hpl + 8 B chi2 written; per point 4 B rowptr + 24 B xyz read, 72 B Hll/bl written;
  // per key-frame 96 B pose cache + 20 B intrinsics read.
  if (algo_bytes) *algo_bytes = (double)g.n_pe * (16 + 4 + 4 + 1 + 144 + 8) + (double)g.n_pt * (4 + 24 + 72) + (double)g.n_kf * (96 + 20);
  return PPO_OK;
}

// Times the dense solve of the reduced pose system (factorisation + both substitutions) on the current linearisation:
// the Schur system is rebuilt (untimed) before every timed solve because the factorisation is in place.
int ppo_ba_time_solve(ppo_ba_handle *h, int reps, double *ms_mean, double *flops, int *n_p_out) {
  if (!h || !h->have_graph || reps <= 0) return PPO_E_INVALID;
  CK(cudaSetDevice(h->device));
  int rc = init_mapping(h);
  if (rc) return rc;
  if ((rc = linearize_sync(h))) return rc;
  const double lambda = h->P.lm_tau * h->h_scal->max_diag;
  if ((rc = set_device_lambda(h, lambda))) return rc;
  double total = 0;
  for (int i = 0; i < reps + 2; i++) {
    if ((rc = schur_system(h))) return rc;
    CK(cudaEventRecord(h->ev0, h->st));
    if ((rc = enqueue_dense_solve(h))) return rc;
    CK(cudaEventRecord(h->ev1, h->st));
    CK(cudaEventSynchronize(h->ev1));
    float ms;
    CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    if (i >= 2) total += ms;
  }
  if (ms_mean) *ms_mean = total / reps;
  const double n = h->n_p;
  if (flops) *flops = n * n * n / 3.0 + 2.0 * n * n;  // Cholesky + two triangular solves
  if (n_p_out) *n_p_out = h->n_p;
  return PPO_OK;
}

// ---- parity / debugging exports (mirrored by the oracle's ppo_oracle_debug_*) ----------------------------------
int ppo_ba_debug_linearize(ppo_ba_handle *h, int32_t dims[2], double *Hpp, double *b, double *Hll, double *chi2) {
  if (!h || !h->have_graph) return PPO_E_INVALID;
  CK(cudaSetDevice(h->device));
  DevGraph &g = h->g;
  int rc = init_mapping(h);
  if (rc) return rc;
  {  // chi2 includes the constant cuboid-plane term: hand it to the device-side controller state
    const double cc = cpe_chi_const(h);
    CK(cudaMemcpyAsync(&h->d_lm->chi_const, &cc, sizeof(double), cudaMemcpyHostToDevice, h->st));
  }
  if ((rc = linearize_sync(h))) return rc;
  const int n_p = h->n_p;
  dims[0] = n_p;
  dims[1] = h->n_l;
  if (chi2) *chi2 = h->h_scal->chi2;
  if (!Hpp && !b && !Hll) return PPO_OK;
  std::vector<int> kf_idx(g.n_kf), cu_off(g.n_cu), pl_act(g.n_pl), pt_act(g.n_pt), cbe_kf(g.n_cbe), cbe_cu(g.n_cbe);
  std::vector<uint8_t> pt_fixed(g.n_pt), cbe_flags(g.n_cbe);
  std::vector<double> hkf(36 * (size_t)g.n_kf), hcu(81 * (size_t)g.n_cu), hpc(54 * (size_t)g.n_cbe), bp(h->max_np), hll(6 * (size_t)g.n_lm), bl(3 * (size_t)g.n_lm);
#define DL(dst, src, n) if ((n) > 0) CK(cudaMemcpy((dst).data(), (src), sizeof((dst)[0]) * (size_t)(n), cudaMemcpyDeviceToHost))
  DL(kf_idx, g.kf_idx, g.n_kf); DL(cu_off, g.cu_off, g.n_cu); DL(pl_act, g.pl_act, g.n_pl); DL(pt_act, g.pt_act, g.n_pt);
  DL(cbe_kf, g.cbe_kf, g.n_cbe); DL(cbe_cu, g.cbe_cuboid, g.n_cbe); DL(pt_fixed, g.pt_fixed, g.n_pt); DL(cbe_flags, g.cbe_flags, g.n_cbe);
  DL(hkf, g.Hpp_kf, 36 * (size_t)g.n_kf); DL(hcu, g.Hpp_cu, 81 * (size_t)g.n_cu); DL(hpc, g.Hpc, 54 * (size_t)g.n_cbe); DL(bp, g.bp, h->max_np);
  DL(hll, g.Hll, 6 * (size_t)g.n_lm); DL(bl, g.bl, 3 * (size_t)g.n_lm);
  if (Hpp) {
    std::fill(Hpp, Hpp + (size_t)n_p * n_p, 0.0);
    for (int k = 0; k < g.n_kf; k++)
      if (kf_idx[k] >= 0)
        for (int a = 0; a < 6; a++)
          for (int c = 0; c < 6; c++) Hpp[(size_t)(6 * kf_idx[k] + a) * n_p + 6 * kf_idx[k] + c] = hkf[36 * (size_t)kf_idx[k] + 6 * a + c];
    for (int k = 0; k < g.n_cu; k++)
      if (cu_off[k] >= 0)
        for (int a = 0; a < 9; a++)
          for (int c = 0; c < 9; c++) Hpp[(size_t)(cu_off[k] + a) * n_p + cu_off[k] + c] = hcu[81 * (size_t)k + 9 * a + c];
    for (int e = 0; e < g.n_cbe; e++) {
      if (cbe_flags[e] & PPO_EF_LEVEL1) continue;
      const int idx = kf_idx[cbe_kf[e]], off = cu_off[cbe_cu[e]];
      if (idx < 0 || off < 0) continue;
      for (int a = 0; a < 6; a++)
        for (int c = 0; c < 9; c++) Hpp[(size_t)(6 * idx + a) * n_p + off + c] += hpc[54 * (size_t)e + 9 * a + c];
    }
  }
  int l = 0;
  for (int L = 0; L < g.n_lm; L++) {
    const bool act = L < g.n_pl ? pl_act[L] != 0 : (pt_act[L - g.n_pl] != 0 && !pt_fixed[L - g.n_pl]);
    if (!act) continue;
    if (Hll) {
      const double *s = &hll[6 * (size_t)L];
      double *d = &Hll[9 * (size_t)l];
      d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; d[3] = s[1]; d[4] = s[3]; d[5] = s[4]; d[6] = s[2]; d[7] = s[4]; d[8] = s[5];
    }
    if (b) for (int i = 0; i < 3; i++) b[n_p + 3 * (size_t)l + i] = bl[3 * (size_t)L + i];
    l++;
  }
  if (b) for (int i = 0; i < n_p; i++) b[i] = bp[i];
  return PPO_OK;
}

int ppo_ba_debug_solve(ppo_ba_handle *h, double lambda, double *Hschur_upper, double *bschur, double *x, int32_t *ok) {
  if (!h || !h->have_graph) return PPO_E_INVALID;
  CK(cudaSetDevice(h->device));
  DevGraph &g = h->g;
  const int n_p = h->n_p, ld = h->ld;
  int rc;
  if ((rc = set_device_lambda(h, lambda))) return rc;
  if ((rc = schur_system(h))) return rc;
  CK(cudaStreamSynchronize(h->st));
  if (Hschur_upper || bschur) {
    const int Tc = dense_num_blocks(n_p), grow = 64 * Tc;
    const size_t s_used = dense_tile_index(ld, Tc, Tc) * (size_t)DENSE_TILE;
    std::vector<double> S(s_used + 1);
    if (n_p) CK(cudaMemcpy(S.data(), g.S, 8 * s_used, cudaMemcpyDeviceToHost));
    for (int i = 0; i < n_p; i++) {  // upper (i, j) of the symmetric matrix = lower (j, i) of the tiled storage
      if (Hschur_upper)
        for (int j = 0; j < n_p; j++) Hschur_upper[(size_t)i * n_p + j] = j >= i ? S[dense_elem_index(ld, j, i)] : 0.0;
      if (bschur) bschur[i] = S[dense_elem_index(ld, grow, i)];
    }
  }
  if ((rc = solve_and_backsub(h))) return rc;
  int ns = 0;
  CK(cudaMemcpyAsync(&ns, h->d_not_spd, sizeof(int), cudaMemcpyDeviceToHost, h->st));
  CK(cudaStreamSynchronize(h->st));
  CK(cudaGetLastError());
  if (ok) *ok = !ns;
  if (x) {
    std::vector<double> xl(3 * (size_t)g.n_lm);
    std::vector<int> pl_act(g.n_pl), pt_act(g.n_pt);
    std::vector<uint8_t> pt_fixed(g.n_pt);
    if (n_p) CK(cudaMemcpy(x, g.xp, 8 * (size_t)n_p, cudaMemcpyDeviceToHost));
    DL(xl, g.xl, 3 * (size_t)g.n_lm); DL(pl_act, g.pl_act, g.n_pl); DL(pt_act, g.pt_act, g.n_pt); DL(pt_fixed, g.pt_fixed, g.n_pt);
    int l = 0;
    for (int L = 0; L < g.n_lm; L++) {
      const bool act = L < g.n_pl ? pl_act[L] != 0 : (pt_act[L - g.n_pl] != 0 && !pt_fixed[L - g.n_pl]);
      if (!act) continue;
      for (int i = 0; i < 3; i++) x[n_p + 3 * (size_t)l + i] = xl[3 * (size_t)L + i];
      l++;
    }
  }
#undef DL
  return PPO_OK;
}

int ppo_ba_debug_dense_solve(ppo_ba_handle *h, int32_t n, const double *A_upper, const double *b, double *x, int32_t *ok) {
  if (!h || n <= 0 || !A_upper || !b || !x) return PPO_E_INVALID;
  CK(cudaSetDevice(h->device));
  const int Tm = dense_num_blocks(n), grow = 64 * Tm;
  const size_t nS = dense_matrix_doubles(n);
  std::vector<double> S(nS, 0.0);
  for (int j = 0; j < n; j++) {  // lower (i, j) of the tiled storage = upper (j, i) of the caller's matrix
    for (int i = j; i < n; i++) S[dense_elem_index(Tm, i, j)] = A_upper[(size_t)j * n + i];
    S[dense_elem_index(Tm, grow, j)] = b[j];
  }
  double *dS = nullptr, *dBak = nullptr, *dx = nullptr, *dW = nullptr;
  void *ws = nullptr;
  int *dns = nullptr;
  auto release = [&] { cudaFree(dS), cudaFree(dBak), cudaFree(dx), cudaFree(dW), cudaFree(ws), cudaFree(dns); };
  const bool eigen_flavour = h->P.solver == PPO_SOLVER_6_3;
  cudaError_t e = cudaMalloc((void **)&dS, nS * 8);
  if (e == cudaSuccess && eigen_flavour) e = cudaMalloc((void **)&dBak, nS * 8);
  if (e == cudaSuccess) e = cudaMalloc((void **)&dx, dense_x_doubles(n) * 8);
  if (e == cudaSuccess) e = cudaMalloc((void **)&dW, (size_t)Tm * DENSE_TILE * 8);
  if (e == cudaSuccess) e = cudaMalloc(&ws, dense_workspace_bytes(n));
  if (e == cudaSuccess) e = cudaMalloc((void **)&dns, sizeof(int));
  if (e == cudaSuccess) e = cudaMemsetAsync(dns, 0, sizeof(int), h->st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dS, S.data(), nS * 8, cudaMemcpyHostToDevice, h->st);
  if (e == cudaSuccess && eigen_flavour) e = cudaMemcpyAsync(dBak, dS, 8 * dense_used_doubles(n, n), cudaMemcpyDeviceToDevice, h->st);
  if (e != cudaSuccess) {
    release();
    CK(e);
  }
  dense_workspace_init(ws, n, h->st);
  dense_cholesky_solve(dS, n, n, dx, dW, ws, dns, h->st, &h->launches, h->solve_sm_cap);
  if (eigen_flavour) dense_ldlt_fallback(dBak, n, n, dx, dns, h->st, &h->launches);
  int ns = 0;
  e = cudaMemcpyAsync(&ns, dns, sizeof(int), cudaMemcpyDeviceToHost, h->st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(x, dx, 8 * (size_t)n, cudaMemcpyDeviceToHost, h->st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->st);
  if (e == cudaSuccess) e = cudaGetLastError();
  release();
  CK(e);
  if (ok) *ok = !ns;
  return PPO_OK;
}

int ppo_ba_nccl_unique_id(char out[128]) {
  g_nccl.load();
  if (!g_nccl.ok || !out) return PPO_E_NCCL;
  ncclUniqueId_ id;
  if (g_nccl.GetUniqueId(&id) != 0) return PPO_E_NCCL;
  std::memcpy(out, id.internal, 128);
  return PPO_OK;
}
int ppo_ba_nccl_init(const char id_bytes[128], int rank, int world, int device, void **comm) {
  g_nccl.load();
  if (!g_nccl.ok || !id_bytes || !comm) return PPO_E_NCCL;
  if (cudaSetDevice(device) != cudaSuccess) return PPO_E_CUDA;
  ncclUniqueId_ id;
  std::memcpy(id.internal, id_bytes, 128);
  ncclComm_t c = nullptr;
  if (g_nccl.CommInitRank(&c, world, id, rank) != 0) return PPO_E_NCCL;
  *comm = c;
  return PPO_OK;
}
int ppo_ba_nccl_destroy(void *comm) {
  g_nccl.load();
  if (!g_nccl.ok) return PPO_E_NCCL;
  return comm ? (g_nccl.CommDestroy((ncclComm_t)comm) == 0 ? PPO_OK : PPO_E_NCCL) : PPO_OK;
}
long long ppo_ba_collective_count(const ppo_ba_handle *h) { return h ? h->collectives : 0; }

int ppo_ba_set_shard(ppo_ba_handle *h, void *nccl_comm, int rank, int world) {
  if (!h || world < 1 || rank < 0 || rank >= world) return PPO_E_INVALID;
  if (world > 1) {
    g_nccl.load();
    if (!g_nccl.ok || !nccl_comm) {
      h->err = "NCCL not available (libnccl.so.2 could not be opened) or null communicator";
      return PPO_E_NCCL;
    }
  }
  h->comm = (ncclComm_t)nccl_comm;
  h->rank = rank;
  h->world = world;
  return PPO_OK;
}

}  // extern "C"
