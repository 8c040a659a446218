// Drop-in replacement for the two local-BA entry points of src/Optimizer.cc:
//   void Optimizer::LocalBundleAdjustment(KeyFrame*, bool*, Map*)                         (:461-786)
//   void Optimizer::LocalBACameraPlaneCuboids(KeyFrame*, bool*, Map*, bool, bool)         (:1994-2967)
//   void Optimizer::LocalBACameraPointCuboids2D(KeyFrame*, bool*, Map*, bool, bool)       (:1252-1992)
// Stage A (window collection), stage B (graph flattening instead of g2o vertex/edge construction), stage F
// (erase lists) and stage G (write-back) stay on the host and follow the reference statement by statement;
// stages C-E (optimize(5), outlier pass, optimize(10)) run on the GPU through the C-ABI (include/ppo_ba.h).
// Compiled against the real ORB-SLAM2 headers with -DPPO_WITH_ORB_SLAM2, otherwise against ppo_mock_slam.h.
#ifdef PPO_WITH_ORB_SLAM2
#include "Converter.h"
#include "KeyFrame.h"
#include "Map.h"
#include "MapCuboid.h"
#include "MapPlane.h"
#include "MapPoint.h"
#include "Optimizer.h"
#include "Parameters.h"
#else
#include "ppo_mock_slam.h"
#endif

#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <cstdio>
#include <map>
#include <mutex>
#include <thread>
#include <functional>
#include <condition_variable>
#include <vector>

#include "../../../include/ppo_ba.h"
#include "ppo_convert.h"

#ifdef PPO_SHIM_ON_ORACLE
// TEST BUILD ONLY (tests/shim_lib.py): the same shim source bound to the CPU oracle, so that the host logic (flattening,
// re-levelling loops, write-back) can be exercised end to end without a GPU.  Never part of the product libraries.
extern "C" {
struct ppo_oracle_handle;
void ppo_oracle_default_params(ppo_ba_params *);
int ppo_oracle_create(const ppo_ba_params *, ppo_oracle_handle **);
void ppo_oracle_destroy(ppo_oracle_handle *);
int ppo_oracle_set_graph(ppo_oracle_handle *, const ppo_ba_graph *);
int ppo_oracle_reset(ppo_oracle_handle *);
int ppo_oracle_optimize(ppo_oracle_handle *, int, const volatile unsigned char *, ppo_ba_stats *);
int ppo_oracle_edge_chi2(ppo_oracle_handle *, int, double *, unsigned char *, double *);
int ppo_oracle_recompute_edge_errors(ppo_oracle_handle *, int);
int ppo_oracle_set_edge_flags(ppo_oracle_handle *, int, const unsigned char *);
int ppo_oracle_local_ba(ppo_oracle_handle *, const volatile unsigned char *, ppo_ba_result *);
int ppo_oracle_get_state(ppo_oracle_handle *, ppo_ba_state *);
int ppo_oracle_set_params(ppo_oracle_handle *, const ppo_ba_params *);
}
#define ppo_ba_handle ppo_oracle_handle
#define ppo_ba_default_params ppo_oracle_default_params
#define ppo_ba_create(P, dev, out) ppo_oracle_create((P), (out))
#define ppo_ba_destroy ppo_oracle_destroy
#define ppo_ba_set_params ppo_oracle_set_params
#define ppo_ba_set_graph ppo_oracle_set_graph
#define ppo_ba_reset ppo_oracle_reset
#define ppo_ba_optimize ppo_oracle_optimize
#define ppo_ba_edge_chi2 ppo_oracle_edge_chi2
#define ppo_ba_recompute_edge_errors ppo_oracle_recompute_edge_errors
#define ppo_ba_set_edge_flags ppo_oracle_set_edge_flags
#define ppo_ba_local_ba ppo_oracle_local_ba
#define ppo_ba_get_state ppo_oracle_get_state
#define ppo_ba_host_register(p, n) (-1)  /* (no device in the test build: nothing to page-lock) */
#define ppo_ba_host_unregister(p) (0)
#define ppo_ba_last_error(h) "oracle backend"
#endif

namespace ppo_shim {

// ---- adapters between the reference's value types and flat doubles ------------------------------------
#ifdef PPO_WITH_ORB_SLAM2
inline void cuboid_to10(const g2o::cuboid &c, double o[10]) {
  auto v = c.toVector();  // [t(3) q(xyzw) scale(3)], g2o_cuboid.h:166-172
  for (int i = 0; i < 10; i++) o[i] = v(i);
}
inline void cuboid_from10(const double v[10], g2o::cuboid &c) {
  Eigen::Matrix<double, 10, 1> e;
  for (int i = 0; i < 10; i++) e(i) = v[i];
  c.fromVector(e);
}
#else
inline void cuboid_to10(const g2o::cuboid &c, double o[10]) {
  o[0] = c.pose7[4]; o[1] = c.pose7[5]; o[2] = c.pose7[6];
  o[3] = c.pose7[0]; o[4] = c.pose7[1]; o[5] = c.pose7[2]; o[6] = c.pose7[3];
  o[7] = c.scale[0]; o[8] = c.scale[1]; o[9] = c.scale[2];
}
inline void cuboid_from10(const double v[10], g2o::cuboid &c) {
  c.pose7[4] = v[0]; c.pose7[5] = v[1]; c.pose7[6] = v[2];
  c.pose7[0] = v[3]; c.pose7[1] = v[4]; c.pose7[2] = v[5]; c.pose7[3] = v[6];
  c.scale[0] = v[7]; c.scale[1] = v[8]; c.scale[2] = v[9];
}
#endif
inline void mat_to_float16(const cv::Mat &T, float o[16]) {
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) o[4 * r + c] = T.at<float>(r, c);
}
inline cv::Mat float16_to_mat(const float T[16]) {
  cv::Mat m(4, 4, CV_32F);
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) m.at<float>(r, c) = T[4 * r + c];
  return m;
}

// ---- the flat graph under construction ----------------------------------------------------------------------
struct Flat {
  std::vector<double> kf_pose, pt_xyz, pl_coef, cu_state, ple_meas, ple_info, cbe_meas, cbe_info, pce_pts, cpe_meas, cpe_info;
  std::vector<uint8_t> kf_fixed, pt_fixed, cu_flags, ple_kind, cbe_kind;
  std::vector<float> kf_intr, pe_obs, pe_invsigma2;
  std::vector<int32_t> pt_rowptr, pe_kf, ple_plane, ple_kf, cbe_kf, cbe_cuboid, pce_cuboid, pce_rowptr, cpe_cuboid, cpe_plane;
  ppo_ba_graph g;
  // The arrays as long as the window's points / point edges are kept page-locked (ppo_ba_host_register), so that ppo_ba_set_graph copies
  // them to the device from where they lie instead of staging them.  A vector is released BEFORE it may reallocate and registered again
  // at its new place; in the steady state (capacity reached) neither happens.
  struct Pin {
    void *p = nullptr;
    size_t bytes = 0;
  };
  Pin pins[5];
  template <class V>
  void resize_pinned(V &v, size_t n, Pin &pin) {
    if (n > v.capacity()) {
      if (pin.p) ppo_ba_host_unregister(pin.p), pin = Pin();
      v.reserve(n + n / 4);
    }
    v.resize(n);
    const size_t bytes = v.capacity() * sizeof(typename V::value_type);
    if ((void *)v.data() != pin.p || bytes != pin.bytes) {
      if (pin.p) ppo_ba_host_unregister(pin.p), pin = Pin();
      if (bytes >= 65536 && ppo_ba_host_register((void *)v.data(), bytes) == PPO_OK) pin.p = (void *)v.data(), pin.bytes = bytes;
    }
  }
  // empties the arrays but keeps their memory: a steady-state call then touches no fresh pages (30 MB for a 200-key-frame window)
  // keep_point_arrays: the point / point-edge arrays are then RESIZED and overwritten by the caller (run()), so they keep their size too --
  // a resize to nearly the same length value-initialises nothing, where clear() + resize() would zero 10 MB per call
  void clear(bool keep_point_arrays = false) {
    for (auto *v : {&kf_pose, &pl_coef, &cu_state, &ple_meas, &ple_info, &cbe_meas, &cbe_info, &pce_pts, &cpe_meas, &cpe_info}) v->clear();
    for (auto *v : {&kf_fixed, &cu_flags, &ple_kind, &cbe_kind}) v->clear();
    kf_intr.clear();
    for (auto *v : {&ple_plane, &ple_kf, &cbe_kf, &cbe_cuboid, &pce_cuboid, &pce_rowptr, &cpe_cuboid, &cpe_plane}) v->clear();
    if (!keep_point_arrays) {
      pt_xyz.clear(), pt_fixed.clear(), pt_rowptr.clear();
      pe_kf.clear(), pe_obs.clear(), pe_invsigma2.clear();
    }
  }
  void publish() {
    std::memset(&g, 0, sizeof g);
    g.n_kf = (int32_t)kf_fixed.size(); g.kf_pose = kf_pose.data(); g.kf_fixed = kf_fixed.data(); g.kf_intr = kf_intr.data();
    g.n_pt = (int32_t)(pt_xyz.size() / 3); g.pt_xyz = pt_xyz.data(); g.pt_fixed = pt_fixed.data();
    g.n_pl = (int32_t)(pl_coef.size() / 4); g.pl_coef = pl_coef.data();
    g.n_cu = (int32_t)(cu_state.size() / 10); g.cu_state = cu_state.data(); g.cu_flags = cu_flags.data();
    if (pt_rowptr.empty()) pt_rowptr.push_back(0);
    g.pt_rowptr = pt_rowptr.data(); g.n_pe = (int32_t)pe_kf.size(); g.pe_kf = pe_kf.data(); g.pe_obs = pe_obs.data(); g.pe_invsigma2 = pe_invsigma2.data();
    g.n_ple = (int32_t)ple_plane.size(); g.ple_plane = ple_plane.data(); g.ple_kf = ple_kf.data(); g.ple_kind = ple_kind.data();
    g.ple_meas = ple_meas.data(); g.ple_info = ple_info.data();
    g.n_cbe = (int32_t)cbe_kf.size(); g.cbe_kf = cbe_kf.data(); g.cbe_cuboid = cbe_cuboid.data(); g.cbe_kind = cbe_kind.data();
    g.cbe_meas = cbe_meas.data(); g.cbe_info = cbe_info.data();
    if (pce_rowptr.empty()) pce_rowptr.push_back(0);
    g.n_pce = (int32_t)pce_cuboid.size(); g.pce_cuboid = pce_cuboid.data(); g.pce_rowptr = pce_rowptr.data(); g.pce_pts = pce_pts.data();
    g.n_cpe = (int32_t)cpe_cuboid.size(); g.cpe_cuboid = cpe_cuboid.data(); g.cpe_plane = cpe_plane.data(); g.cpe_meas = cpe_meas.data();
    g.cpe_info = cpe_info.data();
  }
};

// One engine handle, one flattened window and one mutex PER ENTRY POINT: in ORB-SLAM2 Tracking calls PoseOptimization on every
// frame while LocalMapping runs the local BA and LoopClosing may run the global BA; with a single shared handle
// Tracking would block for a whole local BA.  The handle of a slot (streams, events, pinned staging, the device memory pool) is
// kept across calls; the reference's tuning globals are re-snapshotted on every call and, when they changed, set on the existing
// handle (ppo_ba_set_params) instead of re-creating it.
struct Slot {
  std::mutex m;
  ppo_ba_handle *h = nullptr;
  ppo_ba_params P;
  int device = -1;
  Flat last;  // last flattened window (introspection for tests / logging)
  // per-call work arrays of run() that are as long as the window's points / point edges: kept across calls like the flat arrays
  std::vector<std::pair<ORB_SLAM2::KeyFrame *, ORB_SLAM2::MapPoint *>> point_edge_owner;
  std::vector<ORB_SLAM2::MapPoint *> graph_points;
  std::vector<uint32_t> pt_first, e_first;
  ppo_ba_result res;
  int rc = 0;
};
enum { SLOT_LOCAL = 0, SLOT_GLOBAL = 1, SLOT_POSE = 2 };
static Slot g_slots[3];
static std::atomic<Slot *> g_last_slot{&g_slots[SLOT_LOCAL]};
static int g_device = 0;
static ppo_ba_handle *engine(Slot &S, const ppo_ba_params &P) {
  if (S.h && S.device == g_device) {
    if (std::memcmp(&S.P, &P, sizeof P) != 0) {
      if (ppo_ba_set_params(S.h, &P) != PPO_OK) return nullptr;
      S.P = P;
    }
    return S.h;
  }
  if (S.h) ppo_ba_destroy(S.h), S.h = nullptr;
  if (ppo_ba_create(&P, g_device, &S.h) != PPO_OK) S.h = nullptr;
  S.P = P;
  S.device = g_device;
  return S.h;
}

using namespace ORB_SLAM2;
// indices (ascending) of the point edges the local BA erases: chi2 above the threshold of the edge's kind, or a non-positive depth
#ifdef PPO_SHIM_ON_ORACLE
static int point_edge_outliers(ppo_ba_handle *h, const Flat &F, double th_mono, double th_stereo, const int32_t **idx, int32_t *n) {
  static std::vector<int32_t> out;  // (test build: the oracle has no such call; same test on its per-edge outputs)
  std::vector<double> chi2(F.g.n_pe);
  std::vector<unsigned char> dpos(F.g.n_pe);
  const int rc = ppo_ba_edge_chi2(h, PPO_EDGE_POINT, chi2.data(), dpos.data(), nullptr);
  out.clear();
  for (int e = 0; e < F.g.n_pe; e++)
    if (chi2[e] > (F.pe_obs[3 * (size_t)e + 2] < 0 ? th_mono : th_stereo) || !dpos[e]) out.push_back(e);
  *idx = out.data(), *n = (int32_t)out.size();
  return rc;
}
#else
static int point_edge_outliers(ppo_ba_handle *h, const Flat &, double th_mono, double th_stereo, const int32_t **idx, int32_t *n) {
  return ppo_ba_point_edge_outliers(h, th_mono, th_stereo, idx, n);
}
#endif


// ---- observation mirror (second half) ------------------------------------------------------------------------------
// The flattened observation row of a map point -- (key-frame, feature index, undistorted key-point, right coordinate, 1 / sigma^2 of its
// octave), sorted by key-frame id -- only changes when MapPoint::mObservations changes; the key-point data of a key-frame never changes.
// The mirror keeps the rows of all map points seen so far, keyed by mnId and validated by MapPoint::mnObsVersion (the counter the
// reference increments next to every change of mObservations).  A local-BA call then reads the unchanged rows back
// (one version compare + a contiguous copy per point) instead of copying 10^5 std::maps under their mutexes and chasing
// mvKeysUn / mvuRight / mvInvLevelSigma2 per observation: consecutive windows share almost all of their points.
struct ObsRec {
  KeyFrame *kf;
  uint32_t idx, kf_id;  // feature index in the key-frame; KeyFrame::mnId (so that flattening needs no pointer chase per observation)
  float u, v, ur, inv_sigma2;
};
struct Mirror {
  struct Row {
    MapPoint *mp = nullptr;
    unsigned long version = 0;
    uint32_t off = 0, n = 0;
  };
  std::vector<Row> rows;     // by MapPoint::mnId
  std::vector<ObsRec> pool;  // rows are appended; stale rows are dropped when the pool is compacted
  size_t live = 0;           // records referenced by current rows
  long long hits = 0, misses = 0;
  void clear() {
    rows.clear();
    pool.clear();
    live = 0;
  }
  void compact() {  // (only between calls: a window holds offsets into the pool) stale rows exceed the live ones: copy the live rows into a fresh pool
    if (pool.size() <= 2 * live + (1u << 20)) return;
    std::vector<ObsRec> np;
    np.reserve(live + live / 4);
    for (Row &r : rows)
      if (r.mp) {
        const uint32_t off = (uint32_t)np.size();
        np.insert(np.end(), pool.begin() + r.off, pool.begin() + r.off + r.n);
        r.off = off;
      }
    pool.swap(np);
  }
  // the row of pMP if it is current, else null (read-only: safe from several threads at once)
  const Row *find(MapPoint *pMP) const {
#ifdef PPO_HAVE_OBS_VERSION
    const unsigned long ver = pMP->mnObsVersion;
    const size_t id = pMP->mnId;
    if (ver == 0 || id >= rows.size()) return nullptr;
    const Row &r = rows[id];
    return (r.mp == pMP && r.version == ver) ? &r : nullptr;
#else
    (void)pMP;
    return nullptr;
#endif
  }
  // current row of pMP (rebuilt from GetObservations() when the map point changed or is new)
  const Row &row(MapPoint *pMP) {
#ifdef PPO_HAVE_OBS_VERSION
    const unsigned long ver = pMP->mnObsVersion;
#else
    const unsigned long ver = 0;  // no counter in this build of the reference: every call rebuilds every row
#endif
    const size_t id = pMP->mnId;
    if (id >= rows.size()) rows.resize(std::max(id + 1, rows.size() * 2));
    Row &r = rows[id];
    if (ver != 0 && r.mp == pMP && r.version == ver) {
      hits++;
      return r;
    }
    misses++;
    if (r.mp) live -= r.n;
    const std::map<KeyFrame *, size_t> observations = pMP->GetObservations();
    r.mp = pMP, r.version = ver, r.off = (uint32_t)pool.size(), r.n = (uint32_t)observations.size();
    for (auto &mit : observations) {
      KeyFrame *pKFi = mit.first;
      const size_t idx = mit.second;
      const cv::KeyPoint &kpUn = pKFi->mvKeysUn[idx];
      pool.push_back({pKFi, (uint32_t)idx, (uint32_t)pKFi->mnId, kpUn.pt.x, kpUn.pt.y, pKFi->mvuRight[idx] < 0 ? -1.0f : pKFi->mvuRight[idx],
                      pKFi->mvInvLevelSigma2[kpUn.octave]});
    }
    // key-frame slots of a window are ordered by mnId, so rows sorted by mnId come out in edge order
    std::sort(pool.begin() + r.off, pool.end(), [](const ObsRec &a, const ObsRec &b) { return a.kf_id < b.kf_id; });
    live += r.n;
    return r;
  }
};
static Mirror g_mirror;  // guarded by the mutex of the local-BA slot (run() is its only user)

struct Window {
  std::vector<KeyFrame *> lLocalKeyFrames, lFixedCameras;
  std::vector<MapPoint *> lLocalMapPoints;
  std::vector<std::pair<uint32_t, uint32_t>> obs_row;  // per local map point: (offset, count) of its row in the mirror's pool
  std::vector<MapCuboid *> lLocalMapCuboids;
  std::vector<MapPlane *> lLocalMapPlanes;
};

// Host threads of the flattening loops: a small pool of workers that SLEEP between jobs (condition variable, no spinning -- the
// caller is the LocalMapping thread of a process whose other threads track and close loops).  PPO_SHIM_THREADS overrides the
// size (default: min(8, hardware threads)); 1 = the serial loops.  A job is split into equal contiguous index ranges.
class HostPool {
 public:
  HostPool() {
    int n = (int)std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
    if (const char *e = std::getenv("PPO_SHIM_THREADS")) n = std::max(1, std::min(64, std::atoi(e)));
    n_ = n;
  }
  ~HostPool() {
    {
      std::lock_guard<std::mutex> lk(m_);
      quit_ = true;
    }
    cv_.notify_all();
    for (auto &t : th_) t.join();
  }
  int threads() const { return n_; }
  void resize(int n) {  // (not while a job runs: callers hold the mutex of the local-BA slot)
    {
      std::lock_guard<std::mutex> lk(m_);
      quit_ = true;
    }
    cv_.notify_all();
    for (auto &t : th_) t.join();
    th_.clear();
    quit_ = false;
    n_ = std::max(1, std::min(64, n));
  }
  // number of ranges for_ranges / for_parts split n indices into
  int parts_for(long n, long min_per_part = 2048) const { return (int)std::min<long>(n_, std::max<long>(1, n / min_per_part)); }
  // f(part, begin, end): like for_ranges, for loops that leave one result per range (part = 0 .. parts_for(n) - 1, in index order)
  void for_parts(long n, const std::function<void(int, long, long)> &f, long min_per_part = 2048) {
    const int parts = parts_for(n, min_per_part);
    for_ranges(n, [&](long b, long e) { f(parts <= 1 ? 0 : (int)((b * parts + n - 1) / n), b, e); }, min_per_part);
  }
  // f(begin, end) on disjoint ranges covering [0, n); returns when all ranges are done
  void for_ranges(long n, const std::function<void(long, long)> &f, long min_per_part = 2048) {
    const int parts = parts_for(n, min_per_part);
    if (parts <= 1) {
      f(0, n);
      return;
    }
    if (th_.empty())
      for (int i = 1; i < n_; i++) th_.emplace_back([this, i, g = gen_] { worker(i, g); });  // (a new worker must not mistake an old generation for a job)
    {
      std::lock_guard<std::mutex> lk(m_);
      job_ = &f, job_n_ = n, parts_ = parts, pending_ = parts - 1, gen_++;
    }
    cv_.notify_all();
    f(0, n / parts);
    std::unique_lock<std::mutex> lk(m_);
    done_.wait(lk, [this] { return pending_ == 0; });
  }

 private:
  void worker(int id, int seen) {
    for (;;) {
      const std::function<void(long, long)> *f;
      long n;
      int parts;
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return quit_ || gen_ != seen; });
        if (quit_) return;
        seen = gen_, f = job_, n = job_n_, parts = parts_;
      }
      if (id >= parts) continue;
      (*f)(n * id / parts, n * (id + 1) / parts);
      {
        std::lock_guard<std::mutex> lk(m_);
        pending_--;
      }
      done_.notify_one();
    }
  }
  int n_ = 1;
  std::vector<std::thread> th_;
  std::mutex m_;
  std::condition_variable cv_, done_;
  const std::function<void(long, long)> *job_ = nullptr;
  long job_n_ = 0;
  int parts_ = 0, pending_ = 0, gen_ = 0;
  bool quit_ = false;
};
static HostPool g_pool;  // (used under the mutex of the local-BA slot)

// stage A: Optimizer.cc:1997-2100 (mixed) / :463-514 (points only)
static void collect(KeyFrame *pKF, bool mixed, Window &w, bool with_planes = true) {
  static const bool timing = std::getenv("PPO_BA_TIMING") != nullptr;
  auto t_prev = std::chrono::steady_clock::now();
  auto tick = [&](const char *what) {
    if (!timing) return;
    const auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[ppo shim]   A: %-25s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_prev).count());
    t_prev = now;
  };
  w.lLocalKeyFrames.push_back(pKF);
  pKF->mnBALocalForKF = pKF->mnId;
  const std::vector<KeyFrame *> vNeighKFs = pKF->GetVectorCovisibleKeyFrames();
  for (KeyFrame *pKFi : vNeighKFs) {
    pKFi->mnBALocalForKF = pKF->mnId;
    if (!pKFi->isBad()) w.lLocalKeyFrames.push_back(pKFi);
  }
  {
    // Two phases with the result of the reference's loop: (1) every local key-frame's matches are copied (under the key-frame's mutex, as
    // the reference does) and the live, not yet marked map points among them kept with their ids -- the part that walks 10^5..10^6 pointers
    // into the map, spread over the host threads by key-frame; (2) the first-seen order, serially in key-frame order, on a byte map
    // indexed by MapPoint::mnId instead of a second look at every map point.  The mnBALocalForKF marks are set in the pass that reads
    // the mirror rows below (one look at each local map point).
    struct Cand {
      MapPoint *p;
      size_t id;
    };
    // (1) per RANGE of key-frames: candidates in key-frame order, already without the repeats inside the range
    // "seen" maps are indexed by MapPoint::mnId, which only grows over a session: they hold the number of the CALL that last saw the id
    // instead of a flag, so that nothing has to be cleared per call
    static std::vector<std::vector<Cand>> cand;         // (guarded by the mutex of the local-BA slot; keep their memory across calls)
    static std::vector<std::vector<uint32_t>> seen_of;  // per range
    static std::vector<uint32_t> seen;
    static uint32_t call = 0;
    if (++call == 0) {  // (wrapped: forget everything once)
      for (auto &v : seen_of) std::fill(v.begin(), v.end(), 0u);
      std::fill(seen.begin(), seen.end(), 0u);
      call = 1;
    }
    const size_t nkf = w.lLocalKeyFrames.size();
    const int parts = g_pool.parts_for((long)nkf, 8);
    if ((int)cand.size() < parts) cand.resize(parts), seen_of.resize(parts);
    g_pool.for_parts((long)nkf, [&](int part, long i0, long i1) {
      std::vector<Cand> &c = cand[part];
      std::vector<uint32_t> &mine = seen_of[part];
      c.clear();
      for (long i = i0; i < i1; i++) {
        const std::vector<MapPoint *> vpMPs = w.lLocalKeyFrames[i]->GetMapPointMatches();
        for (MapPoint *pMP : vpMPs)
          if (pMP && !pMP->isBad() && pMP->mnBALocalForKF != pKF->mnId) {
            const size_t id = pMP->mnId;
            if (id >= mine.size()) mine.resize(id + 1 + mine.size() / 2, 0u);
            if (mine[id] == call) continue;
            mine[id] = call;
            c.push_back({pMP, id});
          }
      }
    }, 8);
    tick("live matches per key-frame");
    // (2) merge in range order
    for (int q = 0; q < parts; q++)
      for (const Cand &c : cand[q]) {
        if (c.id >= seen.size()) seen.resize(c.id + 1 + seen.size() / 2, 0u);
        if (seen[c.id] != call) {
          seen[c.id] = call;
          w.lLocalMapPoints.push_back(c.p);
        }
      }
    tick("first-seen order");
  }
  if (mixed) {
    for (KeyFrame *kf : w.lLocalKeyFrames)
      for (MapCuboid *pMC : kf->mvpMapCuboid)
        if (pMC && !pMC->isBad() && pMC->mnBALocalForKF != pKF->mnId) {
          w.lLocalMapCuboids.push_back(pMC);
          pMC->mnBALocalForKF = pKF->mnId;
        }
    for (KeyFrame *kf : w.lLocalKeyFrames)
      if (with_planes)
      for (MapPlane *pMP : kf->mvpMapPlanes)
        if (pMP && !pMP->isBad() && pMP->mnBALocalForKF != pKF->mnId) {
          w.lLocalMapPlanes.push_back(pMP);
          pMP->mnBALocalForKF = pKF->mnId;
        }
  }
  auto add_fixed = [&](KeyFrame *pKFi) {
    if (pKFi->mnBALocalForKF != pKF->mnId && pKFi->mnBAFixedForKF != pKF->mnId) {
      pKFi->mnBAFixedForKF = pKF->mnId;
      if (!pKFi->isBad()) w.lFixedCameras.push_back(pKFi);
    }
  };
  tick("cuboids, planes");
  g_mirror.compact();
  {
    // mirror rows: the unchanged ones (one version compare per map point) are looked up by all host threads, the others are rebuilt
    // serially afterwards (a rebuild appends to the pool); the pass also leaves the mnBALocalForKF mark of the reference's loop
    const size_t NP = w.lLocalMapPoints.size();
    const uint32_t MISS = 0xffffffffu;
    w.obs_row.assign(NP, {MISS, 0});
    std::atomic<long long> hits{0};
    g_pool.for_ranges((long)NP, [&](long i0, long i1) {
      long long h = 0;
      for (long ip = i0; ip < i1; ip++) {
        MapPoint *pMP = w.lLocalMapPoints[ip];
        pMP->mnBALocalForKF = pKF->mnId;
        if (const Mirror::Row *r = g_mirror.find(pMP)) {
          w.obs_row[ip] = {r->off, r->n};
          h++;
        }
      }
      hits += h;
    });
    g_mirror.hits += hits.load();
    std::vector<size_t> miss;
    for (size_t ip = 0; ip < NP; ip++)
      if (w.obs_row[ip].first == MISS) miss.push_back(ip);
    if (miss.size() < 4096 || g_pool.threads() == 1) {
      for (size_t ip : miss) {
        const Mirror::Row &r = g_mirror.row(w.lLocalMapPoints[ip]);
        w.obs_row[ip] = {r.off, r.n};
      }
    } else {
      // many rows to (re)build -- a new map, or a build of the reference without the version counter: the observation maps are copied and
      // flattened by all host threads into per-range buffers (what Mirror::row does for one point), then appended to the pool in order
      struct Built {
        size_t id;
        unsigned long version;
        uint32_t n;
      };
      const long NM = (long)miss.size();
      std::vector<Built> built(miss.size());
      std::vector<std::vector<ObsRec>> recs((size_t)g_pool.parts_for(NM));
      g_pool.for_parts(NM, [&](int part, long k0, long k1) {
        std::vector<ObsRec> &out = recs[(size_t)part];
        for (long k = k0; k < k1; k++) {
          MapPoint *pMP = w.lLocalMapPoints[miss[k]];
#ifdef PPO_HAVE_OBS_VERSION
          const unsigned long ver = pMP->mnObsVersion;
#else
          const unsigned long ver = 0;
#endif
          const std::map<KeyFrame *, size_t> observations = pMP->GetObservations();
          const size_t first = out.size();
          for (auto &mit : observations) {
            KeyFrame *pKFi = mit.first;
            const size_t idx = mit.second;
            const cv::KeyPoint &kpUn = pKFi->mvKeysUn[idx];
            out.push_back({pKFi, (uint32_t)idx, (uint32_t)pKFi->mnId, kpUn.pt.x, kpUn.pt.y, pKFi->mvuRight[idx] < 0 ? -1.0f : pKFi->mvuRight[idx],
                           pKFi->mvInvLevelSigma2[kpUn.octave]});
          }
          std::sort(out.begin() + first, out.end(), [](const ObsRec &a, const ObsRec &b) { return a.kf_id < b.kf_id; });
          built[k] = {(size_t)pMP->mnId, ver, (uint32_t)observations.size()};
        }
      });
      size_t max_id = 0, total = 0;
      for (const Built &bt : built) max_id = std::max(max_id, bt.id), total += bt.n;
      if (max_id >= g_mirror.rows.size()) g_mirror.rows.resize(std::max(max_id + 1, g_mirror.rows.size() * 2));
      g_mirror.pool.reserve(g_mirror.pool.size() + total);
      long k = 0;
      for (size_t part = 0; part < recs.size(); part++) {
        const long k1 = recs.size() <= 1 ? NM : (long)(NM * (long)(part + 1) / (long)recs.size());
        uint32_t off = (uint32_t)g_mirror.pool.size();
        for (; k < k1; k++) {
          Mirror::Row &r = g_mirror.rows[built[k].id];
          if (r.mp) g_mirror.live -= r.n;
          r.mp = w.lLocalMapPoints[miss[k]], r.version = built[k].version, r.off = off, r.n = built[k].n;
          g_mirror.live += r.n;
          w.obs_row[miss[k]] = {r.off, r.n};
          off += r.n;
        }
        g_mirror.pool.insert(g_mirror.pool.end(), recs[part].begin(), recs[part].end());
      }
      g_mirror.misses += NM;
    }
  }
  tick("mirror rows");
  {
    // fixed cameras in the order of their first appearance among the observations: every range of map points lists the key-frames it
    // meets (first appearance inside the range); the lists are merged in range order, which is the order of the serial scan
    struct Seen {
      uint32_t id;
      KeyFrame *kf;
    };
    const long NP = (long)w.obs_row.size();
    std::vector<std::vector<Seen>> found((size_t)g_pool.parts_for(NP));
    const ObsRec *pool = g_mirror.pool.data();
    g_pool.for_parts(NP, [&](int part, long i0, long i1) {
      std::vector<uint8_t> seen;  // by KeyFrame::mnId: one entry per key-frame, not per observation
      std::vector<Seen> &out = found[(size_t)part];
      for (long ip = i0; ip < i1; ip++) {
        const auto &r = w.obs_row[ip];
        for (uint32_t q = r.first; q < r.first + r.second; q++) {
          const ObsRec &o = pool[q];
          if (o.kf_id < seen.size() && seen[o.kf_id]) continue;
          if (o.kf_id >= seen.size()) seen.resize((size_t)o.kf_id + 1 + seen.size(), 0);
          seen[o.kf_id] = 1;
          out.push_back({o.kf_id, o.kf});
        }
      }
    });
    std::vector<uint8_t> seen;
    for (auto &list : found)
      for (const Seen &f : list) {
        if (f.id < seen.size() && seen[f.id]) continue;
        if (f.id >= seen.size()) seen.resize((size_t)f.id + 1 + seen.size(), 0);
        seen[f.id] = 1;
        add_fixed(f.kf);
      }
  }
  tick("fixed cameras");
  if (mixed)
    for (MapCuboid *pMC : w.lLocalMapCuboids) {
      std::unordered_map<KeyFrame *, size_t> observations = pMC->GetObservations();
      for (auto &mit : observations) add_fixed(mit.first);
    }
}

// cuboids2d: Optimizer::LocalBACameraPointCuboids2D (Optimizer.cc:1252-1992) -- the mixed window without planes (no plane vertices, plane
// edges or cuboid-plane edges) plus, with optimize_with_cuboid_3d, one EdgeSE3Cuboid per cuboid observation (:1764-1804)
static void run(KeyFrame *pKF, bool *pbStopFlag, Map *pMap, bool mixed, bool fixCamera, bool fixPoint, bool cuboids2d = false) {
  Slot &S = g_slots[0];
  g_last_slot = &S;
  std::lock_guard<std::mutex> lk(S.m);
  // PPO_BA_TIMING=1: host-side phase times of this call on stderr (diagnostics only)
  static const bool timing = std::getenv("PPO_BA_TIMING") != nullptr;
  auto t_prev = std::chrono::steady_clock::now();
  auto tick = [&](const char *what) {
    if (!timing) return;
    const auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[ppo shim] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_prev).count());
    t_prev = now;
  };
  Window w;
  collect(pKF, mixed, w, !cuboids2d);
  tick("collect window (stage A)");

  // ---- stage B: flatten (vertices) ------------------------------------------------------------------
  Flat &F = S.last;
  F.clear(true);
  // key-frame slots ordered by mnId = g2o's Hessian order (core/sparse_optimizer.cpp:166-190,482-487)
  struct Slot { KeyFrame *kf; bool fixed; };
  std::vector<Slot> slots;
  for (KeyFrame *kf : w.lLocalKeyFrames) slots.push_back({kf, kf->mnId == 0 || fixCamera});  // Optimizer.cc:2126-2128
  for (KeyFrame *kf : w.lFixedCameras) slots.push_back({kf, true});                        // :2141
  std::sort(slots.begin(), slots.end(), [](const Slot &a, const Slot &b) { return a.kf->mnId < b.kf->mnId; });
  std::map<KeyFrame *, int> kf_slot;
  long unsigned int max_id = 0;
  for (const Slot &sl : slots) max_id = std::max(max_id, sl.kf->mnId);
  std::vector<int> slot_of_id(slots.empty() ? 0 : max_id + 1, -1);  // O(1) key-frame -> slot for the 10^5..10^6 point edges
  std::vector<char> slot_bad(slots.size(), 0);
  for (size_t i = 0; i < slots.size(); i++) {
    kf_slot[slots[i].kf] = (int)i;
    slot_of_id[slots[i].kf->mnId] = (int)i;
    slot_bad[i] = slots[i].kf->isBad();
    float T[16];
    double p7[7];
    mat_to_float16(slots[i].kf->GetPose(), T);
    ppo::tcw_float_to_pose7(T, p7);  // Converter::toSE3Quat
    F.kf_pose.insert(F.kf_pose.end(), p7, p7 + 7);
    F.kf_fixed.push_back(slots[i].fixed);
    const float in[5] = {slots[i].kf->fx, slots[i].kf->fy, slots[i].kf->cx, slots[i].kf->cy, slots[i].kf->mbf};
    F.kf_intr.insert(F.kf_intr.end(), in, in + 5);
  }
  auto slot_of = [&](KeyFrame *k) -> int {  // slot of a key-frame of the window, -1 otherwise (O(1): the edge loops below look up 10^4 observations)
    const size_t id = k->mnId;
    if (id >= slot_of_id.size()) return -1;
    const int sl = slot_of_id[id];
    return (sl >= 0 && slots[sl].kf == k) ? sl : -1;
  };
  std::map<MapCuboid *, int> cu_index;
  for (MapCuboid *pMC : w.lLocalMapCuboids) {  // :2158-2178: estimate = cuboid_global_data, roll/pitch and height locked
    cu_index[pMC] = (int)cu_index.size();
    double c[10];
    cuboid_to10(pMC->cuboid_global_data, c);
    F.cu_state.insert(F.cu_state.end(), c, c + 10);
    F.cu_flags.push_back(PPO_CU_FIXROLLPITCH | PPO_CU_FIXHEIGHT);
  }
  std::map<MapPlane *, int> pl_index;
  for (MapPlane *pMP : w.lLocalMapPlanes) {  // :2208-2219
    pl_index[pMP] = (int)pl_index.size();
    cv::Mat m = pMP->GetWorldPos();
    const float c4[4] = {m.at<float>(0, 0), m.at<float>(1, 0), m.at<float>(2, 0), m.at<float>(3, 0)};
    double c[4];
    ppo::plane_float_to_coef(c4, c);  // Converter::toPlane3D
    F.pl_coef.insert(F.pl_coef.end(), c, c + 4);
  }
  tick("  B: vertices");
  // ---- plane edges :2222-2309 ---------------------------------------------------------------------------
  ppo_ba_params P;
  ppo_ba_default_params(&P);
  std::vector<std::pair<KeyFrame *, MapPlane *>> plane_edge_owner;  // for EdgePlane edges only (vpEdgeKFPlane / vpMapPlane)
  std::vector<int> plane_edge_is_obs;
  if (mixed) {
    const double angleInfo = 3282.8 / (plane_angle_info * plane_angle_info), disInfo = plane_dist_info * plane_dist_info;
    const double pvInfo = 3282.8 / (0.5 * 0.5);
    P.huber_plane = ppo::huber_delta(plane_chi);
    P.chi2_plane = plane_chi;
    P.huber_vp_plane = ppo::huber_delta(200.0);
    P.chi2_vp_plane = 200.0;
    P.huber_bbox = ppo::huber_delta(thHuberBbox2d);
    P.norm_bbox = thHuberBbox2d;
    P.huber_corner = ppo::huber_delta(thHuberConer2d);
    P.norm_corner = thHuberConer2d;
    P.huber_cuboid_plane = ppo::huber_delta(cuboid_plane_chi);
    P.huber_se3 = thHuberSE3;  // rk->setDelta(thHuberSE3): the threshold itself (:1794)
    P.norm_se3 = thHuberSE3;   // :1878
    if (optimize_with_plane_3d && !cuboids2d) {
      // Per plane the reference copies three observation maps (GetObservations / GetVerObservations / GetParObservations return by
      // value, under the plane's mutex): ranges of planes are flattened by the host threads into their own arrays, which are then
      // appended in plane order -- the edge order of the serial loop.
      struct Part {
        std::vector<int32_t> plane, kf;
        std::vector<uint8_t> kind;
        std::vector<double> meas, info;
        std::vector<std::pair<KeyFrame *, MapPlane *>> owner;
        std::vector<int> is_obs;
      };
      const long npl = (long)w.lLocalMapPlanes.size();
      std::vector<Part> parts((size_t)g_pool.parts_for(npl, 16));
      g_pool.for_parts(npl, [&](int part, long i0, long i1) {
        Part &o = parts[(size_t)part];
        for (long ip = i0; ip < i1; ip++) {
          MapPlane *pMP = w.lLocalMapPlanes[ip];
          const int ipl = (int)ip;  // (= pl_index[pMP]: the planes were numbered in this order)
          auto add = [&](const std::map<KeyFrame *, int> &obs, int kind) {
            for (auto &mit : obs) {
              KeyFrame *pKFi = mit.first;
              if (pKFi->isBad()) continue;
              const int sl_kf = slot_of(pKFi);
              if (sl_kf < 0) continue;  // optimizer.vertex(pKFi->mnId) == NULL -> continue (:2235-2236, q8)
              cv::Mat m = pKFi->mvPlaneCoefficients[mit.second];
              const float c4[4] = {m.at<float>(0, 0), m.at<float>(1, 0), m.at<float>(2, 0), m.at<float>(3, 0)};
              double c[4];
              ppo::plane_float_to_coef(c4, c);
              o.plane.push_back(ipl);
              o.kf.push_back(sl_kf);
              o.kind.push_back((uint8_t)kind);
              o.meas.insert(o.meas.end(), c, c + 4);
              const double info[3] = {kind == PPO_PLANE_OBS ? angleInfo : pvInfo, kind == PPO_PLANE_OBS ? angleInfo : pvInfo, kind == PPO_PLANE_OBS ? disInfo : 0.0};
              o.info.insert(o.info.end(), info, info + 3);
              o.owner.push_back({pKFi, pMP});
              o.is_obs.push_back(kind == PPO_PLANE_OBS);
            }
          };
          add(pMP->GetObservations(), PPO_PLANE_OBS);
          add(pMP->GetVerObservations(), PPO_PLANE_VER);
          add(pMP->GetParObservations(), PPO_PLANE_PAR);
        }
      }, 16);
      for (const Part &o : parts) {
        F.ple_plane.insert(F.ple_plane.end(), o.plane.begin(), o.plane.end());
        F.ple_kf.insert(F.ple_kf.end(), o.kf.begin(), o.kf.end());
        F.ple_kind.insert(F.ple_kind.end(), o.kind.begin(), o.kind.end());
        F.ple_meas.insert(F.ple_meas.end(), o.meas.begin(), o.meas.end());
        F.ple_info.insert(F.ple_info.end(), o.info.begin(), o.info.end());
        plane_edge_owner.insert(plane_edge_owner.end(), o.owner.begin(), o.owner.end());
        plane_edge_is_obs.insert(plane_edge_is_obs.end(), o.is_obs.begin(), o.is_obs.end());
      }
    }
  } else {
    P.solver = PPO_SOLVER_6_3;
  }
  tick("  B: plane edges");
  // ---- points and reprojection edges :2332-2424 (mixed) / :560-650 (points only) ----------------------------
  std::vector<MapPoint *> &graph_points = S.graph_points;  // points that got a vertex (mixed: Observations() != 1, q1)
  std::vector<std::pair<KeyFrame *, MapPoint *>> &point_edge_owner = S.point_edge_owner;
  // Two passes over the local map points, both spread over the host threads (every point only reads its own map point and its own row of
  // the mirror and writes its own range of the arrays): (1) does the point get a vertex, and how many of its observations become edges;
  // prefix sums give every point its vertex index and its edge range; (2) fill.  The result is the one the serial loop produces.
  const size_t NP = w.lLocalMapPoints.size();
  std::vector<uint32_t> &pt_first = S.pt_first, &e_first = S.e_first;  // (counts, then exclusive prefix sums)
  pt_first.assign(NP + 1, 0), e_first.assign(NP + 1, 0);
  auto edge_ok = [&](const ObsRec &o) {
    if (o.kf_id >= slot_of_id.size()) return -1;
    const int sl = slot_of_id[o.kf_id];
    return (sl < 0 || slots[sl].kf != o.kf || slot_bad[sl]) ? -1 : sl;  // (!pKFi->isBad(), :2352, read once per key-frame)
  };
  const ObsRec *pool = g_mirror.pool.data();
  g_pool.for_ranges((long)NP, [&](long ip0, long ip1) {
  for (long ip = ip0; ip < ip1; ip++) {
    MapPoint *pMP = w.lLocalMapPoints[ip];
    if (mixed && pMP->Observations() == 1) continue;  // :2336
    pt_first[ip + 1] = 1;
    const ObsRec *row = pool + w.obs_row[ip].first;  // the row read in stage A, already in key-frame (= slot) order
    uint32_t c = 0;
    for (uint32_t q = 0; q < w.obs_row[ip].second; q++) c += edge_ok(row[q]) >= 0;
    e_first[ip + 1] = c;
  }
  });
  for (size_t ip = 0; ip < NP; ip++) pt_first[ip + 1] += pt_first[ip], e_first[ip + 1] += e_first[ip];
  const size_t npt = pt_first[NP], ne = e_first[NP];
  graph_points.resize(npt);
  point_edge_owner.resize(ne);
  F.resize_pinned(F.pt_xyz, 3 * npt, F.pins[0]); F.pt_fixed.assign(npt, mixed && fixPoint); F.resize_pinned(F.pt_rowptr, npt + 1, F.pins[1]);
  F.resize_pinned(F.pe_kf, ne, F.pins[2]); F.resize_pinned(F.pe_obs, 3 * ne, F.pins[3]); F.resize_pinned(F.pe_invsigma2, ne, F.pins[4]);
  F.pt_rowptr[0] = 0;
  tick("  B: point arrays sized");
  g_pool.for_ranges((long)NP, [&](long ip0, long ip1) {
  for (long ip = ip0; ip < ip1; ip++) {
    if (pt_first[ip + 1] == pt_first[ip]) continue;
    MapPoint *pMP = w.lLocalMapPoints[ip];
    const size_t pi = pt_first[ip];
    graph_points[pi] = pMP;
    cv::Mat X = pMP->GetWorldPos();
    for (int i = 0; i < 3; i++) F.pt_xyz[3 * pi + i] = (double)X.at<float>(i, 0);  // Converter::toVector3d
    const ObsRec *row = pool + w.obs_row[ip].first;
    size_t e = e_first[ip];
    for (uint32_t q = 0; q < w.obs_row[ip].second; q++) {
      const ObsRec &o = row[q];
      const int sl = edge_ok(o);
      if (sl < 0) continue;
      F.pe_kf[e] = sl;
      F.pe_obs[3 * e] = o.u, F.pe_obs[3 * e + 1] = o.v, F.pe_obs[3 * e + 2] = o.ur;
      F.pe_invsigma2[e] = o.inv_sigma2;
      point_edge_owner[e] = {o.kf, pMP};
      e++;
    }
    F.pt_rowptr[pi + 1] = (int32_t)e_first[ip + 1];
  }
  });
  if (mixed) {
    tick("  B: points + point edges");
    // ---- camera-cuboid edges :2433-2551 ------------------------------------------------------------------
    for (int pass = 0; pass < 2; pass++) {
      if (pass == 0 ? !optimize_with_cuboid_2d : !optimize_with_corners_2d) continue;
      for (MapCuboid *pMCuboid : w.lLocalMapCuboids) {
        const std::unordered_map<KeyFrame *, size_t> observations = pMCuboid->GetObservations();
        for (auto &mit : observations) {
          KeyFrame *pKFi = mit.first;
          if (pKFi->isBad()) continue;
          const int sl_kf = slot_of(pKFi);
          if (sl_kf < 0) continue;
          const MapCuboid *local_object = pKFi->local_cuboids[mit.second];
          const int object_boundary_margin = 5;
          const cv::Rect bbox_2d = local_object->bbox_2d;
          if (!((bbox_2d.x > object_boundary_margin) && (bbox_2d.y > object_boundary_margin) &&
                (bbox_2d.x + bbox_2d.width < 640 - object_boundary_margin) && (bbox_2d.y + bbox_2d.height < 480 - object_boundary_margin)))
            continue;
          double m[16] = {0};
          if (pass == 0)
            for (int i = 0; i < 4; i++) m[i] = local_object->bbox_vec(i);
          else
            for (int i = 0; i < 8; i++) m[2 * i] = local_object->box_corners_2d(0, i), m[2 * i + 1] = local_object->box_corners_2d(1, i);
          const double s = (pass == 0 ? ba_weight_bbox : ba_weight_corner) * local_object->meas_quality;
          F.cbe_kf.push_back(sl_kf);
          F.cbe_cuboid.push_back(cu_index[pMCuboid]);
          F.cbe_kind.push_back(pass == 0 ? PPO_CUBOID_BBOX : PPO_CUBOID_CORNER);
          F.cbe_meas.insert(F.cbe_meas.end(), m, m + 16);
          F.cbe_info.push_back(s * s);
        }
      }
    }
    // ---- EdgeSE3Cuboid :1764-1804 (LocalBACameraPointCuboids2D only) --------------------------------------------------
    if (cuboids2d && optimize_with_cuboid_3d)
      for (MapCuboid *pMCuboid : w.lLocalMapCuboids) {
        const std::unordered_map<KeyFrame *, size_t> observations = pMCuboid->GetObservations();
        for (auto &mit : observations) {
          KeyFrame *pKFi = mit.first;
          if (pKFi->isBad()) continue;
          const int sl_kf = slot_of(pKFi);
          if (sl_kf < 0) continue;
          // the reference reads the measurement from the LANDMARK list of the key-frame, mvpMapCuboid[idx] (:1779), not from its local
          // detections, and addresses the vertex by object_graph_id (:1783); a vertex that is not in the window would be a NULL vertex there
          if (mit.second >= pKFi->mvpMapCuboid.size() || !pKFi->mvpMapCuboid[mit.second]) continue;
          const MapCuboid *local_object = pKFi->mvpMapCuboid[mit.second];
          int cu = -1;
          for (MapCuboid *pMC : w.lLocalMapCuboids)
            if (pMC->mnId == (long int)pMCuboid->object_graph_id) cu = cu_index[pMC];
          if (cu < 0) continue;
          double m[16] = {0};
          cuboid_to10(local_object->cuboid_local_meas, m);
          const double s = ba_weight_SE3 * 0.75;  // meas_quality = 0.75 (:1788)
          F.cbe_kf.push_back(sl_kf);
          F.cbe_cuboid.push_back(cu);
          F.cbe_kind.push_back(PPO_CUBOID_SE3);
          F.cbe_meas.insert(F.cbe_meas.end(), m, m + 16);
          F.cbe_info.push_back(s * s);
        }
      }
    tick("  B: camera-cuboid edges");
    // ---- point-cuboid edges :2556-2655 ----------------------------------------------------------------------
    F.pce_rowptr.push_back(0);
    if (optimize_with_pt_obj_3d) {
      const int point_object_threshold = 2;
      const double coarse_threshold = 4, fine_threshold = 3;
      for (MapCuboid *pMC : w.lLocalMapCuboids) {
        std::vector<MapPoint *> pts;
        std::vector<std::array<double, 3>> xyz;
        const std::vector<MapPoint *> &UniquePoints = pMC->GetUniqueMapPoints();
        for (MapPoint *up : UniquePoints)
          if (up && !up->isBad() && up->MapObjObservations[pMC] > point_object_threshold) {
            cv::Mat X = up->GetWorldPos();
            pts.push_back(up);
            xyz.push_back({(double)X.at<float>(0, 0), (double)X.at<float>(1, 0), (double)X.at<float>(2, 0)});
          }
        pMC->used_points_in_BA_filtered.clear();
        double mean[3] = {0, 0, 0}, mean2[3] = {0, 0, 0};
        for (auto &p : xyz) for (int i = 0; i < 3; i++) mean[i] += p[i];
        for (int i = 0; i < 3; i++) mean[i] /= (double)xyz.size();
        int valid = 0;
        auto dist = [](const double a[3], const std::array<double, 3> &b) {
          return std::sqrt((a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) + (a[2] - b[2]) * (a[2] - b[2]));
        };
        for (auto &p : xyz)
          if (dist(mean, p) < coarse_threshold) {
            for (int i = 0; i < 3; i++) mean2[i] += p[i];
            valid++;
          }
        for (int i = 0; i < 3; i++) mean2[i] /= (double)valid;
        std::vector<std::array<double, 3>> good;
        for (size_t j = 0; j < xyz.size(); j++)
          if (dist(mean2, xyz[j]) < fine_threshold) {
            good.push_back(xyz[j]);
            pMC->used_points_in_BA_filtered.push_back(pts[j]);
          }
        if (good.size() > 10) {
          for (auto &p : good) F.pce_pts.insert(F.pce_pts.end(), p.begin(), p.end());
          F.pce_cuboid.push_back(cu_index[pMC]);
          F.pce_rowptr.push_back((int32_t)(F.pce_pts.size() / 3));
        }
      }
    }
    // ---- cuboid-plane edges :2662-2714 -------------------------------------------------------------------------
    if (optimize_with_cuboid_plane && !cuboids2d) {
      const double a = 3282.8 / (cuboid_plane_angle_info * cuboid_plane_angle_info), d = cuboid_plane_dist_info * cuboid_plane_dist_info;
      for (MapPlane *pMP : w.lLocalMapPlanes) {
        if (pMP->asso_cuboid_id == 999) continue;
        int cu = -1;
        for (MapCuboid *pMC : w.lLocalMapCuboids)
          if ((long unsigned int)pMC->mnId == pMP->asso_cuboid_id) cu = cu_index[pMC];
        if (cu < 0) continue;  // cuboid not in the graph (:2688-2690)
        F.cpe_cuboid.push_back(cu);
        F.cpe_plane.push_back(pl_index[pMP]);
        for (int i = 0; i < 3; i++) F.cpe_meas.push_back(pMP->asso_cuboid_meas(i));
        F.cpe_info.push_back(a); F.cpe_info.push_back(a); F.cpe_info.push_back(d);
      }
    }
  }
  F.publish();
  tick("flatten graph (stage B)");

  if (pbStopFlag && *pbStopFlag) return;  // :2723-2725 — no optimisation, no write-back

  // ---- stages C-E on the GPU ----------------------------------------------------------------------------------
  ppo_ba_handle *h = engine(S, P);
  S.rc = PPO_E_NOGPU;
  if (!h) {
    std::fprintf(stderr, "ppo shim: no CUDA engine available; map left untouched\n");
    return;
  }
  std::memset(&S.res, 0, sizeof S.res);
  if ((S.rc = ppo_ba_set_graph(h, &F.g)) != PPO_OK ||
      (S.rc = ppo_ba_local_ba(h, reinterpret_cast<const volatile unsigned char *>(pbStopFlag), &S.res)) != PPO_OK) {
    std::fprintf(stderr, "ppo shim: engine error %d (%s); map left untouched\n", S.rc, ppo_ba_last_error(h));
    return;
  }
  tick("engine (stages C-E)");
  if (S.res.skipped) return;  // the stop flag was raised between our test and the engine's: the reference returns before any write-back (:2723-2725)
  // ---- stage F: erase lists :2840-2887 ---------------------------------------------------------------------------
  std::vector<std::pair<KeyFrame *, MapPoint *>> vToErase;
  std::vector<std::pair<KeyFrame *, MapPlane *>> vToErasePlane;
  {
    // the chi2 / depth test of every point edge runs on the device; only the indices of the edges that fail it come back (ascending,
    // i.e. in the order of the reference's loop over vpEdgesMono / vpEdgesStereo as the shim laid the edges out)
    const int32_t *bad = nullptr;
    int32_t n_bad = 0;
    if (F.g.n_pe && (S.rc = point_edge_outliers(h, F, 5.991, 7.815, &bad, &n_bad)) != PPO_OK) return;
    for (int32_t q = 0; q < n_bad; q++) {
      MapPoint *pMP = point_edge_owner[bad[q]].second;
      if (pMP->isBad()) continue;
      vToErase.push_back(point_edge_owner[bad[q]]);
    }
    if (F.g.n_ple) {
      std::vector<double> pchi(F.g.n_ple);
      ppo_ba_edge_chi2(h, PPO_EDGE_PLANE, pchi.data(), nullptr, nullptr);
      for (int e = 0; e < F.g.n_ple; e++)
        if (plane_edge_is_obs[e] && pchi[e] > P.chi2_plane) vToErasePlane.push_back(plane_edge_owner[e]);
    }
  }
  ppo_ba_state st;
  std::vector<double> o_kf(F.kf_pose.size()), o_pt(F.pt_xyz.size()), o_pl(F.pl_coef.size()), o_cu(F.cu_state.size());
  st.kf_pose = o_kf.data(); st.pt_xyz = o_pt.data(); st.pl_coef = o_pl.data(); st.cu_state = o_cu.data();
  if ((S.rc = ppo_ba_get_state(h, &st)) != PPO_OK) return;

  tick("erase lists + read-back (F)");
  // ---- stage G: write back under the map mutex :2892-2966 --------------------------------------------------------------
  std::unique_lock<std::mutex> lock(pMap->mMutexMapUpdate);
  for (auto &pr : vToErase) {
    pr.first->EraseMapPointMatch(pr.second);
    pr.second->EraseObservation(pr.first);
  }
  for (auto &pr : vToErasePlane) {
    pr.first->EraseMapPlaneMatch(pr.second);
    pr.second->EraseObservation(pr.first);
  }
  for (KeyFrame *kf : w.lLocalKeyFrames) {
    if (mixed) kf->mnBALocalForKF = 0;
    float T[16];
    ppo::pose7_to_tcw_float(&o_kf[7 * (size_t)kf_slot[kf]], T);  // Converter::toCvMat(SE3Quat)
    kf->SetPose(float16_to_mat(T));
  }
  // points without a vertex are skipped (the reference dereferences NULL there, q1).  Every map point is written through its own
  // mutexes (SetWorldPos, UpdateNormalAndDepth) and appears once in graph_points, so the loop is spread over the host threads; the
  // caller still holds mMutexMapUpdate for the whole write-back, as the reference does.
  g_pool.for_ranges((long)graph_points.size(), [&](long i0, long i1) {
    for (long i = i0; i < i1; i++) {
      MapPoint *pMP = graph_points[i];
      if (mixed) pMP->mnBALocalForKF = 0;
      cv::Mat X(3, 1, CV_32F);
      for (int k = 0; k < 3; k++) X.at<float>(k, 0) = (float)o_pt[3 * i + k];
      pMP->SetWorldPos(X);
      pMP->UpdateNormalAndDepth();
    }
  });
  if (mixed) {
    for (KeyFrame *kf : w.lFixedCameras) {
      kf->mnBAFixedForKF = 0;
      kf->mnBALocalForKF = 0;
    }
    for (MapCuboid *pMC : w.lLocalMapCuboids) {
      pMC->mnBALocalForKF = 0;
      pMC->obj_been_optimized = true;
      const double *c = &o_cu[10 * (size_t)cu_index[pMC]];
      const double p7[7] = {c[3], c[4], c[5], c[6], c[0], c[1], c[2]};
      float T[16];
      ppo::pose7_to_tcw_float(p7, T);
      pMC->SetWorldPos(float16_to_mat(T));
      cuboid_from10(c, pMC->cuboid_global_opti);
    }
    for (MapPlane *pMP : w.lLocalMapPlanes) {
      cv::Mat m(4, 1, CV_32F);
      for (int k = 0; k < 4; k++) m.at<float>(k, 0) = (float)o_pl[4 * (size_t)pl_index[pMP] + k];  // Converter::toCvMat(Plane3D)
      pMP->SetWorldPos(m);
    }
  }
}

// ---- Optimizer::BundleAdjustment (global BA after map initialisation / loop closure), Optimizer.cc:54-241 ------------------
// Same projection edges as the local BA, every non-bad key-frame free except mnId == 0, one optimize(nIterations) and no
// outlier pass; results go to the map directly (nLoopKF == 0) or to mTcwGBA / mPosGBA for LoopClosing to merge.
static void run_global(const std::vector<KeyFrame *> &vpKFs, const std::vector<MapPoint *> &vpMP, int nIterations, bool *pbStopFlag,
                       unsigned long nLoopKF, bool bRobust) {
  Slot &S = g_slots[1];
  g_last_slot = &S;
  std::lock_guard<std::mutex> lk(S.m);
  Flat &F = S.last;
  F.clear();
  // key-frame vertices :73-86, slots ordered by mnId = g2o's Hessian order
  std::vector<KeyFrame *> kfs;
  for (KeyFrame *pKF : vpKFs)
    if (!pKF->isBad()) kfs.push_back(pKF);
  std::sort(kfs.begin(), kfs.end(), [](KeyFrame *a, KeyFrame *b) { return a->mnId < b->mnId; });
  long unsigned int maxKFid = 0;
  std::map<KeyFrame *, int> kf_slot;
  for (size_t i = 0; i < kfs.size(); i++) {
    KeyFrame *pKF = kfs[i];
    kf_slot[pKF] = (int)i;
    float T[16];
    double p7[7];
    mat_to_float16(pKF->GetPose(), T);
    ppo::tcw_float_to_pose7(T, p7);
    F.kf_pose.insert(F.kf_pose.end(), p7, p7 + 7);
    F.kf_fixed.push_back(pKF->mnId == 0);  // :81
    const float in[5] = {pKF->fx, pKF->fy, pKF->cx, pKF->cy, pKF->mbf};
    F.kf_intr.insert(F.kf_intr.end(), in, in + 5);
    if (pKF->mnId > maxKFid) maxKFid = pKF->mnId;
  }
  // map-point vertices and their projection edges :92-178
  std::vector<bool> vbNotIncludedMP(vpMP.size(), true);
  std::vector<MapPoint *> graph_points;
  for (size_t i = 0; i < vpMP.size(); i++) {
    MapPoint *pMP = vpMP[i];
    if (pMP->isBad()) continue;
    const std::map<KeyFrame *, size_t> observations = pMP->GetObservations();
    int nEdges = 0;
    for (auto mit = observations.begin(); mit != observations.end(); mit++) {
      KeyFrame *pKF = mit->first;
      if (pKF->isBad() || pKF->mnId > maxKFid) continue;  // :113-114
      auto slot = kf_slot.find(pKF);
      if (slot == kf_slot.end()) continue;  // key-frame without a vertex: the reference would add an edge to NULL
      nEdges++;
      const cv::KeyPoint &kpUn = pKF->mvKeysUn[mit->second];
      F.pe_kf.push_back(slot->second);
      F.pe_obs.push_back(kpUn.pt.x);
      F.pe_obs.push_back(kpUn.pt.y);
      F.pe_obs.push_back(pKF->mvuRight[mit->second]);  // < 0: monocular edge (:120)
      F.pe_invsigma2.push_back(pKF->mvInvLevelSigma2[kpUn.octave]);
    }
    if (nEdges == 0) continue;  // :168-172 (vertex removed again)
    vbNotIncludedMP[i] = false;
    if (F.pt_rowptr.empty()) F.pt_rowptr.push_back(0);
    cv::Mat X = pMP->GetWorldPos();
    for (int k = 0; k < 3; k++) F.pt_xyz.push_back((double)X.at<float>(k, 0));
    F.pt_fixed.push_back(0);
    F.pt_rowptr.push_back((int32_t)F.pe_kf.size());
    graph_points.push_back(pMP);
  }
  F.publish();

  ppo_ba_params P;
  ppo_ba_default_params(&P);
  P.solver = PPO_SOLVER_6_3;                // :62-66 (BlockSolver_6_3 + LinearSolverEigen: tile Cholesky, LDL^T fall-back on a failed factorisation)
  P.huber_mono = ppo::huber_delta(5.99);    // :88  "sqrt(5.99)", not the 5.991 of the local BA
  P.huber_stereo = ppo::huber_delta(7.815);  // :89
  ppo_ba_handle *h = engine(S, P);
  S.rc = PPO_E_NOGPU;
  if (!h) {
    std::fprintf(stderr, "ppo shim: no CUDA engine available; map left untouched\n");
    return;
  }
  std::memset(&S.res, 0, sizeof S.res);
  if ((S.rc = ppo_ba_set_graph(h, &F.g)) != PPO_OK) {
    std::fprintf(stderr, "ppo shim: engine error %d (%s); map left untouched\n", S.rc, ppo_ba_last_error(h));
    return;
  }
  if (!bRobust && F.g.n_pe) {  // :132-137,152-157: no robust kernel on any edge
    std::vector<unsigned char> flags(F.g.n_pe, 0);
    if ((S.rc = ppo_ba_set_edge_flags(h, PPO_EDGE_POINT, flags.data())) != PPO_OK) return;
  }
  // :180-183 initializeOptimization() + optimize(nIterations); the force-stop flag is polled inside (:69-70)
  if ((S.rc = ppo_ba_optimize(h, nIterations, reinterpret_cast<const volatile unsigned char *>(pbStopFlag), &S.res.round1)) != PPO_OK) {
    std::fprintf(stderr, "ppo shim: engine error %d (%s); map left untouched\n", S.rc, ppo_ba_last_error(h));
    return;
  }
  ppo_ba_state st;
  std::vector<double> o_kf(F.kf_pose.size()), o_pt(F.pt_xyz.size());
  st.kf_pose = o_kf.data(); st.pt_xyz = o_pt.data(); st.pl_coef = nullptr; st.cu_state = nullptr;
  if ((S.rc = ppo_ba_get_state(h, &st)) != PPO_OK) return;
  // ---- recover optimised data :187-239 --------------------------------------------------------------------------------------
  for (KeyFrame *pKF : kfs) {
    float T[16];
    ppo::pose7_to_tcw_float(&o_kf[7 * (size_t)kf_slot[pKF]], T);
    if (nLoopKF == 0) {
      pKF->SetPose(float16_to_mat(T));
    } else {
      pKF->mTcwGBA.create(4, 4, CV_32F);
      float16_to_mat(T).copyTo(pKF->mTcwGBA);
      pKF->mnBAGlobalForKF = nLoopKF;
    }
  }
  size_t gp = 0;
  for (size_t i = 0; i < vpMP.size(); i++) {
    if (vbNotIncludedMP[i]) continue;
    MapPoint *pMP = vpMP[i];
    cv::Mat X(3, 1, CV_32F);
    for (int k = 0; k < 3; k++) X.at<float>(k, 0) = (float)o_pt[3 * gp + k];
    gp++;
    if (nLoopKF == 0) {
      pMP->SetWorldPos(X);
      pMP->UpdateNormalAndDepth();
    } else {
      pMP->mPosGBA.create(3, 1, CV_32F);
      X.copyTo(pMP->mPosGBA);
      pMP->mnBAGlobalForKF = nLoopKF;
    }
  }
}

// ---- Optimizer::PoseOptimization (per tracked frame), Optimizer.cc:247-459 -------------------------------------------------------
// One free pose, every associated map point as a FIXED point with one projection edge (EdgeSE3ProjectXYZOnlyPose is the
// binary projection edge with its point held constant).  Four rounds of optimize(10) that each restart from pFrame->mTcw;
// after each round every edge is re-classified by chi2 (level-1 edges are re-evaluated first), the robust kernels go
// after the third round.  Returns the number of inliers.
static int run_pose(Frame *pFrame) {
  Slot &S = g_slots[2];
  g_last_slot = &S;
  std::lock_guard<std::mutex> lk(S.m);
  Flat &F = S.last;
  F.clear();
  {
    float T[16];
    double p7[7];
    mat_to_float16(pFrame->mTcw, T);
    ppo::tcw_float_to_pose7(T, p7);
    F.kf_pose.insert(F.kf_pose.end(), p7, p7 + 7);
    F.kf_fixed.push_back(0);  // :263 vSE3->setFixed(false)
    const float in[5] = {pFrame->fx, pFrame->fy, pFrame->cx, pFrame->cy, pFrame->mbf};
    F.kf_intr.insert(F.kf_intr.end(), in, in + 5);
  }
  int nInitialCorrespondences = 0;
  std::vector<int> vnIndexEdge;  // edge -> feature index (vnIndexEdgeMono / vnIndexEdgeStereo, :270-279)
  F.pt_rowptr.push_back(0);
  for (int i = 0; i < pFrame->N; i++) {
    MapPoint *pMP = pFrame->mvpMapPoints[i];
    if (!pMP) continue;
    nInitialCorrespondences++;
    pFrame->mvbOutlier[i] = false;  // :296,332
    const cv::KeyPoint &kpUn = pFrame->mvKeysUn[i];
    cv::Mat Xw = pMP->GetWorldPos();
    for (int k = 0; k < 3; k++) F.pt_xyz.push_back((double)Xw.at<float>(k, 0));  // e->Xw (:317-320)
    F.pt_fixed.push_back(1);
    F.pe_kf.push_back(0);
    F.pe_obs.push_back(kpUn.pt.x);
    F.pe_obs.push_back(kpUn.pt.y);
    F.pe_obs.push_back(pFrame->mvuRight[i]);  // < 0: monocular (:293)
    F.pe_invsigma2.push_back(pFrame->mvInvLevelSigma2[kpUn.octave]);
    F.pt_rowptr.push_back((int32_t)F.pe_kf.size());
    vnIndexEdge.push_back(i);
  }
  F.publish();
  if (nInitialCorrespondences < 3) return 0;  // :371-372

  ppo_ba_params P;
  ppo_ba_default_params(&P);  // deltaMono = sqrt(5.991), deltaStereo = sqrt(7.815) (:281-282)
  P.solver = PPO_SOLVER_6_3;
  ppo_ba_handle *h = engine(S, P);
  S.rc = PPO_E_NOGPU;
  if (!h) {
    std::fprintf(stderr, "ppo shim: no CUDA engine available; frame pose left untouched\n");
    return 0;
  }
  std::memset(&S.res, 0, sizeof S.res);
  if ((S.rc = ppo_ba_set_graph(h, &F.g)) != PPO_OK) {
    std::fprintf(stderr, "ppo shim: engine error %d (%s); frame pose left untouched\n", S.rc, ppo_ba_last_error(h));
    return 0;
  }
  const int n_e = F.g.n_pe;
  std::vector<unsigned char> flags(n_e, PPO_EF_ROBUST);
  std::vector<double> chi2(n_e);
  const float chi2Mono[4] = {5.991f, 5.991f, 5.991f, 5.991f}, chi2Stereo[4] = {7.815f, 7.815f, 7.815f, 7.815f};  // :376-377
  const int its[4] = {10, 10, 10, 10};
  int nBad = 0;
  for (size_t it = 0; it < 4; it++) {
    // :383-385  vSE3->setEstimate(toSE3Quat(pFrame->mTcw)); initializeOptimization(0); optimize(its[it])
    if ((S.rc = ppo_ba_reset(h)) != PPO_OK || (S.rc = ppo_ba_set_edge_flags(h, PPO_EDGE_POINT, flags.data())) != PPO_OK) return 0;
    S.rc = ppo_ba_optimize(h, its[it], nullptr, it < 2 ? &S.res.round1 : &S.res.round2);
    if (S.rc != PPO_OK && S.rc != PPO_E_EMPTY) return 0;  // (every edge an outlier: nothing to optimise)
    // :396-403 / :427-434  e->computeError() for the current outliers, then chi2 of every edge
    if ((S.rc = ppo_ba_recompute_edge_errors(h, PPO_EDGE_POINT)) != PPO_OK ||
        (S.rc = ppo_ba_edge_chi2(h, PPO_EDGE_POINT, chi2.data(), nullptr, nullptr)) != PPO_OK)
      return 0;
    nBad = 0;
    for (int e = 0; e < n_e; e++) {
      const bool mono = F.pe_obs[3 * (size_t)e + 2] < 0;
      const float c = (float)chi2[e];  // "const float chi2 = e->chi2();"
      const bool out = c > (mono ? chi2Mono[it] : chi2Stereo[it]);
      pFrame->mvbOutlier[vnIndexEdge[e]] = out;
      flags[e] = (unsigned char)((flags[e] & PPO_EF_ROBUST) | (out ? PPO_EF_LEVEL1 : 0));
      nBad += out;
      if (it == 2) flags[e] &= (unsigned char)~PPO_EF_ROBUST;  // :419-420
    }
    if (n_e < 10) break;  // :448-449
  }
  ppo_ba_state st;
  double o_kf[7];
  st.kf_pose = o_kf; st.pt_xyz = nullptr; st.pl_coef = nullptr; st.cu_state = nullptr;
  if ((S.rc = ppo_ba_get_state(h, &st)) != PPO_OK) return 0;
  float T[16];
  ppo::pose7_to_tcw_float(o_kf, T);
  pFrame->SetPose(float16_to_mat(T));  // :452-456
  return nInitialCorrespondences - nBad;
}

}  // namespace ppo_shim

namespace ORB_SLAM2 {
int Optimizer::PoseOptimization(Frame *pFrame) { return ppo_shim::run_pose(pFrame); }
void Optimizer::BundleAdjustment(const std::vector<KeyFrame *> &vpKFs, const std::vector<MapPoint *> &vpMP, int nIterations, bool *pbStopFlag,
                                 const unsigned long nLoopKF, const bool bRobust) {
  ppo_shim::run_global(vpKFs, vpMP, nIterations, pbStopFlag, nLoopKF, bRobust);
}
void Optimizer::GlobalBundleAdjustemnt(Map *pMap, int nIterations, bool *pbStopFlag, const unsigned long nLoopKF, const bool bRobust) {
  std::vector<KeyFrame *> vpKFs = pMap->GetAllKeyFrames();  // Optimizer.cc:46-51
  std::vector<MapPoint *> vpMP = pMap->GetAllMapPoints();
  BundleAdjustment(vpKFs, vpMP, nIterations, pbStopFlag, nLoopKF, bRobust);
}
void Optimizer::LocalBundleAdjustment(KeyFrame *pKF, bool *pbStopFlag, Map *pMap) { ppo_shim::run(pKF, pbStopFlag, pMap, false, false, false); }
void Optimizer::LocalBACameraPlaneCuboids(KeyFrame *pKF, bool *pbStopFlag, Map *pMap, bool fixCamera, bool fixPoint) {
  ppo_shim::run(pKF, pbStopFlag, pMap, true, fixCamera, fixPoint);
}
void Optimizer::LocalBACameraPointCuboids2D(KeyFrame *pKF, bool *pbStopFlag, Map *pMap, bool fixCamera, bool fixPoint) {
  ppo_shim::run(pKF, pbStopFlag, pMap, true, fixCamera, fixPoint, true);
}
}  // namespace ORB_SLAM2

// introspection for tests and logging
extern "C" {
// observation mirror: drop everything (a new map) / counters {rows reused, rows rebuilt, records in the pool}
void ppo_shim_mirror_clear() {
  std::lock_guard<std::mutex> lk(ppo_shim::g_slots[ppo_shim::SLOT_LOCAL].m);
  ppo_shim::g_mirror.clear();
}
void ppo_shim_mirror_stats(long long out[3]) {
  out[0] = ppo_shim::g_mirror.hits, out[1] = ppo_shim::g_mirror.misses, out[2] = (long long)ppo_shim::g_mirror.pool.size();
}
const ppo_ba_graph *ppo_shim_last_graph() { return &ppo_shim::g_last_slot.load()->last.g; }
const ppo_ba_result *ppo_shim_last_result() { return &ppo_shim::g_last_slot.load()->res; }
int ppo_shim_last_rc() { return ppo_shim::g_last_slot.load()->rc; }
void ppo_shim_set_device(int device) { ppo_shim::g_device = device; }
// host threads of the flattening loops of the local-BA entry points (default min(8, hardware threads), PPO_SHIM_THREADS); 1 = serial
void ppo_shim_set_threads(int n) {
  std::lock_guard<std::mutex> lk(ppo_shim::g_slots[ppo_shim::SLOT_LOCAL].m);
  ppo_shim::g_pool.resize(n);
}
int ppo_shim_get_threads() { return ppo_shim::g_pool.threads(); }
void ppo_shim_shutdown() {
  for (auto &S : ppo_shim::g_slots)
    if (S.h) ppo_ba_destroy(S.h), S.h = nullptr;
}
}
