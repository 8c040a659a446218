// Test harness for the Optimizer shim: builds a mock SLAM map (ppo_mock_slam.h) out of a flat synthetic graph —
// the inverse of the flattening the shim performs — calls ORB_SLAM2::Optimizer::LocalBACameraPlaneCuboids /
// LocalBundleAdjustment exactly like LocalMapping::Run does (src/LocalMapping.cc:100,107) and reads the written-back
// map state out again.  Input generation / bookkeeping only.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <vector>

#include "../../../include/ppo_ba.h"
#include "ppo_convert.h"
#include "ppo_mock_slam.h"

namespace ORB_SLAM2 {
// src/Parameters.cc:43-74 defaults relevant to the BA (flags are set per test from the graph content)
bool optimize_with_cuboid_plane = false, optimize_with_plane_3d = false, optimize_with_cuboid_2d = false, optimize_with_corners_2d = false,
     optimize_with_pt_obj_3d = false, optimize_with_cuboid_3d = false;
double ba_weight_bbox = 1.0, ba_weight_corner = 1.0, thHuberBbox2d = 80.0, thHuberConer2d = 10.0, ba_weight_SE3 = 1.0, thHuberSE3 = 900.0;
double plane_angle_info = 1.0, plane_dist_info = 100.0, plane_chi = 500.0, cuboid_plane_angle_info = 2.0, cuboid_plane_dist_info = 100.0,
       cuboid_plane_chi = 500.0;
}  // namespace ORB_SLAM2

using namespace ORB_SLAM2;

// test options of the mock map: every `bad_point_every`-th map point and key-frame slot `bad_kf` are flagged isBad()
// (0 / -1: none); they must be left out of the graph and untouched by the write-back
static int g_bad_point_every = 0, g_bad_kf = -1;
static double g_last_call_ms = 0;
// wall clock of the last Optimizer:: call itself (window collection + flattening + engine + write-back), without the mock map's construction
extern "C" double ppo_mock_last_call_ms() { return g_last_call_ms; }
extern "C" void ppo_mock_set_options(int bad_point_every, int bad_kf) { g_bad_point_every = bad_point_every, g_bad_kf = bad_kf; }

extern "C" void ppo_shim_mirror_clear();
// A mock map that outlives one Optimizer:: call (ppo_mock_world_*): consecutive local-BA calls on the same map, as LocalMapping makes them.
struct MockWorld {
  std::vector<std::unique_ptr<KeyFrame>> kfs;
  std::vector<std::unique_ptr<MapPoint>> pts, extra_pts;
  std::vector<std::unique_ptr<MapPlane>> pls;
  std::vector<std::unique_ptr<MapCuboid>> cus, local_cus;
  Map map;
  int pkf = 0;
  bool flags[6] = {false, false, false, false, false, false};
  const ppo_ba_graph *g = nullptr;  // (counts only, after build)
  int n_kf = 0, n_pt = 0, n_pl = 0, n_cu = 0;
};
static void world_build(MockWorld &W, const ppo_ba_graph *g);
static int world_run(MockWorld &W, int mixed, int fixCamera, int fixPoint, unsigned char *stop);
static void world_read(MockWorld &W, ppo_ba_state *out, int32_t counts[4]);

extern "C" int ppo_mock_run(const ppo_ba_graph *g, int mixed, int fixCamera, int fixPoint, unsigned char *stop, ppo_ba_state *out,
                            int32_t counts[4] /* erased point obs, erased plane obs, SetPose calls, UpdateNormalAndDepth calls */) {
  MockWorld W;
  world_build(W, g);
  world_run(W, mixed, fixCamera, fixPoint, stop);
  world_read(W, out, counts);
  ppo_shim_mirror_clear();  // the map objects die with W
  return 0;
}
extern "C" MockWorld *ppo_mock_world_create(const ppo_ba_graph *g) {
  MockWorld *W = new MockWorld();
  world_build(*W, g);
  return W;
}
extern "C" int ppo_mock_world_run(MockWorld *W, int mixed, int fixCamera, int fixPoint, unsigned char *stop, ppo_ba_state *out, int32_t counts[4]) {
  world_run(*W, mixed, fixCamera, fixPoint, stop);
  if (out) world_read(*W, out, counts);
  return 0;
}
// what Tracking / LocalMapping do to the map between two local-BA calls, reduced to the two operations the mirror must notice
extern "C" int ppo_mock_world_erase_observation(MockWorld *W, int point, int kf) {
  if (point < 0 || point >= W->n_pt || kf < 0 || kf >= W->n_kf) return -1;
  MapPoint *mp = W->pts[point].get();
  if (!mp->mObservations.count(W->kfs[kf].get())) return 0;
  mp->EraseObservation(W->kfs[kf].get());
  mp->erased.pop_back();  // (not an erasure made by the BA)
  mp->nObs -= 1;
  return 1;
}
extern "C" void ppo_mock_world_perturb_point(MockWorld *W, int point, float dx) {
  if (point >= 0 && point < W->n_pt) W->pts[point]->mWorldPos.at<float>(0, 0) += dx;
}
// puts the estimates of the flat graph back into the map (poses, points, planes, cuboids); observations stay as they are
extern "C" void ppo_mock_world_restore_estimates(MockWorld *W, const ppo_ba_graph *g) {
  for (int i = 0; i < g->n_kf && i < W->n_kf; i++) {
    float T[16];
    ppo::pose7_to_tcw_float(&g->kf_pose[7 * i], T);
    cv::Mat m(4, 4, CV_32F);
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 4; c++) m.at<float>(r, c) = T[4 * r + c];
    W->kfs[i]->Tcw = m;
  }
  for (int p = 0; p < g->n_pt && p < W->n_pt; p++)
    for (int k = 0; k < 3; k++) W->pts[p]->mWorldPos.at<float>(k, 0) = (float)g->pt_xyz[3 * p + k];
  for (int p = 0; p < g->n_pl && p < W->n_pl; p++)
    for (int k = 0; k < 4; k++) W->pls[p]->mWorldPos.at<float>(k, 0) = (float)g->pl_coef[4 * p + k];
  for (int c = 0; c < g->n_cu && c < W->n_cu; c++) {
    const double *s = &g->cu_state[10 * c];
    const double p7[7] = {s[3], s[4], s[5], s[6], s[0], s[1], s[2]};
    std::memcpy(W->cus[c]->cuboid_global_data.pose7, p7, sizeof p7);
    for (int k = 0; k < 3; k++) W->cus[c]->cuboid_global_data.scale[k] = s[7 + k];
    W->cus[c]->obj_been_optimized = false;
  }
}
extern "C" void ppo_mock_world_destroy(MockWorld *W) {
  delete W;
  ppo_shim_mirror_clear();
}

static void world_build(MockWorld &W, const ppo_ba_graph *g) {
  ppo_shim_mirror_clear();  // a new map: nothing cached may refer to the objects of an earlier one
  auto &kfs = W.kfs;
  auto &pts = W.pts;
  auto &extra_pts = W.extra_pts;
  auto &pls = W.pls;
  auto &cus = W.cus;
  auto &local_cus = W.local_cus;
  W.n_kf = g->n_kf, W.n_pt = g->n_pt, W.n_pl = g->n_pl, W.n_cu = g->n_cu;
  float inv_sigma2[8];
  {
    float sf = 1.0f;
    for (int i = 0; i < 8; i++) {
      inv_sigma2[i] = 1.0f / (sf * sf);
      sf *= 1.2f;
    }
  }
  // ---- key-frames: slot i -> mnId i; local = not fixed or slot 0; pKF = first free local key-frame --------------
  for (int i = 0; i < g->n_kf; i++) {
    const float *in = &g->kf_intr[5 * i];
    kfs.emplace_back(new KeyFrame(in[0], in[1], in[2], in[3], in[4]));
    KeyFrame *kf = kfs.back().get();
    kf->mnId = i;
    float T[16];
    ppo::pose7_to_tcw_float(&g->kf_pose[7 * i], T);
    cv::Mat m(4, 4, CV_32F);
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 4; c++) m.at<float>(r, c) = T[4 * r + c];
    kf->Tcw = m;
    kf->mvInvLevelSigma2.assign(inv_sigma2, inv_sigma2 + 8);
    kf->mbBad = i == g_bad_kf;
  }
  auto is_local = [&](int i) { return i == 0 || !g->kf_fixed[i]; };
  int pkf = -1;
  for (int i = 0; i < g->n_kf; i++)
    if (!g->kf_fixed[i]) { pkf = i; break; }
  if (pkf < 0) pkf = 0;
  for (int i = 0; i < g->n_kf; i++)
    if (i != pkf && is_local(i)) kfs[pkf]->mvpOrderedConnectedKeyFrames.push_back(kfs[i].get());
  // ---- map points and their observations ---------------------------------------------------------------------------
  for (int p = 0; p < g->n_pt; p++) {
    pts.emplace_back(new MapPoint());
    MapPoint *mp = pts.back().get();
    mp->mnId = p;
    mp->mbBad = g_bad_point_every > 0 && p % g_bad_point_every == g_bad_point_every - 1;
    cv::Mat X(3, 1, CV_32F);
    for (int k = 0; k < 3; k++) X.at<float>(k, 0) = (float)g->pt_xyz[3 * p + k];
    mp->mWorldPos = X;
    for (int e = g->pt_rowptr[p]; e < g->pt_rowptr[p + 1]; e++) {
      KeyFrame *kf = kfs[g->pe_kf[e]].get();
      const size_t idx = kf->mvKeysUn.size();
      int oct = 0;
      float best = 1e30f;
      for (int l = 0; l < 8; l++)
        if (std::fabs(inv_sigma2[l] - g->pe_invsigma2[e]) < best) best = std::fabs(inv_sigma2[l] - g->pe_invsigma2[e]), oct = l;
      kf->mvKeysUn.push_back(cv::KeyPoint{{g->pe_obs[3 * e], g->pe_obs[3 * e + 1]}, oct});
      kf->mvuRight.push_back(g->pe_obs[3 * e + 2]);
      mp->mObservations[kf] = idx;
      mp->nObs += g->pe_obs[3 * e + 2] >= 0 ? 2 : 1;  // MapPoint.cc:108-111
      if (is_local(g->pe_kf[e])) kf->mvpMapPoints.push_back(mp);
    }
  }
  // ---- planes -----------------------------------------------------------------------------------------------------------
  for (int p = 0; p < g->n_pl; p++) {
    pls.emplace_back(new MapPlane());
    MapPlane *pl = pls.back().get();
    pl->mnId = p;
    cv::Mat m(4, 1, CV_32F);
    for (int k = 0; k < 4; k++) m.at<float>(k, 0) = (float)g->pl_coef[4 * p + k];
    pl->mWorldPos = m;
  }
  for (int e = 0; e < g->n_ple; e++) {
    KeyFrame *kf = kfs[g->ple_kf[e]].get();
    MapPlane *pl = pls[g->ple_plane[e]].get();
    cv::Mat m(4, 1, CV_32F);
    for (int k = 0; k < 4; k++) m.at<float>(k, 0) = (float)g->ple_meas[4 * e + k];
    const int idx = (int)kf->mvPlaneCoefficients.size();
    kf->mvPlaneCoefficients.push_back(m);
    if (g->ple_kind[e] == PPO_PLANE_OBS) {
      pl->mObservations[kf] = idx;
      if (is_local(g->ple_kf[e])) kf->mvpMapPlanes.push_back(pl);
    } else if (g->ple_kind[e] == PPO_PLANE_VER) pl->mVerObservations[kf] = idx;
    else pl->mParObservations[kf] = idx;
  }
  // ---- cuboids --------------------------------------------------------------------------------------------------------------
  for (int c = 0; c < g->n_cu; c++) {
    cus.emplace_back(new MapCuboid());
    MapCuboid *cu = cus.back().get();
    cu->mnId = c;
    cu->object_graph_id = c;
    const double *s = &g->cu_state[10 * c];
    const double p7[7] = {s[3], s[4], s[5], s[6], s[0], s[1], s[2]};
    std::memcpy(cu->cuboid_global_data.pose7, p7, sizeof p7);
    for (int k = 0; k < 3; k++) cu->cuboid_global_data.scale[k] = s[7 + k];
  }
  bool any_bbox = false, any_corner = false, any_se3 = false;
  std::vector<char> any_se3_of((size_t)std::max(g->n_cu, 1), 0);
  for (int e = 0; e < g->n_cbe; e++) {
    KeyFrame *kf = kfs[g->cbe_kf[e]].get();
    MapCuboid *cu = cus[g->cbe_cuboid[e]].get();
    MapCuboid *lo = nullptr;
    auto it = cu->mObservations.find(kf);
    if (it == cu->mObservations.end()) {
      local_cus.emplace_back(new MapCuboid());
      lo = local_cus.back().get();
      lo->bbox_2d = cv::Rect{50, 50, 100, 100};  // inside the 5 px margin (the flat graph only holds edges that passed the test)
      if (g->cbe_kind[e] != PPO_CUBOID_SE3) lo->meas_quality = std::sqrt(g->cbe_info[e]);
      cu->mObservations[kf] = kf->local_cuboids.size();
      kf->local_cuboids.push_back(lo);
      // mvpMapCuboid runs parallel to the key-frame's detections (Tracking associates detection i with landmark mvpMapCuboid[i]); stage A
      // only walks the lists of the LOCAL key-frames, and skips landmarks already marked
      kf->mvpMapCuboid.push_back(cu);
    } else {
      lo = kf->local_cuboids[it->second];
    }
    if (g->cbe_kind[e] == PPO_CUBOID_SE3) {
      // the reference takes the 3-D measurement from the LANDMARK (mvpMapCuboid[idx]->cuboid_local_meas, Optimizer.cc:1779,1785): one
      // measurement per landmark, whichever key-frame observes it -- the first such edge of the flat graph provides it
      if (!any_se3_of[g->cbe_cuboid[e]]) {
        const double *m = &g->cbe_meas[16 * e];
        const double p7[7] = {m[3], m[4], m[5], m[6], m[0], m[1], m[2]};
        std::memcpy(cu->cuboid_local_meas.pose7, p7, sizeof p7);
        for (int k = 0; k < 3; k++) cu->cuboid_local_meas.scale[k] = m[7 + k];
        any_se3_of[g->cbe_cuboid[e]] = 1;
      }
      any_se3 = true;
    } else if (g->cbe_kind[e] == PPO_CUBOID_BBOX) {
      any_bbox = true;
      for (int k = 0; k < 4; k++) lo->bbox_vec(k) = g->cbe_meas[16 * e + k];
    } else {
      any_corner = true;
      for (int k = 0; k < 8; k++) lo->box_corners_2d(0, k) = g->cbe_meas[16 * e + 2 * k], lo->box_corners_2d(1, k) = g->cbe_meas[16 * e + 2 * k + 1];
    }
  }
  for (int e = 0; e < g->n_pce; e++) {
    MapCuboid *cu = cus[g->pce_cuboid[e]].get();
    for (int j = g->pce_rowptr[e]; j < g->pce_rowptr[e + 1]; j++) {
      extra_pts.emplace_back(new MapPoint());
      MapPoint *mp = extra_pts.back().get();
      mp->mnId = 1000000 + j;
      cv::Mat X(3, 1, CV_32F);
      for (int k = 0; k < 3; k++) X.at<float>(k, 0) = (float)g->pce_pts[3 * j + k];
      mp->mWorldPos = X;
      mp->MapObjObservations[cu] = 3;  // > point_object_threshold (Optimizer.cc:2559,2573)
      cu->mappoints_unique_own.push_back(mp);
    }
  }
  for (int e = 0; e < g->n_cpe; e++) {
    MapPlane *pl = pls[g->cpe_plane[e]].get();
    pl->asso_cuboid_id = g->cpe_cuboid[e];
    for (int k = 0; k < 3; k++) pl->asso_cuboid_meas(k) = g->cpe_meas[3 * e + k];
  }
  W.pkf = pkf;
  W.flags[0] = g->n_ple > 0, W.flags[1] = any_bbox, W.flags[2] = any_corner, W.flags[3] = g->n_pce > 0, W.flags[4] = g->n_cpe > 0, W.flags[5] = any_se3;
}

// ---- the call LocalMapping::Run makes ------------------------------------------------------------------------------------------
static int world_run(MockWorld &W, int mixed, int fixCamera, int fixPoint, unsigned char *stop) {
  optimize_with_plane_3d = W.flags[0];
  optimize_with_cuboid_2d = W.flags[1];
  optimize_with_corners_2d = W.flags[2];
  optimize_with_pt_obj_3d = W.flags[3];
  optimize_with_cuboid_plane = W.flags[4];
  optimize_with_cuboid_3d = W.flags[5];
  // every real call has a new pKF->mnId, which invalidates the mnBALocalForKF / mnBAFixedForKF marks of the call before; the mock calls
  // with the same key-frame again, so the marks are invalidated by hand
  const unsigned long none = ~0ul;
  for (auto &k : W.kfs) k->mnBALocalForKF = k->mnBAFixedForKF = none;
  for (auto &q : W.pts) q->mnBALocalForKF = none;
  for (auto &q : W.pls) q->mnBALocalForKF = none;
  for (auto &q : W.cus) q->mnBALocalForKF = none;
  bool stop_flag = stop ? (*stop != 0) : false;
  const auto t_call = std::chrono::steady_clock::now();
  if (mixed == 2) Optimizer::LocalBACameraPointCuboids2D(W.kfs[W.pkf].get(), &stop_flag, &W.map, fixCamera != 0, fixPoint != 0);
  else if (mixed) Optimizer::LocalBACameraPlaneCuboids(W.kfs[W.pkf].get(), &stop_flag, &W.map, fixCamera != 0, fixPoint != 0);
  else Optimizer::LocalBundleAdjustment(W.kfs[W.pkf].get(), &stop_flag, &W.map);
  g_last_call_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call).count();
  return 0;
}

// ---- read the map back ------------------------------------------------------------------------------------------------------------
static void world_read(MockWorld &W, ppo_ba_state *out, int32_t counts[4]) {
  auto &kfs = W.kfs;
  auto &pts = W.pts;
  auto &pls = W.pls;
  auto &cus = W.cus;
  struct { int n_kf, n_pt, n_pl, n_cu; } gs{W.n_kf, W.n_pt, W.n_pl, W.n_cu}, *g = &gs;
  counts[0] = counts[1] = counts[2] = counts[3] = 0;
  for (int i = 0; i < g->n_kf; i++) {
    float T[16];
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 4; c++) T[4 * r + c] = kfs[i]->Tcw.at<float>(r, c);
    ppo::tcw_float_to_pose7(T, &out->kf_pose[7 * i]);
    counts[2] += kfs[i]->n_setpose;
    counts[1] += (int)kfs[i]->erased_planes.size();
  }
  for (int p = 0; p < g->n_pt; p++) {
    for (int k = 0; k < 3; k++) out->pt_xyz[3 * p + k] = pts[p]->mWorldPos.at<float>(k, 0);
    counts[0] += (int)pts[p]->erased.size();
    counts[3] += pts[p]->n_updates;
  }
  for (int p = 0; p < g->n_pl; p++)
    for (int k = 0; k < 4; k++) out->pl_coef[4 * p + k] = pls[p]->mWorldPos.at<float>(k, 0);
  for (int c = 0; c < g->n_cu; c++) {
    const g2o::cuboid &q = cus[c]->obj_been_optimized ? cus[c]->cuboid_global_opti : cus[c]->cuboid_global_data;
    double *s = &out->cu_state[10 * c];
    s[0] = q.pose7[4]; s[1] = q.pose7[5]; s[2] = q.pose7[6];
    s[3] = q.pose7[0]; s[4] = q.pose7[1]; s[5] = q.pose7[2]; s[6] = q.pose7[3];
    s[7] = q.scale[0]; s[8] = q.scale[1]; s[9] = q.scale[2];
  }
}

// Global BA (Optimizer::GlobalBundleAdjustemnt, src/Optimizer.cc:46-241) on a mock map made of the key-frames and points
// of a flat graph: slot i -> mnId i (mnId 0 is the fixed one, whatever kf_fixed says), every key-frame and point registered
// in the Map.  Results are read from the map (nLoopKF == 0) or from mTcwGBA / mPosGBA (nLoopKF != 0).
extern "C" int ppo_mock_run_global(const ppo_ba_graph *g, int nIterations, unsigned long nLoopKF, int bRobust, unsigned char *stop, ppo_ba_state *out,
                                   int32_t counts[4] /* KFs tagged, points tagged, SetPose calls, UpdateNormalAndDepth calls */) {
  std::vector<std::unique_ptr<KeyFrame>> kfs;
  std::vector<std::unique_ptr<MapPoint>> pts;
  Map map;
  float inv_sigma2[8];
  {
    float sf = 1.0f;
    for (int i = 0; i < 8; i++) {
      inv_sigma2[i] = 1.0f / (sf * sf);
      sf *= 1.2f;
    }
  }
  for (int i = 0; i < g->n_kf; i++) {
    const float *in = &g->kf_intr[5 * i];
    kfs.emplace_back(new KeyFrame(in[0], in[1], in[2], in[3], in[4]));
    KeyFrame *kf = kfs.back().get();
    kf->mnId = i;
    float T[16];
    ppo::pose7_to_tcw_float(&g->kf_pose[7 * i], T);
    cv::Mat m(4, 4, CV_32F);
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 4; c++) m.at<float>(r, c) = T[4 * r + c];
    kf->Tcw = m;
    kf->mvInvLevelSigma2.assign(inv_sigma2, inv_sigma2 + 8);
    map.mvpKeyFrames.push_back(kf);
  }
  for (int p = 0; p < g->n_pt; p++) {
    pts.emplace_back(new MapPoint());
    MapPoint *mp = pts.back().get();
    mp->mnId = p;
    cv::Mat X(3, 1, CV_32F);
    for (int k = 0; k < 3; k++) X.at<float>(k, 0) = (float)g->pt_xyz[3 * p + k];
    mp->mWorldPos = X;
    for (int e = g->pt_rowptr[p]; e < g->pt_rowptr[p + 1]; e++) {
      KeyFrame *kf = kfs[g->pe_kf[e]].get();
      const size_t idx = kf->mvKeysUn.size();
      int oct = 0;
      float best = 1e30f;
      for (int l = 0; l < 8; l++)
        if (std::fabs(inv_sigma2[l] - g->pe_invsigma2[e]) < best) best = std::fabs(inv_sigma2[l] - g->pe_invsigma2[e]), oct = l;
      kf->mvKeysUn.push_back(cv::KeyPoint{{g->pe_obs[3 * e], g->pe_obs[3 * e + 1]}, oct});
      kf->mvuRight.push_back(g->pe_obs[3 * e + 2]);
      mp->mObservations[kf] = idx;
    }
    map.mvpMapPoints.push_back(mp);
  }
  bool stop_flag = stop ? (*stop != 0) : false;
  Optimizer::GlobalBundleAdjustemnt(&map, nIterations, &stop_flag, nLoopKF, bRobust != 0);

  counts[0] = counts[1] = counts[2] = counts[3] = 0;
  for (int i = 0; i < g->n_kf; i++) {
    const cv::Mat &M = nLoopKF ? kfs[i]->mTcwGBA : kfs[i]->Tcw;
    float T[16];
    if (M.empty()) {  // not written (engine unavailable): report the input
      for (int k = 0; k < 7; k++) out->kf_pose[7 * i + k] = g->kf_pose[7 * i + k];
    } else {
      for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) T[4 * r + c] = M.at<float>(r, c);
      ppo::tcw_float_to_pose7(T, &out->kf_pose[7 * i]);
    }
    counts[0] += nLoopKF && kfs[i]->mnBAGlobalForKF == nLoopKF;
    counts[2] += kfs[i]->n_setpose;
  }
  for (int p = 0; p < g->n_pt; p++) {
    const cv::Mat &X = (nLoopKF && !pts[p]->mPosGBA.empty()) ? pts[p]->mPosGBA : pts[p]->mWorldPos;
    for (int k = 0; k < 3; k++) out->pt_xyz[3 * p + k] = X.at<float>(k, 0);
    counts[1] += nLoopKF && pts[p]->mnBAGlobalForKF == nLoopKF;
    counts[3] += pts[p]->n_updates;
  }
  return 0;
}

// Tracking-style Optimizer::PoseOptimization call (src/Optimizer.cc:247-459; callers Tracking.cc:1006,1130,1173) on a
// mock Frame made of key-frame slot `kf` of a flat graph: one feature per point edge of that key-frame (its map point at
// the graph's position, rounded to float like MapPoint::GetWorldPos), plus a feature WITHOUT a map point after every
// seventh one.  outlier[] receives mvbOutlier of the associated features in edge order.
extern "C" int ppo_mock_run_pose(const ppo_ba_graph *g, int kf, double out_pose[7], unsigned char *outlier, int32_t counts[3] /* return value, SetPose calls, associated features */) {
  const float *in = &g->kf_intr[5 * kf];
  Frame frame(in[0], in[1], in[2], in[3], in[4]);
  float inv_sigma2[8];
  {
    float sf = 1.0f;
    for (int i = 0; i < 8; i++) {
      inv_sigma2[i] = 1.0f / (sf * sf);
      sf *= 1.2f;
    }
  }
  frame.mvInvLevelSigma2.assign(inv_sigma2, inv_sigma2 + 8);
  {
    float T[16];
    ppo::pose7_to_tcw_float(&g->kf_pose[7 * kf], T);
    cv::Mat m(4, 4, CV_32F);
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 4; c++) m.at<float>(r, c) = T[4 * r + c];
    frame.mTcw = m;
  }
  std::vector<std::unique_ptr<MapPoint>> pts;
  std::vector<int> assoc;  // feature index of every associated feature, in edge order
  for (int p = 0; p < g->n_pt; p++)
    for (int e = g->pt_rowptr[p]; e < g->pt_rowptr[p + 1]; e++) {
      if (g->pe_kf[e] != kf) continue;
      pts.emplace_back(new MapPoint());
      MapPoint *mp = pts.back().get();
      cv::Mat X(3, 1, CV_32F);
      for (int k = 0; k < 3; k++) X.at<float>(k, 0) = (float)g->pt_xyz[3 * p + k];
      mp->mWorldPos = X;
      int oct = 0;
      float best = 1e30f;
      for (int l = 0; l < 8; l++)
        if (std::fabs(inv_sigma2[l] - g->pe_invsigma2[e]) < best) best = std::fabs(inv_sigma2[l] - g->pe_invsigma2[e]), oct = l;
      assoc.push_back((int)frame.mvKeysUn.size());
      frame.mvKeysUn.push_back(cv::KeyPoint{{g->pe_obs[3 * e], g->pe_obs[3 * e + 1]}, oct});
      frame.mvuRight.push_back(g->pe_obs[3 * e + 2]);
      frame.mvpMapPoints.push_back(mp);
      frame.mvbOutlier.push_back(true);  // must be cleared by the call (:296,332)
      if (assoc.size() % 7 == 0) {  // a feature without a map point
        frame.mvKeysUn.push_back(cv::KeyPoint{{1.0f, 2.0f}, 0});
        frame.mvuRight.push_back(-1.0f);
        frame.mvpMapPoints.push_back(nullptr);
        frame.mvbOutlier.push_back(false);
      }
    }
  frame.N = (int)frame.mvKeysUn.size();
  counts[0] = Optimizer::PoseOptimization(&frame);
  counts[1] = frame.n_setpose;
  counts[2] = (int)assoc.size();
  float T[16];
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) T[4 * r + c] = frame.mTcw.at<float>(r, c);
  ppo::tcw_float_to_pose7(T, out_pose);
  for (size_t i = 0; i < assoc.size(); i++) outlier[i] = frame.mvbOutlier[assoc[i]] ? 1 : 0;
  return 0;
}
