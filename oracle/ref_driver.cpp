// ref_driver.cpp -- TEST INFRASTRUCTURE.  Drives the reference's OWN, UNMODIFIED solver code on the flat graph of
// include/ppo_ba.h: g2o (SparseOptimizer, BlockSolverX / BlockSolver_6_3, LinearSolverDense,
// OptimizationAlgorithmLevenberg, RobustKernelHuber, the numeric-Jacobian base edges) and the reference's vertex / edge
// types (VertexSE3Expmap, VertexSBAPointXYZ, EdgeSE3ProjectXYZ, EdgeStereoSE3ProjectXYZ, VertexPlane, EdgePlane,
// EdgeVerticalPlane, EdgeParallelPlane, VertexCuboid, EdgeSE3CuboidProj, EdgeSE3CuboidCornerProj,
// EdgePointCuboidOnlyObject, EdgeCuboidPlane), compiled from the sources where they lie under /root/reference by
// oracle/Makefile.ref against oracle/ref_stub/Eigen (Eigen itself is not installed here).  What this file adds is only
// the graph construction, written after Optimizer::LocalBACameraPlaneCuboids (src/Optimizer.cc:2107-2714: vertex ids,
// fixed / marginalised flags, information matrices, Huber kernels, insertion order), the call schedule
// optimize(5) -> re-levelling pass (:2736-2833) -> optimize(10), and accessors.  Optimizer.cc itself cannot be compiled
// (OpenCV / PCL / Pangolin / the ORB-SLAM2 map types).  The exported C API has the shape of the oracle's
// (ppo_oracle_*), so the same Python wrapper drives both and tests/test_ref_pin.py diffs them.
//
// Deviation (stated in DESIGN.md): the points-only solver flavour (PPO_SOLVER_6_3) uses BlockSolver_6_3 with
// LinearSolverDense instead of LinearSolverEigen (Eigen's sparse module is outside the stub); on the SPD systems of the
// tests both give the same solution up to rounding.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "../include/ppo_ba.h"

#include "Thirdparty/g2o/g2o/core/block_solver.h"
#include "Thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.h"
#include "Thirdparty/g2o/g2o/core/robust_kernel_impl.h"
#include "Thirdparty/g2o/g2o/core/sparse_optimizer.h"
#include "Thirdparty/g2o/g2o/solvers/linear_solver_dense.h"
#include "Thirdparty/g2o/g2o/types/types_six_dof_expmap.h"
#include "G2O_Plane3D.h"
#include "g2o_cuboid.h"

namespace {

struct PointEdgeRef {
  g2o::EdgeSE3ProjectXYZ *mono;
  g2o::EdgeStereoSE3ProjectXYZ *stereo;
  g2o::OptimizableGraph::Edge *e() const { return mono ? (g2o::OptimizableGraph::Edge *)mono : (g2o::OptimizableGraph::Edge *)stereo; }
  bool depth_positive() const { return mono ? mono->isDepthPositive() : stereo->isDepthPositive(); }
};

struct Handle {
  ppo_ba_params P;
  g2o::SparseOptimizer *opt = nullptr;
  g2o::OptimizationAlgorithmLevenberg *lm = nullptr;
  // flat copy of the input (for reset)
  ppo_ba_graph g;
  std::vector<double> kf_pose, pt_xyz, pl_coef, cu_state, ple_meas, ple_info, cbe_meas, cbe_info, pce_pts, cpe_meas, cpe_info;
  std::vector<uint8_t> kf_fixed, pt_fixed, cu_flags, ple_kind, cbe_kind;
  std::vector<float> kf_intr, pe_obs, pe_is2;
  std::vector<int32_t> pt_rowptr, pe_kf, ple_plane, ple_kf, cbe_kf, cbe_cuboid, pce_cuboid, pce_rowptr, cpe_cuboid, cpe_plane;
  // graph objects
  std::vector<g2o::VertexSE3Expmap *> v_kf;
  std::vector<g2o::VertexCuboid *> v_cu;
  std::vector<g2o::VertexPlane *> v_pl;
  std::vector<g2o::VertexSBAPointXYZ *> v_pt;
  std::vector<PointEdgeRef> e_pt;
  std::vector<g2o::OptimizableGraph::Edge *> e_pl, e_cb, e_pc, e_cp;

  ~Handle() { delete opt; }

  void build();
};

template <typename T>
void keep(std::vector<T> &dst, const T *&src, size_t n) {
  if (src && n) dst.assign(src, src + n);
  else dst.clear();
  src = dst.empty() ? nullptr : dst.data();
}

double huber_delta(g2o::OptimizableGraph::Edge *e) { return e->robustKernel() ? e->robustKernel()->delta() : 0.0; }

void Handle::build() {
  delete opt;
  v_kf.clear(), v_cu.clear(), v_pl.clear(), v_pt.clear(), e_pt.clear(), e_pl.clear(), e_cb.clear(), e_pc.clear(), e_cp.clear();
  opt = new g2o::SparseOptimizer();
  // Optimizer.cc:2108-2113 (BlockSolverX + LinearSolverDense) / :516-522 (BlockSolver_6_3)
  if (P.solver == PPO_SOLVER_6_3) {
    g2o::BlockSolver_6_3::LinearSolverType *ls = new g2o::LinearSolverDense<g2o::BlockSolver_6_3::PoseMatrixType>();
    lm = new g2o::OptimizationAlgorithmLevenberg(new g2o::BlockSolver_6_3(ls));
  } else {
    g2o::BlockSolverX::LinearSolverType *ls = new g2o::LinearSolverDense<g2o::BlockSolverX::PoseMatrixType>();
    lm = new g2o::OptimizationAlgorithmLevenberg(new g2o::BlockSolverX(ls));
  }
  opt->setAlgorithm(lm);
  opt->setVerbose(false);
  // vertex ids: key-frames 0 .. n_kf-1 (slot order = mnId order), cuboids, planes, points (typed spaces laid end to end, which
  // keeps g2o's ordering [poses | cuboids] [planes | points] of Optimizer.cc:2126-2348)
  const int id_cu = g.n_kf, id_pl = id_cu + g.n_cu, id_pt = id_pl + g.n_pl;
  for (int i = 0; i < g.n_kf; i++) {  // :2120-2146
    g2o::VertexSE3Expmap *v = new g2o::VertexSE3Expmap();
    const double *p = &kf_pose[7 * (size_t)i];
    v->setEstimate(g2o::SE3Quat(Eigen::Quaterniond(p[3], p[0], p[1], p[2]), Eigen::Vector3d(p[4], p[5], p[6])));
    v->setId(i);
    v->setFixed(kf_fixed[i] != 0);
    opt->addVertex(v);
    v_kf.push_back(v);
  }
  for (int i = 0; i < g.n_cu; i++) {  // :2155-2180
    g2o::VertexCuboid *v = new g2o::VertexCuboid();
    Vector10d c;
    for (int k = 0; k < 10; k++) c(k) = cu_state[10 * (size_t)i + k];
    g2o::cuboid cube;
    cube.fromVector(c);
    v->setEstimate(cube);
    v->whether_fixrollpitch = (cu_flags[i] & PPO_CU_FIXROLLPITCH) != 0;
    v->whether_fixheight = (cu_flags[i] & PPO_CU_FIXHEIGHT) != 0;
    v->setId(id_cu + i);
    v->setFixed(false);
    opt->addVertex(v);
    v_cu.push_back(v);
  }
  for (int i = 0; i < g.n_pl; i++) {  // :2208-2219
    g2o::VertexPlane *v = new g2o::VertexPlane();
    g2o::Vector4D c;
    for (int k = 0; k < 4; k++) c(k) = pl_coef[4 * (size_t)i + k];
    v->setEstimate(g2o::Plane3D(c));
    v->setId(id_pl + i);
    v->setMarginalized(true);
    opt->addVertex(v);
    v_pl.push_back(v);
  }
  for (int e = 0; e < g.n_ple; e++) {  // :2222-2309
    g2o::Vector4D m;
    for (int k = 0; k < 4; k++) m(k) = ple_meas[4 * (size_t)e + k];
    g2o::OptimizableGraph::Vertex *vp = v_pl[ple_plane[e]], *vk = v_kf[ple_kf[e]];
    g2o::RobustKernelHuber *rk = new g2o::RobustKernelHuber;
    if (ple_kind[e] == PPO_PLANE_OBS) {
      g2o::EdgePlane *ed = new g2o::EdgePlane();
      ed->setVertex(0, vp), ed->setVertex(1, vk);
      ed->setMeasurement(g2o::Plane3D(m));
      Eigen::Matrix3d Info;
      Info << ple_info[3 * (size_t)e], 0, 0, 0, ple_info[3 * (size_t)e + 1], 0, 0, 0, ple_info[3 * (size_t)e + 2];
      ed->setInformation(Info);
      ed->setRobustKernel(rk);
      rk->setDelta(P.huber_plane);
      opt->addEdge(ed);
      e_pl.push_back(ed);
    } else if (ple_kind[e] == PPO_PLANE_VER) {
      g2o::EdgeVerticalPlane *ed = new g2o::EdgeVerticalPlane();
      ed->setVertex(0, vp), ed->setVertex(1, vk);
      ed->setMeasurement(g2o::Plane3D(m));
      Eigen::Matrix2d Info;
      Info << ple_info[3 * (size_t)e], 0, 0, ple_info[3 * (size_t)e + 1];
      ed->setInformation(Info);
      ed->setRobustKernel(rk);
      rk->setDelta(P.huber_vp_plane);
      opt->addEdge(ed);
      e_pl.push_back(ed);
    } else {
      g2o::EdgeParallelPlane *ed = new g2o::EdgeParallelPlane();
      ed->setVertex(0, vp), ed->setVertex(1, vk);
      ed->setMeasurement(g2o::Plane3D(m));
      Eigen::Matrix2d Info;
      Info << ple_info[3 * (size_t)e], 0, 0, ple_info[3 * (size_t)e + 1];
      ed->setInformation(Info);
      ed->setRobustKernel(rk);
      rk->setDelta(P.huber_vp_plane);
      opt->addEdge(ed);
      e_pl.push_back(ed);
    }
  }
  for (int p = 0; p < g.n_pt; p++) {  // :2331-2424
    g2o::VertexSBAPointXYZ *v = new g2o::VertexSBAPointXYZ();
    v->setEstimate(Eigen::Vector3d(pt_xyz[3 * (size_t)p], pt_xyz[3 * (size_t)p + 1], pt_xyz[3 * (size_t)p + 2]));
    v->setId(id_pt + p);
    if (!pt_fixed.empty() && pt_fixed[p]) v->setFixed(true);
    else v->setMarginalized(true);
    opt->addVertex(v);
    v_pt.push_back(v);
    for (int e = pt_rowptr[p]; e < pt_rowptr[p + 1]; e++) {
      const int kf = pe_kf[e];
      const float *ob = &pe_obs[3 * (size_t)e], *in = &kf_intr[5 * (size_t)kf];
      const float &invSigma2 = pe_is2[e];
      g2o::RobustKernelHuber *rk = new g2o::RobustKernelHuber;
      PointEdgeRef ref = {nullptr, nullptr};
      if (ob[2] < 0) {
        Eigen::Matrix<double, 2, 1> obs;
        obs << ob[0], ob[1];
        g2o::EdgeSE3ProjectXYZ *ed = new g2o::EdgeSE3ProjectXYZ();
        ed->setVertex(0, v), ed->setVertex(1, v_kf[kf]);
        ed->setMeasurement(obs);
        ed->setInformation(Eigen::Matrix2d::Identity() * invSigma2);
        ed->setRobustKernel(rk);
        rk->setDelta(P.huber_mono);
        ed->fx = in[0], ed->fy = in[1], ed->cx = in[2], ed->cy = in[3];
        opt->addEdge(ed);
        ref.mono = ed;
      } else {
        Eigen::Matrix<double, 3, 1> obs;
        obs << ob[0], ob[1], ob[2];
        g2o::EdgeStereoSE3ProjectXYZ *ed = new g2o::EdgeStereoSE3ProjectXYZ();
        ed->setVertex(0, v), ed->setVertex(1, v_kf[kf]);
        ed->setMeasurement(obs);
        Eigen::Matrix3d Info = Eigen::Matrix3d::Identity() * invSigma2;
        ed->setInformation(Info);
        ed->setRobustKernel(rk);
        rk->setDelta(P.huber_stereo);
        ed->fx = in[0], ed->fy = in[1], ed->cx = in[2], ed->cy = in[3], ed->bf = in[4];
        opt->addEdge(ed);
        ref.stereo = ed;
      }
      e_pt.push_back(ref);
    }
  }
  for (int e = 0; e < g.n_cbe; e++) {  // :2433-2551
    const int kf = cbe_kf[e];
    const float *in = &kf_intr[5 * (size_t)kf];
    Eigen::Matrix3d calib;  // KeyFrame::mK as the shim's flattening passes it: [fx 0 cx; 0 fy cy; 0 0 1] in float
    calib << in[0], 0, in[2], 0, in[1], in[3], 0, 0, 1;
    g2o::RobustKernelHuber *rk = new g2o::RobustKernelHuber;
    if (cbe_kind[e] == PPO_CUBOID_BBOX) {
      g2o::EdgeSE3CuboidProj *ed = new g2o::EdgeSE3CuboidProj();
      ed->setVertex(0, v_kf[kf]), ed->setVertex(1, v_cu[cbe_cuboid[e]]);
      ed->Kalib = calib;
      Eigen::Vector4d m;
      for (int k = 0; k < 4; k++) m(k) = cbe_meas[16 * (size_t)e + k];
      ed->setMeasurement(m);
      Eigen::Matrix4d info = Eigen::Matrix4d::Identity() * cbe_info[e];
      ed->setInformation(info);
      ed->setRobustKernel(rk);
      rk->setDelta(P.huber_bbox);
      ed->setId(e);
      opt->addEdge(ed);
      e_cb.push_back(ed);
    } else if (cbe_kind[e] == PPO_CUBOID_SE3) {  // Optimizer.cc:1781-1798 (LocalBACameraPointCuboids2D)
      g2o::EdgeSE3Cuboid *ed = new g2o::EdgeSE3Cuboid();
      ed->setVertex(0, v_kf[kf]), ed->setVertex(1, v_cu[cbe_cuboid[e]]);
      Vector10d mv;
      for (int k = 0; k < 10; k++) mv(k) = cbe_meas[16 * (size_t)e + k];
      g2o::cuboid mc;
      mc.fromVector(mv);  // [t q scale], like the cuboid states
      ed->setMeasurement(mc);
      Eigen::Matrix<double, 9, 9> info = Eigen::Matrix<double, 9, 9>::Identity() * cbe_info[e];
      ed->setInformation(info);
      ed->setRobustKernel(rk);
      rk->setDelta(P.huber_se3);
      ed->setId(e);
      opt->addEdge(ed);
      e_cb.push_back(ed);
    } else {
      g2o::EdgeSE3CuboidCornerProj *ed = new g2o::EdgeSE3CuboidCornerProj();
      ed->setVertex(0, v_kf[kf]), ed->setVertex(1, v_cu[cbe_cuboid[e]]);
      ed->Kalib = calib;
      Eigen::Matrix<double, 16, 1> m;
      for (int k = 0; k < 16; k++) m(k) = cbe_meas[16 * (size_t)e + k];
      ed->setMeasurement(m);
      Eigen::Matrix<double, 16, 16> info = Eigen::Matrix<double, 16, 16>::Identity() * cbe_info[e];
      ed->setInformation(info);
      ed->setRobustKernel(rk);
      rk->setDelta(P.huber_corner);
      ed->setId(e);
      opt->addEdge(ed);
      e_cb.push_back(ed);
    }
  }
  for (int e = 0; e < g.n_pce; e++) {  // :2636-2655
    g2o::EdgePointCuboidOnlyObject *ed = new g2o::EdgePointCuboidOnlyObject();
    for (int j = pce_rowptr[e]; j < pce_rowptr[e + 1]; j++) ed->object_points.push_back(Eigen::Vector3d(pce_pts[3 * (size_t)j], pce_pts[3 * (size_t)j + 1], pce_pts[3 * (size_t)j + 2]));
    ed->setVertex(0, v_cu[pce_cuboid[e]]);
    Eigen::Matrix3d info;
    info.setIdentity();
    ed->setInformation(info);
    ed->max_outside_margin_ratio = P.ptcu_max_outside_margin_ratio;
    opt->addEdge(ed);
    e_pc.push_back(ed);
  }
  for (int e = 0; e < g.n_cpe; e++) {  // :2660-2712
    g2o::EdgeCuboidPlane *ed = new g2o::EdgeCuboidPlane();
    ed->setVertex(0, v_cu[cpe_cuboid[e]]), ed->setVertex(1, v_pl[cpe_plane[e]]);
    ed->setMeasurement(Eigen::Vector3d(cpe_meas[3 * (size_t)e], cpe_meas[3 * (size_t)e + 1], cpe_meas[3 * (size_t)e + 2]));
    Eigen::Matrix3d Info;
    Info << cpe_info[3 * (size_t)e], 0, 0, 0, cpe_info[3 * (size_t)e + 1], 0, 0, 0, cpe_info[3 * (size_t)e + 2];
    ed->setInformation(Info);
    g2o::RobustKernelHuber *rk = new g2o::RobustKernelHuber;
    ed->setRobustKernel(rk);
    rk->setDelta(P.huber_cuboid_plane);
    opt->addEdge(ed);
    ed->computeError();  // :2707 (debug print of the reference; it also leaves the error vector filled)
    e_cp.push_back(ed);
  }
}

std::vector<g2o::OptimizableGraph::Edge *> edges_of(Handle *h, int kind) {
  std::vector<g2o::OptimizableGraph::Edge *> v;
  switch (kind) {
    case PPO_EDGE_POINT:
      for (auto &r : h->e_pt) v.push_back(r.e());
      break;
    case PPO_EDGE_PLANE: v = h->e_pl; break;
    case PPO_EDGE_CUBOID_CAM: v = h->e_cb; break;
    case PPO_EDGE_POINT_CUBOID: v = h->e_pc; break;
    case PPO_EDGE_CUBOID_PLANE: v = h->e_cp; break;
  }
  return v;
}

double err_norm(g2o::OptimizableGraph::Edge *e) {
  double s = 0;
  const double *d = e->errorData();
  for (int i = 0; i < e->dimension(); i++) s += d[i] * d[i];
  return std::sqrt(s);
}

}  // namespace

extern "C" {

void ppo_ref_default_params(ppo_ba_params *p) {
  std::memset(p, 0, sizeof *p);
  auto hd = [](double th) { return (double)(float)std::sqrt(th); };  // "const float th = sqrt(..)" in Optimizer.cc
  p->huber_mono = hd(5.991), p->huber_stereo = hd(7.815), p->huber_plane = hd(500.0), p->huber_vp_plane = hd(200.0);
  p->huber_bbox = hd(80.0), p->huber_corner = hd(10.0), p->huber_cuboid_plane = hd(500.0);
  p->chi2_mono = 5.991, p->chi2_stereo = 7.815, p->chi2_plane = 500.0, p->chi2_vp_plane = 200.0, p->norm_bbox = 80.0, p->norm_corner = 10.0;
  p->lm_tau = 1e-5, p->lm_good_upper = 2. / 3., p->lm_good_lower = 1. / 3., p->lm_max_trials = 10;
  p->solver = PPO_SOLVER_DENSE_X, p->iters_round1 = 5, p->iters_round2 = 10;
  p->ptcu_max_outside_margin_ratio = 1.0, p->ptcu_prior_weight = 0.2;
  p->huber_se3 = 900.0, p->norm_se3 = 900.0;  // thHuberSE3 (Parameters.cc:65; Optimizer.cc:1794,1878)
}

int ppo_ref_create(const ppo_ba_params *params, void **out) {
  if (!params || !out) return PPO_E_INVALID;
  Handle *h = new Handle();
  h->P = *params;
  std::memset(&h->g, 0, sizeof h->g);
  *out = h;
  return PPO_OK;
}
void ppo_ref_destroy(void *hv) { delete (Handle *)hv; }

int ppo_ref_set_graph(void *hv, const ppo_ba_graph *gi) {
  Handle *h = (Handle *)hv;
  if (!h || !gi) return PPO_E_INVALID;
  h->g = *gi;
  ppo_ba_graph &g = h->g;
  keep(h->kf_pose, g.kf_pose, 7 * (size_t)g.n_kf), keep(h->kf_fixed, g.kf_fixed, (size_t)g.n_kf), keep(h->kf_intr, g.kf_intr, 5 * (size_t)g.n_kf);
  keep(h->pt_xyz, g.pt_xyz, 3 * (size_t)g.n_pt), keep(h->pt_fixed, g.pt_fixed, (size_t)g.n_pt);
  keep(h->pl_coef, g.pl_coef, 4 * (size_t)g.n_pl);
  keep(h->cu_state, g.cu_state, 10 * (size_t)g.n_cu), keep(h->cu_flags, g.cu_flags, (size_t)g.n_cu);
  keep(h->pt_rowptr, g.pt_rowptr, g.n_pt ? (size_t)g.n_pt + 1 : 0), keep(h->pe_kf, g.pe_kf, (size_t)g.n_pe), keep(h->pe_obs, g.pe_obs, 3 * (size_t)g.n_pe),
      keep(h->pe_is2, g.pe_invsigma2, (size_t)g.n_pe);
  keep(h->ple_plane, g.ple_plane, (size_t)g.n_ple), keep(h->ple_kf, g.ple_kf, (size_t)g.n_ple), keep(h->ple_kind, g.ple_kind, (size_t)g.n_ple),
      keep(h->ple_meas, g.ple_meas, 4 * (size_t)g.n_ple), keep(h->ple_info, g.ple_info, 3 * (size_t)g.n_ple);
  keep(h->cbe_kf, g.cbe_kf, (size_t)g.n_cbe), keep(h->cbe_cuboid, g.cbe_cuboid, (size_t)g.n_cbe), keep(h->cbe_kind, g.cbe_kind, (size_t)g.n_cbe),
      keep(h->cbe_meas, g.cbe_meas, 16 * (size_t)g.n_cbe), keep(h->cbe_info, g.cbe_info, (size_t)g.n_cbe);
  keep(h->pce_cuboid, g.pce_cuboid, (size_t)g.n_pce), keep(h->pce_rowptr, g.pce_rowptr, g.n_pce ? (size_t)g.n_pce + 1 : 0);
  keep(h->pce_pts, g.pce_pts, g.n_pce ? 3 * (size_t)h->pce_rowptr.back() : 0);
  keep(h->cpe_cuboid, g.cpe_cuboid, (size_t)g.n_cpe), keep(h->cpe_plane, g.cpe_plane, (size_t)g.n_cpe), keep(h->cpe_meas, g.cpe_meas, 3 * (size_t)g.n_cpe),
      keep(h->cpe_info, g.cpe_info, 3 * (size_t)g.n_cpe);
  h->build();
  return PPO_OK;
}
int ppo_ref_reset(void *hv) {
  Handle *h = (Handle *)hv;
  if (!h || !h->opt) return PPO_E_INVALID;
  h->build();
  return PPO_OK;
}

// SparseOptimizer::initializeOptimization(0) + optimize(iters)
int ppo_ref_optimize(void *hv, int iters, const volatile unsigned char *stop, ppo_ba_stats *st) {
  Handle *h = (Handle *)hv;
  if (!h || !h->opt) return PPO_E_INVALID;
  if (st) std::memset(st, 0, sizeof *st);
  bool flag = false;
  if (stop) {
    flag = *stop != 0;
    h->opt->setForceStopFlag(&flag);  // (the flag is sampled at entry: these tests do not raise it while the solver runs)
  }
  if (!h->opt->initializeOptimization(0)) return PPO_E_EMPTY;
  if (h->opt->indexMapping().size() == 0) return PPO_E_EMPTY;
  h->opt->computeActiveErrors();
  const double chi0 = h->opt->activeRobustChi2();
  const int done = h->opt->optimize(iters);
  if (st) {
    st->iterations = done < 0 ? 0 : done;
    st->chi2_initial = chi0;
    st->chi2_final = h->opt->activeRobustChi2();  // from the stored per-edge errors (stale after a rejected last trial, like e->chi2())
    st->n_active_edges = (int)h->opt->activeEdges().size();
    int np = 0, nl = 0;
    for (auto *v : h->opt->indexMapping()) {
      if (v->marginalized()) nl++;
      else np += v->dimension();
    }
    st->n_pose_dim = np, st->n_landmarks = nl;
    if (done > 0 && done <= PPO_TRACE_MAX) {
      st->trace[done - 1].lambda = h->lm->currentLambda();
      st->trace[done - 1].trials = h->lm->levenbergIteration();
    }
  }
  h->opt->setForceStopFlag(nullptr);
  return PPO_OK;
}

int ppo_ref_edge_count(const void *hv, int kind) {
  Handle *h = (Handle *)hv;
  switch (kind) {
    case PPO_EDGE_POINT: return (int)h->e_pt.size();
    case PPO_EDGE_PLANE: return (int)h->e_pl.size();
    case PPO_EDGE_CUBOID_CAM: return (int)h->e_cb.size();
    case PPO_EDGE_POINT_CUBOID: return (int)h->e_pc.size();
    case PPO_EDGE_CUBOID_PLANE: return (int)h->e_cp.size();
  }
  return -1;
}

int ppo_ref_edge_chi2(void *hv, int kind, double *chi2, unsigned char *depth_positive, double *norm) {
  Handle *h = (Handle *)hv;
  if (!h || !h->opt) return PPO_E_INVALID;
  std::vector<g2o::OptimizableGraph::Edge *> ev = edges_of(h, kind);
  for (size_t i = 0; i < ev.size(); i++) {
    if (chi2) chi2[i] = ev[i]->chi2();
    if (norm) norm[i] = err_norm(ev[i]);
    if (depth_positive) {
      if (kind == PPO_EDGE_POINT) depth_positive[i] = h->e_pt[i].depth_positive();
      else if (kind == PPO_EDGE_PLANE && dynamic_cast<g2o::EdgePlane *>(ev[i])) depth_positive[i] = dynamic_cast<g2o::EdgePlane *>(ev[i])->isDepthPositive();
      else depth_positive[i] = 1;
    }
  }
  return PPO_OK;
}
int ppo_ref_recompute_edge_errors(void *hv, int kind) {
  Handle *h = (Handle *)hv;
  if (!h || !h->opt || kind != PPO_EDGE_POINT) return PPO_E_INVALID;
  for (auto &r : h->e_pt)
    if (r.e()->level() == 1) r.e()->computeError();
  return PPO_OK;
}
int ppo_ref_get_edge_flags(void *hv, int kind, unsigned char *flags) {
  Handle *h = (Handle *)hv;
  std::vector<g2o::OptimizableGraph::Edge *> ev = edges_of(h, kind);
  for (size_t i = 0; i < ev.size(); i++) flags[i] = (ev[i]->level() == 1 ? PPO_EF_LEVEL1 : 0) | (ev[i]->robustKernel() ? PPO_EF_ROBUST : 0);
  return PPO_OK;
}
int ppo_ref_set_edge_flags(void *hv, int kind, const unsigned char *flags) {
  Handle *h = (Handle *)hv;
  std::vector<g2o::OptimizableGraph::Edge *> ev = edges_of(h, kind);
  for (size_t i = 0; i < ev.size(); i++) {
    ev[i]->setLevel((flags[i] & PPO_EF_LEVEL1) ? 1 : 0);
    if (!(flags[i] & PPO_EF_ROBUST)) ev[i]->setRobustKernel(0);
  }
  return PPO_OK;
}

// Optimizer.cc:2736-2833
int ppo_ref_outlier_pass(void *hv, int32_t n_out[3]) {
  Handle *h = (Handle *)hv;
  if (!h || !h->opt) return PPO_E_INVALID;
  int np = 0, npl = 0, ncb = 0;
  for (auto &r : h->e_pt) {
    g2o::OptimizableGraph::Edge *e = r.e();
    if (e->chi2() > (r.mono ? h->P.chi2_mono : h->P.chi2_stereo) || !r.depth_positive()) {
      np += e->level() != 1;
      e->setLevel(1);
    }
    e->setRobustKernel(0);
  }
  for (size_t i = 0; i < h->e_cb.size(); i++) {
    g2o::OptimizableGraph::Edge *e = h->e_cb[i];
    if (err_norm(e) > (h->cbe_kind[i] == PPO_CUBOID_BBOX ? h->P.norm_bbox : (h->cbe_kind[i] == PPO_CUBOID_SE3 ? h->P.norm_se3 : h->P.norm_corner))) {
      ncb += e->level() != 1;
      e->setLevel(1);
    }
  }
  for (size_t i = 0; i < h->e_pl.size(); i++) {
    g2o::OptimizableGraph::Edge *e = h->e_pl[i];
    if (e->chi2() > (h->ple_kind[i] == PPO_PLANE_OBS ? h->P.chi2_plane : h->P.chi2_vp_plane)) {
      npl += e->level() != 1;
      e->setLevel(1);
    }
    e->setRobustKernel(0);
  }
  if (n_out) n_out[0] = np, n_out[1] = npl, n_out[2] = ncb;
  return PPO_OK;
}

int ppo_ref_local_ba(void *hv, const volatile unsigned char *stop, ppo_ba_result *res) {
  Handle *h = (Handle *)hv;
  if (!h || !res) return PPO_E_INVALID;
  std::memset(res, 0, sizeof *res);
  if (stop && *stop) {
    res->skipped = 1;
    return PPO_OK;
  }
  int rc = ppo_ref_optimize(hv, h->P.iters_round1, stop, &res->round1);
  if (rc != PPO_OK && rc != PPO_E_EMPTY) return rc;
  if (!(stop && *stop)) {
    int32_t n_out[3];
    ppo_ref_outlier_pass(hv, n_out);
    res->n_outlier_point_edges = n_out[0], res->n_outlier_plane_edges = n_out[1], res->n_outlier_cuboid_edges = n_out[2];
    rc = ppo_ref_optimize(hv, h->P.iters_round2, stop, &res->round2);
    if (rc == PPO_E_EMPTY) rc = PPO_OK;
  }
  return rc;
}

int ppo_ref_get_state(void *hv, ppo_ba_state *out) {
  Handle *h = (Handle *)hv;
  if (!h || !h->opt || !out) return PPO_E_INVALID;
  if (out->kf_pose)
    for (size_t i = 0; i < h->v_kf.size(); i++) {
      const g2o::SE3Quat &T = h->v_kf[i]->estimate();
      double *p = out->kf_pose + 7 * i;
      p[0] = T.rotation().x(), p[1] = T.rotation().y(), p[2] = T.rotation().z(), p[3] = T.rotation().w();
      p[4] = T.translation()(0), p[5] = T.translation()(1), p[6] = T.translation()(2);
    }
  if (out->pt_xyz)
    for (size_t i = 0; i < h->v_pt.size(); i++)
      for (int k = 0; k < 3; k++) out->pt_xyz[3 * i + k] = h->v_pt[i]->estimate()(k);
  if (out->pl_coef)
    for (size_t i = 0; i < h->v_pl.size(); i++)
      for (int k = 0; k < 4; k++) out->pl_coef[4 * i + k] = h->v_pl[i]->estimate().coeffs()(k);
  if (out->cu_state)
    for (size_t i = 0; i < h->v_cu.size(); i++) {
      Vector10d v = h->v_cu[i]->estimate().toVector();
      for (int k = 0; k < 10; k++) out->cu_state[10 * i + k] = v(k);
    }
  return PPO_OK;
}

// ------------------------------------------------------------------------------------------------
// formula-level entry points: the same signatures as the oracle's (ppo_oracle_*), evaluated by the reference's own classes
// ------------------------------------------------------------------------------------------------
static g2o::SE3Quat se3_in(const double p[7]) { return g2o::SE3Quat(Eigen::Quaterniond(p[3], p[0], p[1], p[2]), Eigen::Vector3d(p[4], p[5], p[6])); }
static void se3_out(const g2o::SE3Quat &T, double p[7]) {
  p[0] = T.rotation().x(), p[1] = T.rotation().y(), p[2] = T.rotation().z(), p[3] = T.rotation().w();
  for (int i = 0; i < 3; i++) p[4 + i] = T.translation()(i);
}
static g2o::cuboid cu_in(const double c[10]) {
  Vector10d v;
  for (int k = 0; k < 10; k++) v(k) = c[k];
  g2o::cuboid q;
  q.fromVector(v);
  return q;
}
static void cu_out(const g2o::cuboid &q, double c[10]) {
  Vector10d v = q.toVector();
  for (int k = 0; k < 10; k++) c[k] = v(k);
}
static g2o::Plane3D pl_in(const double c[4]) {
  g2o::Vector4D v;
  for (int k = 0; k < 4; k++) v(k) = c[k];
  return g2o::Plane3D(v);
}
static Eigen::Matrix3d K_in(const float intr[5]) {
  Eigen::Matrix3d K;
  K << intr[0], 0, intr[2], 0, intr[1], intr[3], 0, 0, 1;
  return K;
}
void ppo_ref_se3_exp(const double u[6], double out[7]) {
  g2o::Vector6d v;
  for (int i = 0; i < 6; i++) v(i) = u[i];
  se3_out(g2o::SE3Quat::exp(v), out);
}
void ppo_ref_se3_from_Rt(const double R[9], const double t[3], double out[7]) {
  Eigen::Matrix3d m;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) m(i, j) = R[3 * i + j];
  se3_out(g2o::SE3Quat(m, Eigen::Vector3d(t[0], t[1], t[2])), out);
}
void ppo_ref_se3_oplus(const double pose[7], const double u[6], double out[7]) {
  g2o::VertexSE3Expmap v;
  v.setEstimate(se3_in(pose));
  v.oplusImpl(u);
  se3_out(v.estimate(), out);
}
void ppo_ref_se3_map(const double pose[7], const double p[3], double out[3]) {
  Eigen::Vector3d r = se3_in(pose).map(Eigen::Vector3d(p[0], p[1], p[2]));
  for (int i = 0; i < 3; i++) out[i] = r(i);
}
void ppo_ref_se3_matrix(const double pose[7], double R[9]) {
  Eigen::Matrix3d m = se3_in(pose).rotation().toRotationMatrix();
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) R[3 * i + j] = m(i, j);
}
void ppo_ref_plane_normalize(const double c[4], double out[4]) {
  g2o::Plane3D p = pl_in(c);
  for (int i = 0; i < 4; i++) out[i] = p.coeffs()(i);
}
void ppo_ref_plane_oplus(const double c[4], const double v[3], double out[4]) {
  g2o::VertexPlane vp;
  vp.setEstimate(pl_in(c));
  vp.oplusImpl(v);
  for (int i = 0; i < 4; i++) out[i] = vp.estimate().coeffs()(i);
}
void ppo_ref_plane_ominus(int kind, const double a[4], const double bq[4], double out[3]) {
  g2o::Plane3D p = pl_in(a), q = pl_in(bq);
  out[2] = 0;
  if (kind == PPO_PLANE_OBS) {
    g2o::Vector3D r = p.ominus(q);
    for (int i = 0; i < 3; i++) out[i] = r(i);
  } else {
    g2o::Vector2D r = kind == PPO_PLANE_VER ? p.ominus_ver(q) : p.ominus_par(q);
    out[0] = r(0), out[1] = r(1);
  }
}
void ppo_ref_plane_transform(const double pose[7], const double c[4], double out[4]) {
  g2o::Isometry3D w2n = se3_in(pose);
  g2o::Plane3D p = w2n * pl_in(c);
  for (int i = 0; i < 4; i++) out[i] = p.coeffs()(i);
}
void ppo_ref_cuboid_oplus(const double c[10], unsigned flags, const double u[9], double out[10]) {
  g2o::VertexCuboid v;
  v.setEstimate(cu_in(c));
  v.whether_fixrollpitch = (flags & PPO_CU_FIXROLLPITCH) != 0;
  v.whether_fixheight = (flags & PPO_CU_FIXHEIGHT) != 0;
  v.oplusImpl(u);
  cu_out(v.estimate(), out);
}
void ppo_ref_cuboid_corners(const double c[10], double out[24]) {
  Eigen::Matrix3Xd w = cu_in(c).compute3D_BoxCorner();
  for (int i = 0; i < 3; i++)
    for (int k = 0; k < 8; k++) out[8 * i + k] = w(i, k);
}
void ppo_ref_cuboid_project(const double c[10], const double pose[7], const float intr[5], double corners[16], double bbox[4]) {
  Eigen::Matrix2Xd p = cu_in(c).projectOntoImage(se3_in(pose), K_in(intr));
  for (int k = 0; k < 8; k++) corners[2 * k] = p(0, k), corners[2 * k + 1] = p(1, k);
  Eigen::Vector4d b = cu_in(c).projectOntoImageBbox(se3_in(pose), K_in(intr));
  for (int k = 0; k < 4; k++) bbox[k] = b(k);
}
void ppo_ref_cuboid_point_error(const double c[10], const double *pts, int n, double ratio, double prior_weight, double out[3]) {
  (void)prior_weight;  // 0.2 is a literal inside EdgePointCuboidOnlyObject::computeError (g2o_cuboid.cc:147)
  g2o::VertexCuboid v;
  v.setEstimate(cu_in(c));
  g2o::EdgePointCuboidOnlyObject e;
  e.setVertex(0, &v);
  for (int j = 0; j < n; j++) e.object_points.push_back(Eigen::Vector3d(pts[3 * j], pts[3 * j + 1], pts[3 * j + 2]));
  e.max_outside_margin_ratio = ratio;
  e.computeError();
  for (int i = 0; i < 3; i++) out[i] = e.error()(i);
  e.setVertex(0, nullptr);
}
void ppo_ref_cuboid_to_minimal(const double c[10], double out[9]) {
  Vector9d v = cu_in(c).toMinimalVector();
  for (int i = 0; i < 9; i++) out[i] = v(i);
}
void ppo_ref_huber(double e, double delta, double rho[3]) {
  g2o::RobustKernelHuber rk;
  rk.setDelta(delta);
  Eigen::Vector3d r;
  rk.robustify(e, r);
  for (int i = 0; i < 3; i++) rho[i] = r(i);
}
int ppo_ref_point_edge(const double pose[7], const double X[3], const float intr[5], const float obs[3], double err[3], double Jpt[9], double Jkf[18]) {
  g2o::VertexSBAPointXYZ vp;
  vp.setEstimate(Eigen::Vector3d(X[0], X[1], X[2]));
  g2o::VertexSE3Expmap vk;
  vk.setEstimate(se3_in(pose));
  int D;
  err[2] = 0;
  if (obs[2] < 0) {
    g2o::EdgeSE3ProjectXYZ e;
    e.setVertex(0, &vp), e.setVertex(1, &vk);
    Eigen::Matrix<double, 2, 1> o;
    o << obs[0], obs[1];
    e.setMeasurement(o);
    e.fx = intr[0], e.fy = intr[1], e.cx = intr[2], e.cy = intr[3];
    e.computeError();
    err[0] = e.error()(0), err[1] = e.error()(1);
    D = 2;
    if (Jpt && Jkf) {
      g2o::JacobianWorkspace ws;
      ws.updateSize(&e);
      ws.allocate();
      e.g2o::BaseBinaryEdge<2, Eigen::Vector2d, g2o::VertexSBAPointXYZ, g2o::VertexSE3Expmap>::linearizeOplus(ws);
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) Jpt[3 * r + c] = r < 2 ? e.jacobianOplusXi()(r, c) : 0.0;
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 6; c++) Jkf[6 * r + c] = r < 2 ? e.jacobianOplusXj()(r, c) : 0.0;
    }
    e.setVertex(0, nullptr), e.setVertex(1, nullptr);
  } else {
    g2o::EdgeStereoSE3ProjectXYZ e;
    e.setVertex(0, &vp), e.setVertex(1, &vk);
    Eigen::Matrix<double, 3, 1> o;
    o << obs[0], obs[1], obs[2];
    e.setMeasurement(o);
    e.fx = intr[0], e.fy = intr[1], e.cx = intr[2], e.cy = intr[3], e.bf = intr[4];
    e.computeError();
    for (int i = 0; i < 3; i++) err[i] = e.error()(i);
    D = 3;
    if (Jpt && Jkf) {
      g2o::JacobianWorkspace ws;
      ws.updateSize(&e);
      ws.allocate();
      e.g2o::BaseBinaryEdge<3, Eigen::Vector3d, g2o::VertexSBAPointXYZ, g2o::VertexSE3Expmap>::linearizeOplus(ws);
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) Jpt[3 * r + c] = e.jacobianOplusXi()(r, c);
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 6; c++) Jkf[6 * r + c] = e.jacobianOplusXj()(r, c);
    }
    e.setVertex(0, nullptr), e.setVertex(1, nullptr);
  }
  return D;
}
int ppo_ref_plane_edge(int kind, const double pl[4], const double pose[7], const double meas[4], double err[3]) {
  g2o::VertexPlane vp;
  vp.setEstimate(pl_in(pl));
  g2o::VertexSE3Expmap vk;
  vk.setEstimate(se3_in(pose));
  err[2] = 0;
  if (kind == PPO_PLANE_OBS) {
    g2o::EdgePlane e;
    e.setVertex(0, &vp), e.setVertex(1, &vk);
    e.setMeasurement(pl_in(meas));
    e.computeError();
    for (int i = 0; i < 3; i++) err[i] = e.error()(i);
    e.setVertex(0, nullptr), e.setVertex(1, nullptr);
    return 3;
  } else if (kind == PPO_PLANE_VER) {
    g2o::EdgeVerticalPlane e;
    e.setVertex(0, &vp), e.setVertex(1, &vk);
    e.setMeasurement(pl_in(meas));
    e.computeError();
    err[0] = e.error()(0), err[1] = e.error()(1);
    e.setVertex(0, nullptr), e.setVertex(1, nullptr);
    return 2;
  }
  g2o::EdgeParallelPlane e;
  e.setVertex(0, &vp), e.setVertex(1, &vk);
  e.setMeasurement(pl_in(meas));
  e.computeError();
  err[0] = e.error()(0), err[1] = e.error()(1);
  e.setVertex(0, nullptr), e.setVertex(1, nullptr);
  return 2;
}
int ppo_ref_cuboid_cam_edge(int kind, const double pose[7], const double c[10], const float intr[5], const double *meas, double err[16]) {
  g2o::VertexSE3Expmap vk;
  vk.setEstimate(se3_in(pose));
  g2o::VertexCuboid vc;
  vc.setEstimate(cu_in(c));
  if (kind == PPO_CUBOID_BBOX) {
    g2o::EdgeSE3CuboidProj e;
    e.setVertex(0, &vk), e.setVertex(1, &vc);
    e.Kalib = K_in(intr);
    e.setMeasurement(Eigen::Vector4d(meas[0], meas[1], meas[2], meas[3]));
    e.computeError();
    for (int i = 0; i < 4; i++) err[i] = e.error()(i);
    e.setVertex(0, nullptr), e.setVertex(1, nullptr);
    return 4;
  }
  if (kind == PPO_CUBOID_SE3) {
    g2o::EdgeSE3Cuboid e;
    e.setVertex(0, &vk), e.setVertex(1, &vc);
    e.setMeasurement(cu_in(meas));
    e.computeError();
    for (int i = 0; i < 9; i++) err[i] = e.error()(i);
    e.setVertex(0, nullptr), e.setVertex(1, nullptr);
    return 9;
  }
  g2o::EdgeSE3CuboidCornerProj e;
  e.setVertex(0, &vk), e.setVertex(1, &vc);
  e.Kalib = K_in(intr);
  Eigen::Matrix<double, 16, 1> m;
  for (int i = 0; i < 16; i++) m(i) = meas[i];
  e.setMeasurement(m);
  e.computeError();
  for (int i = 0; i < 16; i++) err[i] = e.error()(i);
  e.setVertex(0, nullptr), e.setVertex(1, nullptr);
  return 16;
}

}  // extern "C"
