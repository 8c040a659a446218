// TEST INFRASTRUCTURE — NOT PRODUCT CODE.  (see the header of ppo_oracle_math.h)
// CPU restatement of the reference's local bundle adjustment:
//   Optimizer::LocalBACameraPlaneCuboids   src/Optimizer.cc:1994-2967 (stages C-F)
//   Optimizer::LocalBundleAdjustment       src/Optimizer.cc:461-786
// on the flat graph of include/ppo_ba.h, i.e. g2o's SparseOptimizer + BlockSolver (Schur) +
// LinearSolverDense + OptimizationAlgorithmLevenberg with the reference's vertex / edge types.
// Single-threaded like the reference (G2O_OPENMP undefined, Thirdparty/g2o/config.h:4).
// PARITY UNPINNED by the reference's own tests (it has none): pinned by tests/test_oracle_kat.py.
//
// g2o file:line followed (relative to /root/reference/Thirdparty/g2o/g2o unless noted):
//   initializeOptimization / buildIndexMapping  core/sparse_optimizer.cpp:166-267
//   computeActiveErrors / activeRobustChi2      core/sparse_optimizer.cpp:61-114
//   optimize / update / push / pop              core/sparse_optimizer.cpp:354-435,600-613
//   buildSystem / setLambda / solve (Schur)     core/block_solver.hpp:354-604
//   LM iteration                                core/optimization_algorithm_levenberg.cpp:61-189
//   numeric Jacobians                           core/base_binary_edge.hpp:216-320, base_unary_edge.hpp:81-123
//   quadratic forms                             core/base_binary_edge.hpp:54-120, base_unary_edge.hpp:42-72
//   dense solve                                 solvers/linear_solver_dense.h:65-113 (Eigen::LDLT + isPositive)
#include <algorithm>
#include <chrono>
#include <cfloat>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/ppo_ba.h"
#include "ppo_oracle_math.h"

using namespace ppo_oracle;

namespace {

// ------------------------------------------------------------------------------------------------
// edge residuals
// ------------------------------------------------------------------------------------------------
// EdgeSE3ProjectXYZ::computeError      types/types_six_dof_expmap.h:174-179, .cpp:172-179
// EdgeStereoSE3ProjectXYZ::computeError types/types_six_dof_expmap.h:206-211, .cpp:182-189
//   (stereo: invz and bf are float)
inline int point_edge_error(const SE3 &T, const V3 &X, const float intr[5], const float obs[3], double err[3]) {
  V3 p = se3_map(T, X);
  const double fx = intr[0], fy = intr[1], cx = intr[2], cy = intr[3];
  if (obs[2] < 0) {
    double px = p[0] / p[2], py = p[1] / p[2];  // project2d
    err[0] = (double)obs[0] - (px * fx + cx);
    err[1] = (double)obs[1] - (py * fy + cy);
    err[2] = 0;
    return 2;
  }
  const double bfd = intr[4];
  const float bf = (float)bfd;
  const float invz = (float)(1.0f / p[2]);
  double r0 = p[0] * invz * fx + cx;
  double r1 = p[1] * invz * fy + cy;
  double r2 = r0 - (double)(bf * invz);
  err[0] = (double)obs[0] - r0;
  err[1] = (double)obs[1] - r1;
  err[2] = (double)obs[2] - r2;
  return 3;
}
// linearizeOplus: types_six_dof_expmap.cpp:135-170 (mono), :216-266 (stereo). Jpt d x 3, Jkf d x 6 row-major.
inline void point_edge_jacobian(const SE3 &T, const V3 &X, const float intr[5], bool stereo, double Jpt[9], double Jkf[18]) {
  V3 p = se3_map(T, X);
  M3 R = quat_to_matrix(T.r);
  const double fx = intr[0], fy = intr[1], bf = intr[4];
  double x = p[0], y = p[1], z = p[2], z_2 = z * z;
  if (!stereo) {
    double tmp[2][3] = {{fx, 0, -x / z * fx}, {0, fy, -y / z * fy}};
    for (int i = 0; i < 2; i++)
      for (int j = 0; j < 3; j++) {
        double a = 0;
        for (int k = 0; k < 3; k++) a += (-1. / z * tmp[i][k]) * R(k, j);
        Jpt[i * 3 + j] = a;
      }
  } else {
    for (int j = 0; j < 3; j++) {
      Jpt[0 * 3 + j] = -fx * R(0, j) / z + fx * x * R(2, j) / z_2;
      Jpt[1 * 3 + j] = -fy * R(1, j) / z + fy * y * R(2, j) / z_2;
      Jpt[2 * 3 + j] = Jpt[0 * 3 + j] - bf * R(2, j) / z_2;
    }
  }
  Jkf[0] = x * y / z_2 * fx;
  Jkf[1] = -(1 + (x * x / z_2)) * fx;
  Jkf[2] = y / z * fx;
  Jkf[3] = -1. / z * fx;
  Jkf[4] = 0;
  Jkf[5] = x / z_2 * fx;
  Jkf[6] = (1 + y * y / z_2) * fy;
  Jkf[7] = -x * y / z_2 * fy;
  Jkf[8] = -x / z * fy;
  Jkf[9] = 0;
  Jkf[10] = -1. / z * fy;
  Jkf[11] = y / z_2 * fy;
  if (stereo) {
    Jkf[12] = Jkf[0] - bf * y / z_2;
    Jkf[13] = Jkf[1] + bf * x / z_2;
    Jkf[14] = Jkf[2];
    Jkf[15] = Jkf[3];
    Jkf[16] = 0;
    Jkf[17] = Jkf[5] - bf / z_2;
  }
}
// EdgePlane / EdgeVerticalPlane / EdgeParallelPlane::computeError  include/G2O_Plane3D.h:181-193,220-232,279-291
inline int plane_edge_error(int kind, const Plane &pl, const SE3 &T, const Plane &meas, double err[3]) {
  Plane local = plane_transform(T, pl);
  err[2] = 0;
  if (kind == PPO_PLANE_OBS) {
    plane_ominus(local, meas, err);
    return 3;
  } else if (kind == PPO_PLANE_VER) {
    plane_ominus_ver(local, meas, err);
    return 2;
  }
  plane_ominus_par(local, meas, err);
  return 2;
}
inline void K_from_intr(const float intr[5], double K[9]) {
  // Optimizer.cc:2450-2454 copies the float KeyFrame::mK; mK = [fx 0 cx; 0 fy cy; 0 0 1] (Frame.cc:62)
  K[0] = intr[0]; K[1] = 0; K[2] = intr[2];
  K[3] = 0; K[4] = intr[1]; K[5] = intr[3];
  K[6] = 0; K[7] = 0; K[8] = 1;
}
// EdgeSE3CuboidProj::computeError src/g2o_cuboid.cc:70-80 ; EdgeSE3CuboidCornerProj :103-120
inline int cbe_dim(int kind) { return kind == PPO_CUBOID_BBOX ? 4 : (kind == PPO_CUBOID_SE3 ? 9 : 16); }
inline int cuboid_cam_error(int kind, const SE3 &T, const Cuboid &c, const double K[9], const double *meas, double err[16]) {
  if (kind == PPO_CUBOID_SE3) {  // EdgeSE3Cuboid::computeError include/g2o_cuboid.h:330-340; measurement laid out like a cuboid state
    Cuboid m;
    m.pose = se3_from_qt(Quat{meas[3], meas[4], meas[5], meas[6]}, v3(meas[0], meas[1], meas[2]));
    m.scale = v3(meas[7], meas[8], meas[9]);
    cuboid_se3_error(T, c, m, err);
    return 9;
  }
  if (kind == PPO_CUBOID_BBOX) {
    double b[4];
    cuboid_project_bbox(c, T, K, b);
    for (int i = 0; i < 4; i++) err[i] = b[i] - meas[i];
    return 4;
  }
  double p[2][8];
  cuboid_project(c, T, K, p);
  for (int i = 0; i < 8; i++) {
    err[2 * i] = p[0][i] - meas[2 * i];
    err[2 * i + 1] = p[1][i] - meas[2 * i + 1];
  }
  return 16;
}
// EdgePointCuboidOnlyObject::computeError src/g2o_cuboid.cc:132-160 (prior_object_half_size unset)
inline void point_cuboid_error(const Cuboid &c, const double *pts, int n, double ratio, double prior_weight, double err[3]) {
  V3 acc = v3(0, 0, 0);
  for (int i = 0; i < n; i++) {
    V3 e = cuboid_point_boundary_error(c, v3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]), ratio);
    acc = acc + v3(std::fabs(e[0]), std::fabs(e[1]), std::fabs(e[2]));
  }
  if (n > 0) acc = v3(acc[0] / n, acc[1] / n, acc[2] / n);
  for (int i = 0; i < 3; i++) err[i] = 1.0 * (acc[i] / c.scale[i]) + prior_weight * c.scale[i];
}

// general 3x3 inverse (Eigen's MatrixBase::inverse() for fixed size 3: cofactors / determinant)
inline void inv3(const double A[9], double R[9]) {
  double c00 = A[4] * A[8] - A[5] * A[7], c01 = A[5] * A[6] - A[3] * A[8], c02 = A[3] * A[7] - A[4] * A[6];
  double det = A[0] * c00 + A[1] * c01 + A[2] * c02;
  double id = 1.0 / det;
  R[0] = c00 * id;
  R[1] = (A[2] * A[7] - A[1] * A[8]) * id;
  R[2] = (A[1] * A[5] - A[2] * A[4]) * id;
  R[3] = c01 * id;
  R[4] = (A[0] * A[8] - A[2] * A[6]) * id;
  R[5] = (A[2] * A[3] - A[0] * A[5]) * id;
  R[6] = c02 * id;
  R[7] = (A[1] * A[6] - A[0] * A[7]) * id;
  R[8] = (A[0] * A[4] - A[1] * A[3]) * id;
}

struct HplBlock {
  int pose;      // block index among free KFs
  double m[18];  // 6 x 3 row-major  (d b_pose / d landmark)
};

}  // namespace

struct ppo_oracle_handle {
  ppo_ba_params P;
  // graph (owned copies)
  int n_kf = 0, n_pt = 0, n_pl = 0, n_cu = 0, n_pe = 0, n_ple = 0, n_cbe = 0, n_pce = 0, n_cpe = 0;
  std::vector<uint8_t> kf_fixed, pt_fixed, cu_flags;
  std::vector<float> kf_intr;
  std::vector<int> pt_rowptr, pe_kf, pe_pt;
  std::vector<float> pe_obs, pe_invsigma2;
  std::vector<int> ple_plane, ple_kf;
  std::vector<uint8_t> ple_kind;
  std::vector<Plane> ple_meas;
  std::vector<double> ple_info;
  std::vector<int> cbe_kf, cbe_cuboid;
  std::vector<uint8_t> cbe_kind;
  std::vector<double> cbe_meas, cbe_info;
  std::vector<int> pce_cuboid, pce_rowptr;
  std::vector<double> pce_pts;
  std::vector<int> cpe_cuboid, cpe_plane;
  std::vector<double> cpe_meas, cpe_info;
  // estimates
  std::vector<SE3> kf, kf0;
  std::vector<V3> pt, pt0;
  std::vector<Plane> pl, pl0;
  std::vector<Cuboid> cu, cu0;
  // edge flags and stored errors (g2o's Edge::_error after the last computeError of that edge)
  std::vector<uint8_t> ef[PPO_EDGE_KINDS];
  std::vector<double> pe_err, ple_err, cbe_err, pce_err, cpe_err;
  // index mapping of the current optimize()
  std::vector<int> kf_off, cu_off;  // scalar offset in the pose block, -1 if fixed / inactive
  std::vector<int> pl_l, pt_l;      // landmark index, -1 if inactive / fixed
  int n_p = 0, n_l = 0, n_kf_free = 0;
  // linear system
  std::vector<double> Hpp, Hschur, b, x, Hll, Dinv, bschur;
  std::vector<std::vector<HplBlock>> Hpl;  // per landmark, sorted by pose
  std::vector<int> pe_blk, ple_blk;        // edge -> block index inside Hpl[landmark], -1 if KF fixed
  double lambda = -1, ni = 2;
  int nBad = 0;

  int edge_count(int kind) const {
    switch (kind) {
      case PPO_EDGE_POINT: return n_pe;
      case PPO_EDGE_PLANE: return n_ple;
      case PPO_EDGE_CUBOID_CAM: return n_cbe;
      case PPO_EDGE_POINT_CUBOID: return n_pce;
      case PPO_EDGE_CUBOID_PLANE: return n_cpe;
    }
    return -1;
  }
  // Host threads (OpenMP).  1 = the reference's configuration (g2o is built without OpenMP, Thirdparty/g2o/config.h:4): every
  // loop then runs in the original order and all results are bit-reproducible.  > 1 is a reported variant only
  // ("not the reference configuration"): errors, point-edge linearisation, Schur rows, the dense factorisation's
  // row updates and the back-substitution run in parallel; partial sums are combined in thread order.
  int threads = 1;
  bool lvl0(int kind, int e) const { return !(ef[kind][e] & PPO_EF_LEVEL1); }
  bool robust(int kind, int e) const { return ef[kind][e] & PPO_EF_ROBUST; }
  bool pe_active(int e) const { return lvl0(PPO_EDGE_POINT, e) && !(kf_fixed[pe_kf[e]] && pt_fixed[pe_pt[e]]); }

  // ---- core/sparse_optimizer.cpp:199-267 + :166-190 -------------------------------------------
  void initialize_optimization() {
    std::vector<char> kf_act(n_kf, 0), cu_act(n_cu, 0), pl_act(n_pl, 0), pt_act(n_pt, 0);
    for (int e = 0; e < n_pe; e++)
      if (pe_active(e)) kf_act[pe_kf[e]] = 1, pt_act[pe_pt[e]] = 1;
    for (int e = 0; e < n_ple; e++)
      if (lvl0(PPO_EDGE_PLANE, e)) kf_act[ple_kf[e]] = 1, pl_act[ple_plane[e]] = 1;
    for (int e = 0; e < n_cbe; e++)
      if (lvl0(PPO_EDGE_CUBOID_CAM, e)) kf_act[cbe_kf[e]] = 1, cu_act[cbe_cuboid[e]] = 1;
    for (int e = 0; e < n_pce; e++)
      if (lvl0(PPO_EDGE_POINT_CUBOID, e)) cu_act[pce_cuboid[e]] = 1;
    for (int e = 0; e < n_cpe; e++)
      if (lvl0(PPO_EDGE_CUBOID_PLANE, e)) cu_act[cpe_cuboid[e]] = 1, pl_act[cpe_plane[e]] = 1;
    kf_off.assign(n_kf, -1);
    cu_off.assign(n_cu, -1);
    pl_l.assign(n_pl, -1);
    pt_l.assign(n_pt, -1);
    int off = 0;
    n_kf_free = 0;
    for (int i = 0; i < n_kf; i++)
      if (kf_act[i] && !kf_fixed[i]) kf_off[i] = off, off += 6, n_kf_free++;
    for (int i = 0; i < n_cu; i++)
      if (cu_act[i]) cu_off[i] = off, off += 9;
    n_p = off;
    int l = 0;
    for (int i = 0; i < n_pl; i++)
      if (pl_act[i]) pl_l[i] = l++;
    for (int i = 0; i < n_pt; i++)
      if (pt_act[i] && !pt_fixed[i]) pt_l[i] = l++;
    n_l = l;
  }
  // ---- BlockSolver::buildStructure core/block_solver.hpp:143-295 -------------------------------
  void build_structure() {
    Hpp.assign((size_t)n_p * n_p, 0.0);
    Hschur.assign((size_t)n_p * n_p, 0.0);
    b.assign(n_p + 3 * (size_t)n_l, 0.0);
    x.assign(n_p + 3 * (size_t)n_l, 0.0);
    bschur.assign(n_p, 0.0);
    Hll.assign(9 * (size_t)n_l, 0.0);
    Dinv.assign(9 * (size_t)n_l, 0.0);
    Hpl.assign(n_l, {});
    pe_blk.assign(n_pe, -1);
    ple_blk.assign(n_ple, -1);
    auto add_block = [&](int l, int pose) {
      auto &v = Hpl[l];
      for (auto &bk : v)
        if (bk.pose == pose) return;
      HplBlock nb;
      nb.pose = pose;
      std::memset(nb.m, 0, sizeof nb.m);
      v.push_back(nb);
    };
    for (int e = 0; e < n_pe; e++)
      if (pe_active(e) && pt_l[pe_pt[e]] >= 0 && kf_off[pe_kf[e]] >= 0) add_block(pt_l[pe_pt[e]], kf_off[pe_kf[e]] / 6);
    for (int e = 0; e < n_ple; e++)
      if (lvl0(PPO_EDGE_PLANE, e) && kf_off[ple_kf[e]] >= 0) add_block(pl_l[ple_plane[e]], kf_off[ple_kf[e]] / 6);
    for (auto &v : Hpl) std::sort(v.begin(), v.end(), [](const HplBlock &a, const HplBlock &c) { return a.pose < c.pose; });
    auto find_block = [&](int l, int pose) {
      auto &v = Hpl[l];
      for (size_t i = 0; i < v.size(); i++)
        if (v[i].pose == pose) return (int)i;
      return -1;
    };
    for (int e = 0; e < n_pe; e++)
      if (pe_active(e) && pt_l[pe_pt[e]] >= 0 && kf_off[pe_kf[e]] >= 0) pe_blk[e] = find_block(pt_l[pe_pt[e]], kf_off[pe_kf[e]] / 6);
    for (int e = 0; e < n_ple; e++)
      if (lvl0(PPO_EDGE_PLANE, e) && kf_off[ple_kf[e]] >= 0) ple_blk[e] = find_block(pl_l[ple_plane[e]], kf_off[ple_kf[e]] / 6);
  }

  // ---- residuals ------------------------------------------------------------------------------
  int pe_eval(int e, double err[3]) const { return point_edge_error(kf[pe_kf[e]], pt[pe_pt[e]], &kf_intr[5 * pe_kf[e]], &pe_obs[3 * e], err); }
  int ple_eval(int e, const Plane &p, const SE3 &T, double err[3]) const { return plane_edge_error(ple_kind[e], p, T, ple_meas[e], err); }
  int cbe_eval(int e, const SE3 &T, const Cuboid &c, double err[16]) const {
    double K[9];
    K_from_intr(&kf_intr[5 * cbe_kf[e]], K);
    return cuboid_cam_error(cbe_kind[e], T, c, K, &cbe_meas[16 * e], err);
  }
  void pce_eval(int e, const Cuboid &c, double err[3]) const {
    point_cuboid_error(c, &pce_pts[3 * pce_rowptr[e]], pce_rowptr[e + 1] - pce_rowptr[e], P.ptcu_max_outside_margin_ratio, P.ptcu_prior_weight, err);
  }
  double delta_of(int kind, int e) const {
    switch (kind) {
      case PPO_EDGE_POINT: return pe_obs[3 * e + 2] < 0 ? P.huber_mono : P.huber_stereo;
      case PPO_EDGE_PLANE: return ple_kind[e] == PPO_PLANE_OBS ? P.huber_plane : P.huber_vp_plane;
      case PPO_EDGE_CUBOID_CAM: return cbe_kind[e] == PPO_CUBOID_BBOX ? P.huber_bbox : (cbe_kind[e] == PPO_CUBOID_SE3 ? P.huber_se3 : P.huber_corner);
      case PPO_EDGE_CUBOID_PLANE: return P.huber_cuboid_plane;
    }
    return 0;
  }
  // e->chi2() from the stored _error  (core/base_edge.h: chi2 = _error . (information * _error))
  double chi2_of(int kind, int e) const {
    switch (kind) {
      case PPO_EDGE_POINT: {
        const double *r = &pe_err[3 * e];
        double s = pe_invsigma2[e];
        int d = pe_obs[3 * e + 2] < 0 ? 2 : 3;
        double c = 0;
        for (int i = 0; i < d; i++) c += r[i] * (s * r[i]);
        return c;
      }
      case PPO_EDGE_PLANE: {
        const double *r = &ple_err[3 * e];
        int d = ple_kind[e] == PPO_PLANE_OBS ? 3 : 2;
        double c = 0;
        for (int i = 0; i < d; i++) c += r[i] * (ple_info[3 * e + i] * r[i]);
        return c;
      }
      case PPO_EDGE_CUBOID_CAM: {
        const double *r = &cbe_err[16 * e];
        int d = cbe_dim(cbe_kind[e]);
        double c = 0;
        for (int i = 0; i < d; i++) c += r[i] * (cbe_info[e] * r[i]);
        return c;
      }
      case PPO_EDGE_POINT_CUBOID: {
        const double *r = &pce_err[3 * e];
        return r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
      }
      case PPO_EDGE_CUBOID_PLANE: {
        const double *r = &cpe_err[3 * e];
        double c = 0;
        for (int i = 0; i < 3; i++) c += r[i] * (cpe_info[3 * e + i] * r[i]);
        return c;
      }
    }
    return 0;
  }
  // core/sparse_optimizer.cpp:61-76
  void compute_active_errors() {
#pragma omp parallel for schedule(static) num_threads(threads) if (threads > 1)
    for (int e = 0; e < n_pe; e++)
      if (pe_active(e)) pe_eval(e, &pe_err[3 * e]);
    for (int e = 0; e < n_ple; e++)
      if (lvl0(PPO_EDGE_PLANE, e)) ple_eval(e, pl[ple_plane[e]], kf[ple_kf[e]], &ple_err[3 * e]);
    for (int e = 0; e < n_cbe; e++)
      if (lvl0(PPO_EDGE_CUBOID_CAM, e)) cbe_eval(e, kf[cbe_kf[e]], cu[cbe_cuboid[e]], &cbe_err[16 * e]);
    for (int e = 0; e < n_pce; e++)
      if (lvl0(PPO_EDGE_POINT_CUBOID, e)) pce_eval(e, cu[pce_cuboid[e]], &pce_err[3 * e]);
    for (int e = 0; e < n_cpe; e++)
      if (lvl0(PPO_EDGE_CUBOID_PLANE, e))
        for (int i = 0; i < 3; i++) cpe_err[3 * e + i] = cpe_meas[3 * e + i];  // G2O_Plane3D.h:470-473
  }
  // core/sparse_optimizer.cpp:100-114
  double active_robust_chi2() const {
    double chi = 0;
    for (int kind = 0; kind < PPO_EDGE_KINDS; kind++) {
      int n = edge_count(kind);
      for (int e = 0; e < n; e++) {
        bool act = kind == PPO_EDGE_POINT ? pe_active(e) : lvl0(kind, e);
        if (!act) continue;
        double c = chi2_of(kind, e);
        if (robust(kind, e)) {
          double rho[3];
          huber(c, delta_of(kind, e), rho);
          chi += rho[0];
        } else
          chi += c;
      }
    }
    return chi;
  }
  int n_active_edges() const {
    int n = 0;
    for (int kind = 0; kind < PPO_EDGE_KINDS; kind++)
      for (int e = 0; e < edge_count(kind); e++) n += kind == PPO_EDGE_POINT ? pe_active(e) : lvl0(kind, e);
    return n;
  }

  // ---- quadratic form helpers -------------------------------------------------------------------
  // H(off_i.., off_j..) += Ji^T W Jj with W = diag(w) ; row-major Jacobians (D x di), (D x dj)
  void add_pp(int oi, int di, const double *Ji, int oj, int dj, const double *Jj, int D, const double *w) {
    for (int a = 0; a < di; a++)
      for (int c = 0; c < dj; c++) {
        double s = 0;
        for (int r = 0; r < D; r++) s += Ji[r * di + a] * w[r] * Jj[r * dj + c];
        Hpp[(size_t)(oi + a) * n_p + (oj + c)] += s;
      }
  }
  void add_b(double *bv, int d, const double *J, int D, const double *w, const double *err) {
    for (int a = 0; a < d; a++) {
      double s = 0;
      for (int r = 0; r < D; r++) s += J[r * d + a] * (-(w[r] * err[r]));
      bv[a] += s;
    }
  }
  // threaded variant of the point-edge part of build_system: a thread owns whole points (their Hll / b_l / Hpl blocks are
  // private to the point); the pose-side 6x6 diagonal blocks and gradients go to per-thread buffers summed in thread order
  void build_system_points_mt() {
    const int T = threads;
    std::vector<std::vector<double>> Hk(T, std::vector<double>(36 * (size_t)n_kf, 0.0)), bk(T, std::vector<double>(6 * (size_t)n_kf, 0.0));
#pragma omp parallel num_threads(T)
    {
      int t = 0;
#ifdef _OPENMP
      t = omp_get_thread_num();
#endif
      double *Ht = Hk[t].data(), *bt = bk[t].data();
#pragma omp for schedule(static)
      for (int p = 0; p < n_pt; p++) {
        for (int e = pt_rowptr[p]; e < pt_rowptr[p + 1]; e++) {
          if (!pe_active(e)) continue;
          const int k = pe_kf[e];
          const bool stereo = pe_obs[3 * e + 2] >= 0;
          const int D = stereo ? 3 : 2;
          double Jpt[9], Jkf[18];
          point_edge_jacobian(kf[k], pt[p], &kf_intr[5 * k], stereo, Jpt, Jkf);
          double w[3], rho1 = 1.0;
          if (robust(PPO_EDGE_POINT, e)) {
            double rho[3];
            huber(chi2_of(PPO_EDGE_POINT, e), delta_of(PPO_EDGE_POINT, e), rho);
            rho1 = rho[1];
          }
          for (int i = 0; i < 3; i++) w[i] = rho1 * (double)pe_invsigma2[e];
          const double *err = &pe_err[3 * e];
          const int l = pt_l[p], ko = kf_off[k];
          if (l >= 0) {
            for (int a = 0; a < 3; a++)
              for (int c = 0; c < 3; c++) {
                double sacc = 0;
                for (int r = 0; r < D; r++) sacc += Jpt[r * 3 + a] * w[r] * Jpt[r * 3 + c];
                Hll[9 * (size_t)l + 3 * a + c] += sacc;
              }
            add_b(&b[n_p + 3 * (size_t)l], 3, Jpt, D, w, err);
          }
          if (ko >= 0) {
            for (int a = 0; a < 6; a++)
              for (int c = 0; c < 6; c++) {
                double sacc = 0;
                for (int r = 0; r < D; r++) sacc += Jkf[r * 6 + a] * w[r] * Jkf[r * 6 + c];
                Ht[36 * (size_t)k + 6 * a + c] += sacc;
              }
            add_b(&bt[6 * (size_t)k], 6, Jkf, D, w, err);
            if (l >= 0) {
              double *m = Hpl[l][pe_blk[e]].m;
              for (int a = 0; a < 6; a++)
                for (int c = 0; c < 3; c++) {
                  double sacc = 0;
                  for (int r = 0; r < D; r++) sacc += Jkf[r * 6 + a] * w[r] * Jpt[r * 3 + c];
                  m[a * 3 + c] += sacc;
                }
            }
          }
        }
      }
    }
    for (int t = 0; t < T; t++)
      for (int k = 0; k < n_kf; k++) {
        const int ko = kf_off[k];
        if (ko < 0) continue;
        for (int a = 0; a < 6; a++) {
          for (int c = 0; c < 6; c++) Hpp[(size_t)(ko + a) * n_p + (ko + c)] += Hk[t][36 * (size_t)k + 6 * a + c];
          b[ko + a] += bk[t][6 * (size_t)k + a];
        }
      }
  }
  // ---- BlockSolver::buildSystem core/block_solver.hpp:502-560 -----------------------------------
  void build_system() {
    std::fill(Hpp.begin(), Hpp.end(), 0.0);
    std::fill(b.begin(), b.end(), 0.0);
    std::fill(Hll.begin(), Hll.end(), 0.0);
    for (auto &v : Hpl)
      for (auto &bk : v) std::memset(bk.m, 0, sizeof bk.m);
    const double delta = 1e-9, scalar = 1.0 / (2 * delta);
    if (threads > 1) build_system_points_mt();
    // point edges: analytic Jacobians
    for (int e = 0; e < (threads > 1 ? 0 : n_pe); e++) {
      if (!pe_active(e)) continue;
      int k = pe_kf[e], p = pe_pt[e];
      bool stereo = pe_obs[3 * e + 2] >= 0;
      int D = stereo ? 3 : 2;
      double Jpt[9], Jkf[18];
      point_edge_jacobian(kf[k], pt[p], &kf_intr[5 * k], stereo, Jpt, Jkf);
      double w[3], rho1 = 1.0;
      if (robust(PPO_EDGE_POINT, e)) {
        double rho[3];
        huber(chi2_of(PPO_EDGE_POINT, e), delta_of(PPO_EDGE_POINT, e), rho);
        rho1 = rho[1];
      }
      for (int i = 0; i < 3; i++) w[i] = rho1 * (double)pe_invsigma2[e];
      const double *err = &pe_err[3 * e];
      int l = pt_l[p], ko = kf_off[k];
      if (l >= 0) {
        for (int a = 0; a < 3; a++)
          for (int c = 0; c < 3; c++) {
            double s = 0;
            for (int r = 0; r < D; r++) s += Jpt[r * 3 + a] * w[r] * Jpt[r * 3 + c];
            Hll[9 * (size_t)l + 3 * a + c] += s;
          }
        add_b(&b[n_p + 3 * (size_t)l], 3, Jpt, D, w, err);
      }
      if (ko >= 0) {
        add_pp(ko, 6, Jkf, ko, 6, Jkf, D, w);
        add_b(&b[ko], 6, Jkf, D, w, err);
        if (l >= 0) {
          double *m = Hpl[l][pe_blk[e]].m;
          for (int a = 0; a < 6; a++)
            for (int c = 0; c < 3; c++) {
              double s = 0;
              for (int r = 0; r < D; r++) s += Jkf[r * 6 + a] * w[r] * Jpt[r * 3 + c];
              m[a * 3 + c] += s;
            }
        }
      }
    }
    // plane edges: numeric Jacobians, vertex0 = plane (3), vertex1 = KF (6)
    for (int e = 0; e < n_ple; e++) {
      if (!lvl0(PPO_EDGE_PLANE, e)) continue;
      int k = ple_kf[e], p = ple_plane[e];
      int D = ple_kind[e] == PPO_PLANE_OBS ? 3 : 2;
      double Jpl[9] = {0}, Jkf[18] = {0};
      double ep[3], em[3];
      for (int d = 0; d < 3; d++) {
        double add[3] = {0, 0, 0};
        add[d] = delta;
        Plane pp = pl[p];
        plane_oplus(pp, add);
        ple_eval(e, pp, kf[k], ep);
        add[d] = -delta;
        Plane pm = pl[p];
        plane_oplus(pm, add);
        ple_eval(e, pm, kf[k], em);
        for (int r = 0; r < D; r++) Jpl[r * 3 + d] = scalar * (ep[r] - em[r]);
      }
      int ko = kf_off[k];
      if (!kf_fixed[k]) {
        for (int d = 0; d < 6; d++) {
          double add[6] = {0, 0, 0, 0, 0, 0};
          add[d] = delta;
          SE3 Tp = se3_oplus(kf[k], add);
          ple_eval(e, pl[p], Tp, ep);
          add[d] = -delta;
          SE3 Tm = se3_oplus(kf[k], add);
          ple_eval(e, pl[p], Tm, em);
          for (int r = 0; r < D; r++) Jkf[r * 6 + d] = scalar * (ep[r] - em[r]);
        }
      }
      double rho1 = 1.0;
      if (robust(PPO_EDGE_PLANE, e)) {
        double rho[3];
        huber(chi2_of(PPO_EDGE_PLANE, e), delta_of(PPO_EDGE_PLANE, e), rho);
        rho1 = rho[1];
      }
      double w[3];
      for (int i = 0; i < 3; i++) w[i] = rho1 * ple_info[3 * e + i];
      const double *err = &ple_err[3 * e];
      int l = pl_l[p];
      for (int a = 0; a < 3; a++)
        for (int c = 0; c < 3; c++) {
          double s = 0;
          for (int r = 0; r < D; r++) s += Jpl[r * 3 + a] * w[r] * Jpl[r * 3 + c];
          Hll[9 * (size_t)l + 3 * a + c] += s;
        }
      add_b(&b[n_p + 3 * (size_t)l], 3, Jpl, D, w, err);
      if (ko >= 0) {
        add_pp(ko, 6, Jkf, ko, 6, Jkf, D, w);
        add_b(&b[ko], 6, Jkf, D, w, err);
        double *m = Hpl[l][ple_blk[e]].m;
        for (int a = 0; a < 6; a++)
          for (int c = 0; c < 3; c++) {
            double s = 0;
            for (int r = 0; r < D; r++) s += Jkf[r * 6 + a] * w[r] * Jpl[r * 3 + c];
            m[a * 3 + c] += s;
          }
      }
    }
    // camera-cuboid edges: numeric, vertex0 = KF (6), vertex1 = cuboid (9)
    for (int e = 0; e < n_cbe; e++) {
      if (!lvl0(PPO_EDGE_CUBOID_CAM, e)) continue;
      int k = cbe_kf[e], c = cbe_cuboid[e];
      int D = cbe_dim(cbe_kind[e]);
      double Jkf[16 * 6] = {0}, Jcu[16 * 9] = {0}, ep[16], em[16];
      if (!kf_fixed[k]) {
        for (int d = 0; d < 6; d++) {
          double add[6] = {0, 0, 0, 0, 0, 0};
          add[d] = delta;
          SE3 Tp = se3_oplus(kf[k], add);
          cbe_eval(e, Tp, cu[c], ep);
          add[d] = -delta;
          SE3 Tm = se3_oplus(kf[k], add);
          cbe_eval(e, Tm, cu[c], em);
          for (int r = 0; r < D; r++) Jkf[r * 6 + d] = scalar * (ep[r] - em[r]);
        }
      }
      for (int d = 0; d < 9; d++) {
        double add[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        add[d] = delta;
        Cuboid cp = cuboid_oplus(cu[c], cu_flags[c], add);
        cbe_eval(e, kf[k], cp, ep);
        add[d] = -delta;
        Cuboid cm = cuboid_oplus(cu[c], cu_flags[c], add);
        cbe_eval(e, kf[k], cm, em);
        for (int r = 0; r < D; r++) Jcu[r * 9 + d] = scalar * (ep[r] - em[r]);
      }
      double rho1 = 1.0;
      if (robust(PPO_EDGE_CUBOID_CAM, e)) {
        double rho[3];
        huber(chi2_of(PPO_EDGE_CUBOID_CAM, e), delta_of(PPO_EDGE_CUBOID_CAM, e), rho);
        rho1 = rho[1];
      }
      double w[16];
      for (int i = 0; i < 16; i++) w[i] = rho1 * cbe_info[e];
      const double *err = &cbe_err[16 * e];
      int ko = kf_off[k], co = cu_off[c];
      add_pp(co, 9, Jcu, co, 9, Jcu, D, w);
      add_b(&b[co], 9, Jcu, D, w, err);
      if (ko >= 0) {
        add_pp(ko, 6, Jkf, ko, 6, Jkf, D, w);
        add_b(&b[ko], 6, Jkf, D, w, err);
        add_pp(ko, 6, Jkf, co, 9, Jcu, D, w);  // upper block (KF index < cuboid index)
      }
    }
    // point-cuboid unary edges: numeric, no robust kernel (Optimizer.cc:2640-2653)
    for (int e = 0; e < n_pce; e++) {
      if (!lvl0(PPO_EDGE_POINT_CUBOID, e)) continue;
      int c = pce_cuboid[e];
      double J[27], ep[3], em[3];
      for (int d = 0; d < 9; d++) {
        double add[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        add[d] = delta;
        Cuboid cp = cuboid_oplus(cu[c], cu_flags[c], add);
        pce_eval(e, cp, ep);
        add[d] = -delta;
        Cuboid cm = cuboid_oplus(cu[c], cu_flags[c], add);
        pce_eval(e, cm, em);
        for (int r = 0; r < 3; r++) J[r * 9 + d] = scalar * (ep[r] - em[r]);
      }
      double rho1 = 1.0;
      if (robust(PPO_EDGE_POINT_CUBOID, e)) {  // never set by the reference; kept general
        double rho[3];
        huber(chi2_of(PPO_EDGE_POINT_CUBOID, e), 1.0, rho);
        rho1 = rho[1];
      }
      double w[3] = {rho1, rho1, rho1};
      int co = cu_off[c];
      add_pp(co, 9, J, co, 9, J, 3, w);
      add_b(&b[co], 9, J, 3, w, &pce_err[3 * e]);
    }
    // cuboid-plane edges: constant residual => numeric Jacobians are exactly 0: nothing to add.
  }

  // ---- dense LDLT (solvers/linear_solver_dense.h:65-113: Eigen::LDLT, fail if !isPositive) -------
  // Unpivoted LDL^T on the mirrored upper triangle; for SPD input it is the same decomposition
  // up to rounding as Eigen's pivoted one.  Returns false if a pivot is <= 0.
  // PPO_SOLVER_6_3 = LinearSolverEigen (solvers/linear_solver_eigen.h:94-124): Eigen::SimplicialLDLT, an LDL^T without pivoting
  // whose numeric factorisation only fails on a pivot that is exactly zero -- an indefinite system is solved, not rejected.  (Its
  // fill-reducing ordering changes the rounding of the result, not the result.)
  bool dense_solve(std::vector<double> &A, const double *rhs, double *sol) {
    const bool eigen_flavour = P.solver == PPO_SOLVER_6_3;
    const int n = n_p;
    // work on lower triangle, row-major: L(i,j) j<i stored in A[i*n+j]
    for (int i = 0; i < n; i++)
      for (int j = 0; j < i; j++) A[(size_t)i * n + j] = A[(size_t)j * n + i];
    std::vector<double> d(n), tmp(n);
    for (int j = 0; j < n; j++) {
      double *Aj = &A[(size_t)j * n];
      double dj = Aj[j];
      for (int k = 0; k < j; k++) {
        tmp[k] = Aj[k] * d[k];
        dj -= Aj[k] * tmp[k];
      }
      if (eigen_flavour ? dj == 0.0 : !(dj > 0.0)) return false;
      d[j] = dj;
      double inv = 1.0 / dj;
#pragma omp parallel for schedule(static) num_threads(threads) if (threads > 1 && n - j > 64)
      for (int i = j + 1; i < n; i++) {
        double *Ai = &A[(size_t)i * n];
        double s = Ai[j];
        for (int k = 0; k < j; k++) s -= Ai[k] * tmp[k];
        Ai[j] = s * inv;
      }
    }
    for (int i = 0; i < n; i++) {
      double s = rhs[i];
      const double *Ai = &A[(size_t)i * n];
      for (int k = 0; k < i; k++) s -= Ai[k] * sol[k];
      sol[i] = s;
    }
    for (int i = 0; i < n; i++) sol[i] /= d[i];
    for (int i = n - 1; i >= 0; i--) {
      double s = sol[i];
      for (int k = i + 1; k < n; k++) s -= A[(size_t)k * n + i] * sol[k];
      sol[i] = s;
    }
    return true;
  }
  // threaded Schur complement: Dinv per landmark in parallel, then every thread owns the pose block rows i1 with
  // (pose index % T == t) and scans all landmarks for them: no two threads write the same entry, no reduction needed
  void schur_mt(double lam, std::vector<double> &coeff) {
    const int T = threads;
#pragma omp parallel for schedule(static) num_threads(T)
    for (int l = 0; l < n_l; l++) {
      double D[9];
      for (int i = 0; i < 9; i++) D[i] = Hll[9 * (size_t)l + i];
      D[0] += lam;
      D[4] += lam;
      D[8] += lam;
      inv3(D, &Dinv[9 * (size_t)l]);
    }
#pragma omp parallel num_threads(T)
    {
      int t = 0;
#ifdef _OPENMP
      t = omp_get_thread_num();
#endif
      for (int l = 0; l < n_l; l++) {
        const double *Di = &Dinv[9 * (size_t)l];
        const double *bl = &b[n_p + 3 * (size_t)l];
        auto &blocks = Hpl[l];
        for (size_t i1 = 0; i1 < blocks.size(); i1++) {
          if (blocks[i1].pose % T != t) continue;
          double db[3];
          for (int i = 0; i < 3; i++) db[i] = Di[3 * i] * bl[0] + Di[3 * i + 1] * bl[1] + Di[3 * i + 2] * bl[2];
          const double *Bi = blocks[i1].m;
          double BDinv[18];
          for (int a = 0; a < 6; a++)
            for (int c = 0; c < 3; c++) BDinv[a * 3 + c] = Bi[a * 3] * Di[c] + Bi[a * 3 + 1] * Di[3 + c] + Bi[a * 3 + 2] * Di[6 + c];
          const int o1 = blocks[i1].pose * 6;
          for (int a = 0; a < 6; a++) coeff[o1 + a] += Bi[a * 3] * db[0] + Bi[a * 3 + 1] * db[1] + Bi[a * 3 + 2] * db[2];
          for (size_t i2 = i1; i2 < blocks.size(); i2++) {
            const double *Bj = blocks[i2].m;
            const int o2 = blocks[i2].pose * 6;
            for (int a = 0; a < 6; a++)
              for (int c = 0; c < 6; c++)
                Hschur[(size_t)(o1 + a) * n_p + (o2 + c)] -= BDinv[a * 3] * Bj[c * 3] + BDinv[a * 3 + 1] * Bj[c * 3 + 1] + BDinv[a * 3 + 2] * Bj[c * 3 + 2];
          }
        }
      }
    }
  }
  // ---- setLambda + BlockSolver::solve + restoreDiagonal (core/block_solver.hpp:354-486,564-604) --
  bool schur_only = false;
  bool solve_damped(double lam) {
    Hschur = Hpp;
    for (int i = 0; i < n_p; i++) Hschur[(size_t)i * n_p + i] += lam;
    std::vector<double> coeff(n_p, 0.0);
    if (threads > 1) schur_mt(lam, coeff);
    for (int l = 0; l < (threads > 1 ? 0 : n_l); l++) {
      double D[9];
      for (int i = 0; i < 9; i++) D[i] = Hll[9 * (size_t)l + i];
      D[0] += lam;
      D[4] += lam;
      D[8] += lam;
      double *Di = &Dinv[9 * (size_t)l];
      inv3(D, Di);
      const double *bl = &b[n_p + 3 * (size_t)l];
      double db[3];
      for (int i = 0; i < 3; i++) db[i] = Di[3 * i] * bl[0] + Di[3 * i + 1] * bl[1] + Di[3 * i + 2] * bl[2];
      auto &blocks = Hpl[l];
      for (size_t i1 = 0; i1 < blocks.size(); i1++) {
        const double *Bi = blocks[i1].m;
        double BDinv[18];
        for (int a = 0; a < 6; a++)
          for (int c = 0; c < 3; c++) BDinv[a * 3 + c] = Bi[a * 3] * Di[c] + Bi[a * 3 + 1] * Di[3 + c] + Bi[a * 3 + 2] * Di[6 + c];
        int o1 = blocks[i1].pose * 6;
        for (int a = 0; a < 6; a++) coeff[o1 + a] += Bi[a * 3] * db[0] + Bi[a * 3 + 1] * db[1] + Bi[a * 3 + 2] * db[2];
        for (size_t i2 = i1; i2 < blocks.size(); i2++) {
          const double *Bj = blocks[i2].m;
          int o2 = blocks[i2].pose * 6;
          for (int a = 0; a < 6; a++)
            for (int c = 0; c < 6; c++)
              Hschur[(size_t)(o1 + a) * n_p + (o2 + c)] -= BDinv[a * 3] * Bj[c * 3] + BDinv[a * 3 + 1] * Bj[c * 3 + 1] + BDinv[a * 3 + 2] * Bj[c * 3 + 2];
        }
      }
    }
    for (int i = 0; i < n_p; i++) bschur[i] = b[i] - coeff[i];
    if (schur_only) return true;  // (tests on very large windows check the reduced system itself; its solution is checked against LAPACK)
    if (n_p > 0 && !dense_solve(Hschur, bschur.data(), x.data())) return false;
    // xl = Dinv (bl - Hpl^T xp)
#pragma omp parallel for schedule(static) num_threads(threads) if (threads > 1)
    for (int l = 0; l < n_l; l++) {
      double cl[3] = {b[n_p + 3 * (size_t)l], b[n_p + 3 * (size_t)l + 1], b[n_p + 3 * (size_t)l + 2]};
      for (auto &bk : Hpl[l]) {
        const double *xp = &x[bk.pose * 6];
        for (int c = 0; c < 3; c++)
          for (int a = 0; a < 6; a++) cl[c] -= bk.m[a * 3 + c] * xp[a];
      }
      const double *Di = &Dinv[9 * (size_t)l];
      for (int i = 0; i < 3; i++) x[n_p + 3 * (size_t)l + i] = Di[3 * i] * cl[0] + Di[3 * i + 1] * cl[1] + Di[3 * i + 2] * cl[2];
    }
    return true;
  }
  // SparseOptimizer::update core/sparse_optimizer.cpp:422-435
  void apply_update() {
    for (int i = 0; i < n_kf; i++)
      if (kf_off[i] >= 0) kf[i] = se3_oplus(kf[i], &x[kf_off[i]]);
    for (int i = 0; i < n_cu; i++)
      if (cu_off[i] >= 0) cu[i] = cuboid_oplus(cu[i], cu_flags[i], &x[cu_off[i]]);
    for (int i = 0; i < n_pl; i++)
      if (pl_l[i] >= 0) plane_oplus(pl[i], &x[n_p + 3 * (size_t)pl_l[i]]);
    for (int i = 0; i < n_pt; i++)
      if (pt_l[i] >= 0) {
        const double *d = &x[n_p + 3 * (size_t)pt_l[i]];
        pt[i] = v3(pt[i][0] + d[0], pt[i][1] + d[1], pt[i][2] + d[2]);  // types_sba.h:51-55
      }
  }
  // computeLambdaInit levenberg.cpp:166-180
  double lambda_init() const {
    double mx = 0;
    for (int i = 0; i < n_p; i++) mx = std::max(std::fabs(Hpp[(size_t)i * n_p + i]), mx);
    for (int l = 0; l < n_l; l++)
      for (int j = 0; j < 3; j++) mx = std::max(std::fabs(Hll[9 * (size_t)l + 4 * j]), mx);
    return P.lm_tau * mx;
  }
  // computeScale levenberg.cpp:182-189
  double compute_scale() const {
    double s = 0;
    for (size_t j = 0; j < x.size(); j++) s += x[j] * (lambda * x[j] + b[j]);
    return s;
  }

  // SparseOptimizer::optimize + OptimizationAlgorithmLevenberg::solve
  int optimize(int iters, const volatile unsigned char *stop, ppo_ba_stats *st) {
    auto t0 = std::chrono::steady_clock::now();
    if (st) std::memset(st, 0, sizeof *st);
    initialize_optimization();
    if (n_p + n_l == 0) return PPO_E_EMPTY;
    auto terminate = [&]() { return stop && *stop; };
    int done = 0;
    bool ok = true;
    int term = 0;
    for (int it = 0; it < iters && !terminate() && ok; it++) {
      if (it == 0) build_structure();
      compute_active_errors();
      double currentChi = active_robust_chi2();
      double tempChi = currentChi;
      double iniChi = currentChi;
      if (st && it == 0) st->chi2_initial = currentChi;
      build_system();
      if (it == 0) {
        lambda = lambda_init();
        ni = 2;
        nBad = 0;
      }
      double rho = 0;
      int qmax = 0;
      bool accepted = false;
      do {
        std::vector<SE3> kf_bak = kf;  // push
        std::vector<V3> pt_bak = pt;
        std::vector<Plane> pl_bak = pl;
        std::vector<Cuboid> cu_bak = cu;
        bool ok2 = solve_damped(lambda);
        apply_update();  // g2o applies x even if the solve failed; x is then stale — we skip nothing
        compute_active_errors();
        tempChi = active_robust_chi2();
        if (!ok2) tempChi = std::numeric_limits<double>::max();
        rho = (currentChi - tempChi);
        double scale = compute_scale();
        scale += 1e-3;
        rho /= scale;
        if (rho > 0 && std::isfinite(tempChi)) {
          double alpha = 1. - std::pow((2 * rho - 1), 3);
          alpha = std::min(alpha, P.lm_good_upper);
          double scaleFactor = std::max(P.lm_good_lower, alpha);
          lambda *= scaleFactor;
          ni = 2;
          currentChi = tempChi;
          accepted = true;  // discardTop
        } else {
          lambda *= ni;
          ni *= 2;
          kf = kf_bak;  // pop
          pt = pt_bak;
          pl = pl_bak;
          cu = cu_bak;
          accepted = false;
        }
        qmax++;
      } while (rho < 0 && qmax < P.lm_max_trials && !terminate());
      done++;
      if (st) {
        st->total_trials += qmax;
        if (it < PPO_TRACE_MAX) {
          ppo_ba_iter &r = st->trace[it];
          r.chi2_before = iniChi;
          r.chi2_after = currentChi;
          r.lambda = lambda;
          r.rho = rho;
          r.trials = qmax;
          r.accepted = accepted;
        }
        st->chi2_final = currentChi;
      }
      if (qmax == P.lm_max_trials || rho == 0) {
        ok = false;
        term = 1;
      } else {
        if ((iniChi - currentChi) * 1e3 < iniChi)
          nBad++;
        else
          nBad = 0;
        if (nBad >= 3) {
          ok = false;
          term = 1;
        }
      }
    }
    if (st) {
      st->iterations = done;
      st->terminated = term ? term : (terminate() ? 2 : 0);
      st->n_pose_dim = n_p;
      st->n_landmarks = n_l;
      st->n_active_edges = n_active_edges();
      st->ms_total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    return PPO_OK;
  }

  bool depth_positive_pe(int e) const { return se3_map(kf[pe_kf[e]], pt[pe_pt[e]])[2] > 0.0; }

  // Optimizer.cc:2736-2833
  void outlier_pass(int32_t n_out[3]) {
    n_out[0] = n_out[1] = n_out[2] = 0;
    for (int e = 0; e < n_pe; e++) {
      bool mono = pe_obs[3 * e + 2] < 0;
      if (chi2_of(PPO_EDGE_POINT, e) > (mono ? P.chi2_mono : P.chi2_stereo) || !depth_positive_pe(e)) {
        if (lvl0(PPO_EDGE_POINT, e)) n_out[0]++;
        ef[PPO_EDGE_POINT][e] |= PPO_EF_LEVEL1;
      }
      ef[PPO_EDGE_POINT][e] &= ~PPO_EF_ROBUST;
    }
    for (int e = 0; e < n_cbe; e++) {
      int D = cbe_dim(cbe_kind[e]);
      double s = 0;
      for (int i = 0; i < D; i++) s += cbe_err[16 * e + i] * cbe_err[16 * e + i];
      if (std::sqrt(s) > (cbe_kind[e] == PPO_CUBOID_BBOX ? P.norm_bbox : (cbe_kind[e] == PPO_CUBOID_SE3 ? P.norm_se3 : P.norm_corner))) {  // SE3: Optimizer.cc:1875-1882
        if (lvl0(PPO_EDGE_CUBOID_CAM, e)) n_out[2]++;
        ef[PPO_EDGE_CUBOID_CAM][e] |= PPO_EF_LEVEL1;
      }
    }
    for (int e = 0; e < n_ple; e++) {
      if (chi2_of(PPO_EDGE_PLANE, e) > (ple_kind[e] == PPO_PLANE_OBS ? P.chi2_plane : P.chi2_vp_plane)) {
        if (lvl0(PPO_EDGE_PLANE, e)) n_out[1]++;
        ef[PPO_EDGE_PLANE][e] |= PPO_EF_LEVEL1;
      }
      ef[PPO_EDGE_PLANE][e] &= ~PPO_EF_ROBUST;
    }
  }
};

// ------------------------------------------------------------------------------------------------
// C interface (mirrors include/ppo_ba.h so that tests read symmetrically)
// ------------------------------------------------------------------------------------------------
extern "C" {

void ppo_oracle_default_params(ppo_ba_params *p) {
  std::memset(p, 0, sizeof *p);
  p->huber_mono = (double)(float)std::sqrt(5.991);
  p->huber_stereo = (double)(float)std::sqrt(7.815);
  p->huber_plane = (double)(float)std::sqrt(500.0);
  p->huber_vp_plane = (double)(float)std::sqrt(200.0);
  p->huber_bbox = (double)(float)std::sqrt(80.0);
  p->huber_corner = (double)(float)std::sqrt(10.0);
  p->huber_cuboid_plane = (double)(float)std::sqrt(500.0);
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

int ppo_oracle_create(const ppo_ba_params *params, ppo_oracle_handle **out) {
  if (!params || !out) return PPO_E_INVALID;
  auto *h = new ppo_oracle_handle();
  h->P = *params;
  *out = h;
  return PPO_OK;
}
void ppo_oracle_destroy(ppo_oracle_handle *h) { delete h; }
int ppo_oracle_set_params(ppo_oracle_handle *h, const ppo_ba_params *params) {
  if (!h || !params) return PPO_E_INVALID;
  h->P = *params;
  return PPO_OK;
}

int ppo_oracle_set_graph(ppo_oracle_handle *h, const ppo_ba_graph *g) {
  if (!h || !g) return PPO_E_INVALID;
  h->n_kf = g->n_kf; h->n_pt = g->n_pt; h->n_pl = g->n_pl; h->n_cu = g->n_cu;
  h->n_pe = g->n_pe; h->n_ple = g->n_ple; h->n_cbe = g->n_cbe; h->n_pce = g->n_pce; h->n_cpe = g->n_cpe;
  h->kf.resize(g->n_kf);
  for (int i = 0; i < g->n_kf; i++) {
    const double *p = &g->kf_pose[7 * i];
    h->kf[i] = se3_from_qt(Quat{p[0], p[1], p[2], p[3]}, v3(p[4], p[5], p[6]));
  }
  h->kf_fixed.assign(g->kf_fixed, g->kf_fixed + g->n_kf);
  h->kf_intr.assign(g->kf_intr, g->kf_intr + 5 * (size_t)g->n_kf);
  h->pt.resize(g->n_pt);
  for (int i = 0; i < g->n_pt; i++) h->pt[i] = v3(g->pt_xyz[3 * i], g->pt_xyz[3 * i + 1], g->pt_xyz[3 * i + 2]);
  if (g->pt_fixed) h->pt_fixed.assign(g->pt_fixed, g->pt_fixed + g->n_pt);
  else h->pt_fixed.assign(g->n_pt, 0);
  h->pl.resize(g->n_pl);
  for (int i = 0; i < g->n_pl; i++) h->pl[i] = plane_from_vector(&g->pl_coef[4 * i]);
  h->cu.resize(g->n_cu);
  for (int i = 0; i < g->n_cu; i++) {
    const double *c = &g->cu_state[10 * i];
    h->cu[i].pose = se3_from_qt(Quat{c[3], c[4], c[5], c[6]}, v3(c[0], c[1], c[2]));
    h->cu[i].scale = v3(c[7], c[8], c[9]);
  }
  if (g->n_cu) h->cu_flags.assign(g->cu_flags, g->cu_flags + g->n_cu);
  else h->cu_flags.clear();
  h->pt_rowptr.assign(g->pt_rowptr, g->pt_rowptr + g->n_pt + 1);
  if (h->pt_rowptr[g->n_pt] != g->n_pe) return PPO_E_INVALID;
  h->pe_kf.assign(g->pe_kf, g->pe_kf + g->n_pe);
  h->pe_pt.resize(g->n_pe);
  for (int p = 0; p < g->n_pt; p++)
    for (int e = h->pt_rowptr[p]; e < h->pt_rowptr[p + 1]; e++) h->pe_pt[e] = p;
  h->pe_obs.assign(g->pe_obs, g->pe_obs + 3 * (size_t)g->n_pe);
  h->pe_invsigma2.assign(g->pe_invsigma2, g->pe_invsigma2 + g->n_pe);
  h->ple_plane.assign(g->ple_plane, g->ple_plane + g->n_ple);
  h->ple_kf.assign(g->ple_kf, g->ple_kf + g->n_ple);
  h->ple_kind.assign(g->ple_kind, g->ple_kind + g->n_ple);
  h->ple_meas.resize(g->n_ple);
  for (int e = 0; e < g->n_ple; e++) h->ple_meas[e] = plane_from_vector(&g->ple_meas[4 * e]);
  h->ple_info.assign(g->ple_info, g->ple_info + 3 * (size_t)g->n_ple);
  h->cbe_kf.assign(g->cbe_kf, g->cbe_kf + g->n_cbe);
  h->cbe_cuboid.assign(g->cbe_cuboid, g->cbe_cuboid + g->n_cbe);
  h->cbe_kind.assign(g->cbe_kind, g->cbe_kind + g->n_cbe);
  h->cbe_meas.assign(g->cbe_meas, g->cbe_meas + 16 * (size_t)g->n_cbe);
  h->cbe_info.assign(g->cbe_info, g->cbe_info + g->n_cbe);
  h->pce_cuboid.assign(g->pce_cuboid, g->pce_cuboid + g->n_pce);
  if (g->n_pce) {
    h->pce_rowptr.assign(g->pce_rowptr, g->pce_rowptr + g->n_pce + 1);
    h->pce_pts.assign(g->pce_pts, g->pce_pts + 3 * (size_t)h->pce_rowptr[g->n_pce]);
  } else {
    h->pce_rowptr.assign(1, 0);
    h->pce_pts.clear();
  }
  h->cpe_cuboid.assign(g->cpe_cuboid, g->cpe_cuboid + g->n_cpe);
  h->cpe_plane.assign(g->cpe_plane, g->cpe_plane + g->n_cpe);
  h->cpe_meas.assign(g->cpe_meas, g->cpe_meas + 3 * (size_t)g->n_cpe);
  h->cpe_info.assign(g->cpe_info, g->cpe_info + 3 * (size_t)g->n_cpe);
  for (int e = 0; e < g->n_pe; e++)
    if (h->pe_kf[e] < 0 || h->pe_kf[e] >= g->n_kf) return PPO_E_INVALID;
  h->ef[PPO_EDGE_POINT].assign(g->n_pe, PPO_EF_ROBUST);
  h->ef[PPO_EDGE_PLANE].assign(g->n_ple, PPO_EF_ROBUST);
  h->ef[PPO_EDGE_CUBOID_CAM].assign(g->n_cbe, PPO_EF_ROBUST);
  h->ef[PPO_EDGE_POINT_CUBOID].assign(g->n_pce, 0);
  h->ef[PPO_EDGE_CUBOID_PLANE].assign(g->n_cpe, PPO_EF_ROBUST);
  h->pe_err.assign(3 * (size_t)g->n_pe, 0.0);
  h->ple_err.assign(3 * (size_t)g->n_ple, 0.0);
  h->cbe_err.assign(16 * (size_t)g->n_cbe, 0.0);
  h->pce_err.assign(3 * (size_t)g->n_pce, 0.0);
  h->cpe_err.assign(3 * (size_t)g->n_cpe, 0.0);
  h->kf0 = h->kf; h->pt0 = h->pt; h->pl0 = h->pl; h->cu0 = h->cu;
  return PPO_OK;
}

int ppo_oracle_reset(ppo_oracle_handle *h) {
  h->kf = h->kf0; h->pt = h->pt0; h->pl = h->pl0; h->cu = h->cu0;
  std::fill(h->ef[PPO_EDGE_POINT].begin(), h->ef[PPO_EDGE_POINT].end(), PPO_EF_ROBUST);
  std::fill(h->ef[PPO_EDGE_PLANE].begin(), h->ef[PPO_EDGE_PLANE].end(), PPO_EF_ROBUST);
  std::fill(h->ef[PPO_EDGE_CUBOID_CAM].begin(), h->ef[PPO_EDGE_CUBOID_CAM].end(), PPO_EF_ROBUST);
  std::fill(h->ef[PPO_EDGE_POINT_CUBOID].begin(), h->ef[PPO_EDGE_POINT_CUBOID].end(), 0);
  std::fill(h->ef[PPO_EDGE_CUBOID_PLANE].begin(), h->ef[PPO_EDGE_CUBOID_PLANE].end(), PPO_EF_ROBUST);
  return PPO_OK;
}

int ppo_oracle_optimize(ppo_oracle_handle *h, int iters, const volatile unsigned char *stop, ppo_ba_stats *st) {
  return h->optimize(iters, stop, st);
}

int ppo_oracle_edge_count(const ppo_oracle_handle *h, int kind) { return h->edge_count(kind); }

int ppo_oracle_edge_chi2(ppo_oracle_handle *h, int kind, double *chi2, unsigned char *depth_positive, double *err_norm) {
  int n = h->edge_count(kind);
  if (n < 0) return PPO_E_INVALID;
  for (int e = 0; e < n; e++) {
    if (chi2) chi2[e] = h->chi2_of(kind, e);
    if (depth_positive) {
      if (kind == PPO_EDGE_POINT) depth_positive[e] = h->depth_positive_pe(e);
      else if (kind == PPO_EDGE_PLANE)  // EdgePlane::isDepthPositive G2O_Plane3D.h:199-209
        depth_positive[e] = plane_distance(plane_transform(h->kf[h->ple_kf[e]], h->pl[h->ple_plane[e]])) > 0;
      else depth_positive[e] = 1;
    }
    if (err_norm) {
      const double *r; int D;
      switch (kind) {
        case PPO_EDGE_POINT: r = &h->pe_err[3 * e]; D = h->pe_obs[3 * e + 2] < 0 ? 2 : 3; break;
        case PPO_EDGE_PLANE: r = &h->ple_err[3 * e]; D = h->ple_kind[e] == PPO_PLANE_OBS ? 3 : 2; break;
        case PPO_EDGE_CUBOID_CAM: r = &h->cbe_err[16 * e]; D = cbe_dim(h->cbe_kind[e]); break;
        case PPO_EDGE_POINT_CUBOID: r = &h->pce_err[3 * e]; D = 3; break;
        default: r = &h->cpe_err[3 * e]; D = 3; break;
      }
      double s = 0;
      for (int i = 0; i < D; i++) s += r[i] * r[i];
      err_norm[e] = std::sqrt(s);
    }
  }
  return PPO_OK;
}

// host threads of this handle (1 = reference configuration, the default)
int ppo_oracle_set_threads(ppo_oracle_handle *h, int n) {
  if (!h || n < 1) return PPO_E_INVALID;
  h->threads = n;
  return PPO_OK;
}

// e->computeError() on the level-1 point edges (Optimizer.cc:400-403,431-434)
int ppo_oracle_recompute_edge_errors(ppo_oracle_handle *h, int kind) {
  if (!h || kind != PPO_EDGE_POINT) return PPO_E_INVALID;
  for (int e = 0; e < h->n_pe; e++)
    if (!h->lvl0(PPO_EDGE_POINT, e)) h->pe_eval(e, &h->pe_err[3 * e]);
  return PPO_OK;
}

int ppo_oracle_set_edge_flags(ppo_oracle_handle *h, int kind, const unsigned char *flags) {
  int n = h->edge_count(kind);
  if (n < 0) return PPO_E_INVALID;
  h->ef[kind].assign(flags, flags + n);
  return PPO_OK;
}
int ppo_oracle_get_edge_flags(ppo_oracle_handle *h, int kind, unsigned char *flags) {
  int n = h->edge_count(kind);
  if (n < 0) return PPO_E_INVALID;
  std::copy(h->ef[kind].begin(), h->ef[kind].end(), flags);
  return PPO_OK;
}

int ppo_oracle_outlier_pass(ppo_oracle_handle *h, int32_t n_out[3]) {
  h->outlier_pass(n_out);
  return PPO_OK;
}

// stages C-E: Optimizer.cc:2723-2837
int ppo_oracle_local_ba(ppo_oracle_handle *h, const volatile unsigned char *stop, ppo_ba_result *res) {
  std::memset(res, 0, sizeof *res);
  if (stop && *stop) {
    res->skipped = 1;
    return PPO_OK;
  }
  int rc = h->optimize(h->P.iters_round1, stop, &res->round1);
  if (rc != PPO_OK) return rc;
  bool more = !(stop && *stop);
  if (more) {
    int32_t n_out[3];
    h->outlier_pass(n_out);
    res->n_outlier_point_edges = n_out[0];
    res->n_outlier_plane_edges = n_out[1];
    res->n_outlier_cuboid_edges = n_out[2];
    rc = h->optimize(h->P.iters_round2, stop, &res->round2);
    if (rc == PPO_E_EMPTY) rc = PPO_OK;
  }
  return rc;
}

int ppo_oracle_get_state(ppo_oracle_handle *h, ppo_ba_state *out) {
  if (out->kf_pose)
    for (int i = 0; i < h->n_kf; i++) {
      double *p = &out->kf_pose[7 * i];
      const SE3 &T = h->kf[i];
      p[0] = T.r.x; p[1] = T.r.y; p[2] = T.r.z; p[3] = T.r.w; p[4] = T.t[0]; p[5] = T.t[1]; p[6] = T.t[2];
    }
  if (out->pt_xyz)
    for (int i = 0; i < h->n_pt; i++)
      for (int j = 0; j < 3; j++) out->pt_xyz[3 * i + j] = h->pt[i][j];
  if (out->pl_coef)
    for (int i = 0; i < h->n_pl; i++)
      for (int j = 0; j < 4; j++) out->pl_coef[4 * i + j] = h->pl[i].c[j];
  if (out->cu_state)
    for (int i = 0; i < h->n_cu; i++) {
      double *c = &out->cu_state[10 * i];
      const Cuboid &q = h->cu[i];
      c[0] = q.pose.t[0]; c[1] = q.pose.t[1]; c[2] = q.pose.t[2];
      c[3] = q.pose.r.x; c[4] = q.pose.r.y; c[5] = q.pose.r.z; c[6] = q.pose.r.w;
      c[7] = q.scale[0]; c[8] = q.scale[1]; c[9] = q.scale[2];
    }
  return PPO_OK;
}

// --- block-level export for parity tests: one linearisation at the current estimates -------------
// Hpp: n_p x n_p row-major (upper blocks filled), b: n_p + 3 n_l, Hll: n_l x 9.  Returns n_p, n_l
// through dims[0], dims[1]; call with NULL outputs first to size the buffers.
int ppo_oracle_debug_linearize(ppo_oracle_handle *h, int32_t dims[2], double *Hpp, double *b, double *Hll, double *chi2) {
  h->initialize_optimization();
  h->build_structure();
  h->compute_active_errors();
  double c = h->active_robust_chi2();
  h->build_system();
  dims[0] = h->n_p;
  dims[1] = h->n_l;
  if (chi2) *chi2 = c;
  if (Hpp) std::copy(h->Hpp.begin(), h->Hpp.end(), Hpp);
  if (b) std::copy(h->b.begin(), h->b.end(), b);
  if (Hll) std::copy(h->Hll.begin(), h->Hll.end(), Hll);
  return PPO_OK;
}
// After debug_linearize: Schur complement + solve for a given lambda. Hschur n_p x n_p (upper), x full.
// x == NULL: the reduced system only (no factorisation: n_p^3/3 flops on one core is minutes for a 1000-key-frame window)
int ppo_oracle_debug_solve(ppo_oracle_handle *h, double lambda, double *Hschur_upper, double *bschur, double *x, int32_t *ok) {
  // reproduce solve_damped but keep a copy of Hschur before the factorisation overwrites it
  h->lambda = lambda;
  h->schur_only = x == nullptr;
  if (h->schur_only) {  // Hschur is not overwritten in this mode: hand it out directly
    bool good = h->solve_damped(lambda);
    h->schur_only = false;
    if (ok) *ok = good;
    if (Hschur_upper) std::copy(h->Hschur.begin(), h->Hschur.end(), Hschur_upper);
    if (bschur) std::copy(h->bschur.begin(), h->bschur.end(), bschur);
    return PPO_OK;
  }
  std::vector<double> Hpp_keep = h->Hpp;
  bool good = h->solve_damped(lambda);
  if (ok) *ok = good;
  if (Hschur_upper) {
    // recompute the un-factorised matrix
    std::vector<double> S = Hpp_keep;
    int n_p = h->n_p;
    for (int i = 0; i < n_p; i++) S[(size_t)i * n_p + i] += lambda;
    for (int l = 0; l < h->n_l; l++) {
      const double *Di = &h->Dinv[9 * (size_t)l];
      auto &blocks = h->Hpl[l];
      for (size_t i1 = 0; i1 < blocks.size(); i1++) {
        const double *Bi = blocks[i1].m;
        double BD[18];
        for (int a = 0; a < 6; a++)
          for (int c = 0; c < 3; c++) BD[a * 3 + c] = Bi[a * 3] * Di[c] + Bi[a * 3 + 1] * Di[3 + c] + Bi[a * 3 + 2] * Di[6 + c];
        for (size_t i2 = i1; i2 < blocks.size(); i2++) {
          const double *Bj = blocks[i2].m;
          for (int a = 0; a < 6; a++)
            for (int c = 0; c < 6; c++)
              S[(size_t)(blocks[i1].pose * 6 + a) * n_p + blocks[i2].pose * 6 + c] -= BD[a * 3] * Bj[c * 3] + BD[a * 3 + 1] * Bj[c * 3 + 1] + BD[a * 3 + 2] * Bj[c * 3 + 2];
        }
      }
    }
    std::copy(S.begin(), S.end(), Hschur_upper);
  }
  if (bschur) std::copy(h->bschur.begin(), h->bschur.end(), bschur);
  if (x) std::copy(h->x.begin(), h->x.end(), x);
  return PPO_OK;
}

// ------------------------------------------------------------------------------------------------
// formula-level entry points for the known-answer tests
// ------------------------------------------------------------------------------------------------
static SE3 se3_in(const double p[7]) { return SE3{Quat{p[0], p[1], p[2], p[3]}, v3(p[4], p[5], p[6])}; }
static void se3_out(const SE3 &T, double p[7]) {
  p[0] = T.r.x; p[1] = T.r.y; p[2] = T.r.z; p[3] = T.r.w; p[4] = T.t[0]; p[5] = T.t[1]; p[6] = T.t[2];
}
static Cuboid cu_in(const double c[10]) {
  Cuboid q;
  q.pose = SE3{Quat{c[3], c[4], c[5], c[6]}, v3(c[0], c[1], c[2])};
  q.scale = v3(c[7], c[8], c[9]);
  return q;
}
static void cu_out(const Cuboid &q, double c[10]) {
  c[0] = q.pose.t[0]; c[1] = q.pose.t[1]; c[2] = q.pose.t[2];
  c[3] = q.pose.r.x; c[4] = q.pose.r.y; c[5] = q.pose.r.z; c[6] = q.pose.r.w;
  c[7] = q.scale[0]; c[8] = q.scale[1]; c[9] = q.scale[2];
}
void ppo_oracle_se3_exp(const double u[6], double out[7]) { se3_out(se3_exp(u), out); }
void ppo_oracle_se3_from_Rt(const double R[9], const double t[3], double out[7]) {
  M3 m;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) m(i, j) = R[3 * i + j];
  se3_out(se3_from_Rt(m, v3(t[0], t[1], t[2])), out);
}
void ppo_oracle_se3_oplus(const double pose[7], const double u[6], double out[7]) { se3_out(se3_oplus(se3_in(pose), u), out); }
void ppo_oracle_se3_map(const double pose[7], const double p[3], double out[3]) {
  V3 r = se3_map(se3_in(pose), v3(p[0], p[1], p[2]));
  out[0] = r[0]; out[1] = r[1]; out[2] = r[2];
}
void ppo_oracle_se3_matrix(const double pose[7], double R[9]) {
  M3 m = quat_to_matrix(se3_in(pose).r);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) R[3 * i + j] = m(i, j);
}
void ppo_oracle_plane_normalize(const double c[4], double out[4]) {
  Plane p = plane_from_vector(c);
  for (int i = 0; i < 4; i++) out[i] = p.c[i];
}
void ppo_oracle_plane_oplus(const double c[4], const double v[3], double out[4]) {
  Plane p = plane_from_vector(c);
  plane_oplus(p, v);
  for (int i = 0; i < 4; i++) out[i] = p.c[i];
}
void ppo_oracle_plane_ominus(int kind, const double a[4], const double bq[4], double out[3]) {
  Plane p = plane_from_vector(a), q = plane_from_vector(bq);
  out[2] = 0;
  if (kind == PPO_PLANE_OBS) plane_ominus(p, q, out);
  else if (kind == PPO_PLANE_VER) plane_ominus_ver(p, q, out);
  else plane_ominus_par(p, q, out);
}
void ppo_oracle_plane_transform(const double pose[7], const double c[4], double out[4]) {
  Plane p = plane_transform(se3_in(pose), plane_from_vector(c));
  for (int i = 0; i < 4; i++) out[i] = p.c[i];
}
void ppo_oracle_cuboid_oplus(const double c[10], unsigned flags, const double u[9], double out[10]) {
  cu_out(cuboid_oplus(cu_in(c), flags, u), out);
}
void ppo_oracle_cuboid_corners(const double c[10], double out[24]) {
  double w[3][8];
  cuboid_corners(cu_in(c), w);
  for (int i = 0; i < 3; i++)
    for (int k = 0; k < 8; k++) out[8 * i + k] = w[i][k];
}
void ppo_oracle_cuboid_project(const double c[10], const double pose[7], const float intr[5], double corners[16], double bbox[4]) {
  double K[9];
  K_from_intr(intr, K);
  double p[2][8];
  cuboid_project(cu_in(c), se3_in(pose), K, p);
  for (int k = 0; k < 8; k++) corners[2 * k] = p[0][k], corners[2 * k + 1] = p[1][k];
  cuboid_project_bbox(cu_in(c), se3_in(pose), K, bbox);
}
void ppo_oracle_cuboid_point_error(const double c[10], const double *pts, int n, double ratio, double prior_weight, double out[3]) {
  point_cuboid_error(cu_in(c), pts, n, ratio, prior_weight, out);
}
void ppo_oracle_cuboid_to_minimal(const double c[10], double out[9]) { cuboid_to_minimal(cu_in(c), out); }
void ppo_oracle_huber(double e, double delta, double rho[3]) { huber(e, delta, rho); }
int ppo_oracle_point_edge(const double pose[7], const double X[3], const float intr[5], const float obs[3], double err[3], double Jpt[9], double Jkf[18]) {
  SE3 T = se3_in(pose);
  V3 x = v3(X[0], X[1], X[2]);
  int D = point_edge_error(T, x, intr, obs, err);
  if (Jpt && Jkf) point_edge_jacobian(T, x, intr, D == 3, Jpt, Jkf);
  return D;
}
int ppo_oracle_plane_edge(int kind, const double pl[4], const double pose[7], const double meas[4], double err[3]) {
  return plane_edge_error(kind, plane_from_vector(pl), se3_in(pose), plane_from_vector(meas), err);
}
int ppo_oracle_cuboid_cam_edge(int kind, const double pose[7], const double c[10], const float intr[5], const double *meas, double err[16]) {
  double K[9];
  K_from_intr(intr, K);
  return cuboid_cam_error(kind, se3_in(pose), cu_in(c), K, meas, err);
}
// dense LDLT KAT: A is n x n row-major symmetric (upper used)
// the same with the linear solver of the given stack: PPO_SOLVER_DENSE_X = LinearSolverDense, PPO_SOLVER_6_3 = LinearSolverEigen
int ppo_oracle_dense_solve_flavour(int solver, int n, const double *A, const double *rhs, double *sol) {
  ppo_oracle_handle h;
  h.P.solver = solver;
  h.n_p = n;
  std::vector<double> M(A, A + (size_t)n * n);
  return h.dense_solve(M, rhs, sol) ? 1 : 0;
}
int ppo_oracle_dense_solve(int n, const double *A, const double *rhs, double *sol) { return ppo_oracle_dense_solve_flavour(PPO_SOLVER_DENSE_X, n, A, rhs, sol); }

}  // extern "C"
