// Seeded synthetic local-BA windows.  Produces the flat graph that the
// Optimizer shim would build from KeyFrame / MapPoint / MapPlane / MapCuboid state
// (src/Optimizer.cc:1997-2714 of the reference): float32 map state converted exactly like
// ORB_SLAM2::Converter does.  Input generation only — no solver arithmetic lives here.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../../include/ppo_synth.h"
#include "ppo_convert.h"

namespace {

struct Rng {  // splitmix64 + Box-Muller: identical streams on every platform with the same libm
  uint64_t s;
  bool has_spare = false;
  double spare = 0;
  explicit Rng(uint64_t seed) : s(seed) {}
  uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
  double uni(double a, double b) { return a + (b - a) * uni(); }
  int below(int n) { return (int)(next() % (uint64_t)n); }
  double normal() {
    if (has_spare) {
      has_spare = false;
      return spare;
    }
    double u1 = uni(), u2 = uni();
    if (u1 < 1e-300) u1 = 1e-300;
    double r = std::sqrt(-2.0 * std::log(u1)), a = 6.283185307179586 * u2;
    spare = r * std::sin(a);
    has_spare = true;
    return r * std::cos(a);
  }
};

struct Mat3 {
  double m[3][3];
};
Mat3 mul(const Mat3 &a, const Mat3 &b) {
  Mat3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
  return r;
}
Mat3 rot_axis(int axis, double a) {
  double c = std::cos(a), s = std::sin(a);
  Mat3 r = {{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}};
  int i = (axis + 1) % 3, j = (axis + 2) % 3;
  r.m[i][i] = c;
  r.m[i][j] = -s;
  r.m[j][i] = s;
  r.m[j][j] = c;
  return r;
}
Mat3 rodrigues(const double w[3]) {
  double th = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  Mat3 K = {{{0, -w[2], w[1]}, {w[2], 0, -w[0]}, {-w[1], w[0], 0}}};
  Mat3 K2 = mul(K, K);
  double a = th < 1e-12 ? 1.0 : std::sin(th) / th, b = th < 1e-12 ? 0.5 : (1 - std::cos(th)) / (th * th);
  Mat3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r.m[i][j] = (i == j) + a * K.m[i][j] + b * K2.m[i][j];
  return r;
}
struct Pose {  // world -> camera
  Mat3 R;
  double t[3];
};
void to_float16(const Pose &P, float T[16]) {
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) T[4 * i + j] = (float)P.R.m[i][j];
    T[4 * i + 3] = (float)P.t[i];
  }
  T[12] = T[13] = T[14] = 0;
  T[15] = 1;
}
inline void cam_point(const Pose &P, const double X[3], double p[3]) {
  for (int i = 0; i < 3; i++) p[i] = P.R.m[i][0] * X[0] + P.R.m[i][1] * X[1] + P.R.m[i][2] * X[2] + P.t[i];
}

// Examples/RGB-D/TUM1.yaml:8-11,26 ; 8 levels, scale 1.2 (:45,48)
const float FX = 517.306408f, FY = 516.469215f, CX = 318.643040f, CY = 255.313989f, BF = 40.0f;
const int IMG_W = 640, IMG_H = 480;

}  // namespace

struct ppo_synth {
  ppo_synth_cfg cfg;
  ppo_ba_graph g;
  std::vector<double> kf_pose, pt_xyz, pl_coef, cu_state, ple_meas, ple_info, cbe_meas, cbe_info, pce_pts, cpe_meas, cpe_info;
  std::vector<double> t_kf_pose, t_pt_xyz, t_pl_coef, t_cu_state;
  std::vector<uint8_t> kf_fixed, cu_flags, ple_kind, cbe_kind;
  std::vector<float> kf_intr, pe_obs, pe_invsigma2;
  std::vector<int32_t> pt_rowptr, pe_kf, ple_plane, ple_kf, cbe_kf, cbe_cuboid, pce_cuboid, pce_rowptr, cpe_cuboid, cpe_plane;
};

extern "C" {

void ppo_synth_config(int ci, int window, ppo_synth_cfg *c) {
  std::memset(c, 0, sizeof *c);
  static const int kf[5] = {10, 50, 200, 50, 1000}, pt[5] = {2000, 20000, 80000, 20000, 400000}, pl[5] = {0, 50, 200, 50, 1000},
                   cu[5] = {0, 10, 50, 10, 200};
  if (ci < 0) ci = 0;
  if (ci > 4) ci = 4;
  c->n_kf = kf[ci];
  c->n_fixed = -1;
  c->n_pt = pt[ci];
  c->n_pl = pl[ci];
  c->n_cu = cu[ci];
  c->seed = 0x50504F00ull + (uint64_t)ci + (uint64_t)window;
  c->cuboid_2d = 1;
  c->corners_2d = 0;
  c->pt_obj_3d = 1;
  c->cuboid_plane = 1;
  c->plane_3d = 1;
  c->outlier_frac = 0.03;
  c->stereo_frac = 0.7;
  c->sort_points = 1;
}

ppo_synth *ppo_synth_create(const ppo_synth_cfg *cfgp) {
  ppo_synth *S = new ppo_synth();
  S->cfg = *cfgp;
  const ppo_synth_cfg &c = S->cfg;
  Rng rng(c.seed);
  const int n_loc = c.n_kf;
  const int n_fix = c.n_fixed >= 0 ? c.n_fixed : std::max(2, n_loc / 10);
  const int n_kf = n_loc + n_fix;
  const double deg = M_PI / 180.0;

  // ---- key-frames ------------------------------------------------------------------------------
  std::vector<Pose> Ttrue(n_kf);
  std::vector<double> theta(n_kf);
  for (int i = 0; i < n_kf; i++) {
    double th = i < n_loc ? 2 * M_PI * i / n_loc : 2 * M_PI * ((i - n_loc) + 0.5) / n_fix;
    theta[i] = th;
    double C[3] = {3 * std::cos(th), 0, 3 * std::sin(th)};
    Mat3 Rwc = {{{-std::sin(th), 0, -std::cos(th)}, {0, 1, 0}, {std::cos(th), 0, -std::sin(th)}}};  // columns x_c y_c z_c
    Rwc = mul(mul(Rwc, rot_axis(1, rng.uni(-10, 10) * deg)), rot_axis(0, rng.uni(-10, 10) * deg));
    Pose P;
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) P.R.m[a][b] = Rwc.m[b][a];
    for (int a = 0; a < 3; a++) P.t[a] = -(P.R.m[a][0] * C[0] + P.R.m[a][1] * C[1] + P.R.m[a][2] * C[2]);
    // the map stores float32 poses: truth is the float-rounded pose
    float T[16];
    to_float16(P, T);
    for (int a = 0; a < 3; a++) {
      for (int b = 0; b < 3; b++) P.R.m[a][b] = T[4 * a + b];
      P.t[a] = T[4 * a + 3];
    }
    Ttrue[i] = P;
  }
  // loop order (by angle) of all key-frames
  std::vector<int> loop(n_kf);
  for (int i = 0; i < n_kf; i++) loop[i] = i;
  std::sort(loop.begin(), loop.end(), [&](int a, int b) { return theta[a] < theta[b] || (theta[a] == theta[b] && a < b); });
  std::vector<int> loop_pos(n_kf);
  for (int i = 0; i < n_kf; i++) loop_pos[loop[i]] = i;

  S->kf_pose.resize(7 * (size_t)n_kf);
  S->t_kf_pose.resize(7 * (size_t)n_kf);
  S->kf_fixed.resize(n_kf);
  S->kf_intr.resize(5 * (size_t)n_kf);
  for (int i = 0; i < n_kf; i++) {
    bool fixed = (i == 0) || i >= n_loc;  // mnId == 0 or a fixed camera (Optimizer.cc:2126,2141)
    S->kf_fixed[i] = fixed;
    float T[16];
    to_float16(Ttrue[i], T);
    ppo::tcw_float_to_pose7(T, &S->t_kf_pose[7 * (size_t)i]);
    Pose P = Ttrue[i];
    if (!fixed) {
      double w[3] = {0.01 * rng.normal(), 0.01 * rng.normal(), 0.01 * rng.normal()};
      double v[3] = {0.02 * rng.normal(), 0.02 * rng.normal(), 0.02 * rng.normal()};
      Mat3 dR = rodrigues(w);
      P.R = mul(dR, Ttrue[i].R);
      for (int a = 0; a < 3; a++)
        P.t[a] = dR.m[a][0] * Ttrue[i].t[0] + dR.m[a][1] * Ttrue[i].t[1] + dR.m[a][2] * Ttrue[i].t[2] + v[a];
    }
    to_float16(P, T);
    ppo::tcw_float_to_pose7(T, &S->kf_pose[7 * (size_t)i]);
    float *in = &S->kf_intr[5 * (size_t)i];
    in[0] = FX; in[1] = FY; in[2] = CX; in[3] = CY; in[4] = BF;
  }

  // ORBextractor.cc:416-430 level sigmas in float
  float sf[8], inv_sigma2[8], sigma[8];
  sf[0] = 1.0f;
  for (int i = 1; i < 8; i++) sf[i] = sf[i - 1] * 1.2f;
  for (int i = 0; i < 8; i++) {
    float s2 = sf[i] * sf[i];
    inv_sigma2[i] = 1.0f / s2;
    sigma[i] = sf[i];
  }

  // ---- map points ------------------------------------------------------------------------------
  struct Obs {
    int kf;
    float u, v, ur, is2;
  };
  struct Pt {
    double Xt[3];
    float Xi[3];
    int first;
    std::vector<Obs> obs;
  };
  std::vector<Pt> pts(c.n_pt);
  std::vector<int> vis;
  vis.reserve(n_kf);
  for (int i = 0; i < c.n_pt; i++) {
    int k = 2 + (i % 9);
    if (k > n_kf) k = n_kf;
    Pt &P = pts[i];
    for (int attempt = 0;; attempt++) {
      double X[3] = {rng.uni(-4, 4), rng.uni(-1.5, 1.5), rng.uni(-4, 4)};
      float Xf[3] = {(float)X[0], (float)X[1], (float)X[2]};
      for (int a = 0; a < 3; a++) X[a] = Xf[a];
      vis.clear();
      for (int li = 0; li < n_kf; li++) {
        int kf = loop[li];
        double p[3];
        cam_point(Ttrue[kf], X, p);
        if (p[2] < 0.3) continue;
        double u = FX * p[0] / p[2] + CX, v = FY * p[1] / p[2] + CY;
        if (u < 1 || u > IMG_W - 2 || v < 1 || v > IMG_H - 2) continue;
        vis.push_back(kf);
      }
      if ((int)vis.size() < k) continue;
      for (int a = 0; a < 3; a++) P.Xt[a] = X[a];
      break;
    }
    // k nearest-in-loop-order key-frames around a random anchor, 5 % long-range substitutions
    int nv = (int)vis.size();
    int a0 = rng.below(nv);
    std::vector<int> chosen;
    for (int s = 0; (int)chosen.size() < k; s++) {
      int off = (s + 1) / 2 * ((s & 1) ? 1 : -1);
      chosen.push_back(vis[((a0 + off) % nv + nv) % nv]);
    }
    for (int s = 0; s < k; s++)
      if (rng.uni() < 0.05 && nv > k) {
        for (int tries = 0; tries < 8; tries++) {
          int cand = vis[rng.below(nv)];
          if (std::find(chosen.begin(), chosen.end(), cand) == chosen.end()) {
            chosen[s] = cand;
            break;
          }
        }
      }
    std::sort(chosen.begin(), chosen.end());
    chosen.erase(std::unique(chosen.begin(), chosen.end()), chosen.end());
    P.first = n_kf;
    for (int kf : chosen) {
      double p[3];
      cam_point(Ttrue[kf], P.Xt, p);
      int oct = rng.below(8);
      double sg = sigma[oct];
      double u = FX * p[0] / p[2] + CX, v = FY * p[1] / p[2] + CY;
      double ur = u - BF / p[2];
      u += sg * rng.normal();
      v += sg * rng.normal();
      ur += sg * rng.normal();
      bool stereo = rng.uni() < c.stereo_frac;
      if (rng.uni() < c.outlier_frac) {
        double mag = rng.uni(20, 60), ang = rng.uni(0, 2 * M_PI);
        u += mag * std::cos(ang);
        v += mag * std::sin(ang);
      }
      Obs o;
      o.kf = kf;
      o.u = (float)u;
      o.v = (float)v;
      o.ur = stereo ? (float)std::max(ur, 0.0) : -1.0f;
      o.is2 = inv_sigma2[oct];
      P.obs.push_back(o);
      if (kf < n_loc) P.first = std::min(P.first, kf);
    }
    for (int a = 0; a < 3; a++) P.Xi[a] = (float)(P.Xt[a] + 0.03 * rng.normal());
  }
  if (c.sort_points) {
    // lLocalMapPoints is collected key-frame by key-frame (Optimizer.cc:2012-2030)
    std::stable_sort(pts.begin(), pts.end(), [](const Pt &a, const Pt &b) { return a.first < b.first; });
  }
  S->pt_xyz.resize(3 * (size_t)c.n_pt);
  S->t_pt_xyz.resize(3 * (size_t)c.n_pt);
  S->pt_rowptr.resize(c.n_pt + 1);
  S->pt_rowptr[0] = 0;
  for (int i = 0; i < c.n_pt; i++) {
    for (int a = 0; a < 3; a++) {
      S->pt_xyz[3 * (size_t)i + a] = pts[i].Xi[a];
      S->t_pt_xyz[3 * (size_t)i + a] = pts[i].Xt[a];
    }
    for (const Obs &o : pts[i].obs) {
      S->pe_kf.push_back(o.kf);
      S->pe_obs.push_back(o.u);
      S->pe_obs.push_back(o.v);
      S->pe_obs.push_back(o.ur);
      S->pe_invsigma2.push_back(o.is2);
    }
    S->pt_rowptr[i + 1] = (int32_t)S->pe_kf.size();
  }
  pts.clear();
  pts.shrink_to_fit();

  // ---- planes ----------------------------------------------------------------------------------
  const double angleInfo = 3282.8 / (1.0 * 1.0), disInfo = 100.0 * 100.0;  // Optimizer.cc:2194-2197
  const double pvInfo = 3282.8 / (0.5 * 0.5);                                // :2198-2201
  auto plane_in_cam = [&](const Pose &T, const double pw[4], double pc[4]) {
    for (int a = 0; a < 3; a++) pc[a] = T.R.m[a][0] * pw[0] + T.R.m[a][1] * pw[1] + T.R.m[a][2] * pw[2];
    pc[3] = pw[3] - (T.t[0] * pc[0] + T.t[1] * pc[1] + T.t[2] * pc[2]);
  };
  auto perturb_normal = [&](double n[3], double sd_rad) {
    // add small components along two directions perpendicular to n
    double a[3] = {1, 0, 0};
    if (std::fabs(n[0]) > 0.9) a[0] = 0, a[1] = 1;
    double e1[3] = {n[1] * a[2] - n[2] * a[1], n[2] * a[0] - n[0] * a[2], n[0] * a[1] - n[1] * a[0]};
    double l = std::sqrt(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]);
    for (int i = 0; i < 3; i++) e1[i] /= l;
    double e2[3] = {n[1] * e1[2] - n[2] * e1[1], n[2] * e1[0] - n[0] * e1[2], n[0] * e1[1] - n[1] * e1[0]};
    double d1 = std::tan(sd_rad * rng.normal()), d2 = std::tan(sd_rad * rng.normal());
    for (int i = 0; i < 3; i++) n[i] += d1 * e1[i] + d2 * e2[i];
    l = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    for (int i = 0; i < 3; i++) n[i] /= l;
  };
  S->pl_coef.resize(4 * (size_t)c.n_pl);
  S->t_pl_coef.resize(4 * (size_t)c.n_pl);
  for (int i = 0; i < c.n_pl; i++) {
    double n[3] = {0, 0, 0};
    n[i % 3] = rng.uni() < 0.5 ? 1 : -1;
    perturb_normal(n, 5 * deg);
    double pw[4] = {n[0], n[1], n[2], rng.uni(1, 5)};
    float pf[4] = {(float)pw[0], (float)pw[1], (float)pw[2], (float)pw[3]};
    ppo::plane_float_to_coef(pf, &S->t_pl_coef[4 * (size_t)i]);
    const double *tw = &S->t_pl_coef[4 * (size_t)i];
    // initial estimate
    double ni[3] = {tw[0], tw[1], tw[2]};
    perturb_normal(ni, 1 * deg);
    float pi4[4] = {(float)ni[0], (float)ni[1], (float)ni[2], (float)(tw[3] + 0.02 * rng.normal())};
    ppo::plane_float_to_coef(pi4, &S->pl_coef[4 * (size_t)i]);
    if (!c.plane_3d) continue;
    int span = std::max(2, n_loc / 4), start = rng.below(n_kf);
    for (int s = 0; s < span && s < n_kf; s++) {
      int kf = loop[(start + s) % n_kf];
      double pc[4];
      plane_in_cam(Ttrue[kf], tw, pc);
      if (std::fabs(pc[3]) < 0.3) continue;  // sign of d ambiguous when the camera sits on the plane
      double nm[3] = {pc[0], pc[1], pc[2]};
      perturb_normal(nm, 0.5 * deg);
      float mf[4] = {(float)nm[0], (float)nm[1], (float)nm[2], (float)(pc[3] + (pc[3] < 0 ? -1 : 1) * 0.01 * rng.normal())};
      double m[4];
      ppo::plane_float_to_coef(mf, m);
      S->ple_plane.push_back(i);
      S->ple_kf.push_back(kf);
      S->ple_kind.push_back(PPO_PLANE_OBS);
      S->ple_meas.insert(S->ple_meas.end(), m, m + 4);
      S->ple_info.push_back(angleInfo);
      S->ple_info.push_back(angleInfo);
      S->ple_info.push_back(disInfo);
      for (int kind = PPO_PLANE_VER; kind <= PPO_PLANE_PAR; kind++) {
        if (rng.uni() >= 0.10) continue;
        double q[3];
        if (kind == PPO_PLANE_PAR) {
          double sgn = rng.uni() < 0.5 ? 1 : -1;
          for (int a = 0; a < 3; a++) q[a] = sgn * pc[a];
        } else {
          double r[3] = {rng.normal(), rng.normal(), rng.normal()};
          q[0] = pc[1] * r[2] - pc[2] * r[1];
          q[1] = pc[2] * r[0] - pc[0] * r[2];
          q[2] = pc[0] * r[1] - pc[1] * r[0];
          double l = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
          for (int a = 0; a < 3; a++) q[a] /= l;
        }
        perturb_normal(q, 0.5 * deg);
        float qf[4] = {(float)q[0], (float)q[1], (float)q[2], (float)rng.uni(0.5, 5)};
        double mm[4];
        ppo::plane_float_to_coef(qf, mm);
        S->ple_plane.push_back(i);
        S->ple_kf.push_back(kf);
        S->ple_kind.push_back((uint8_t)kind);
        S->ple_meas.insert(S->ple_meas.end(), mm, mm + 4);
        S->ple_info.push_back(pvInfo);
        S->ple_info.push_back(pvInfo);
        S->ple_info.push_back(0.0);
      }
    }
  }

  // ---- cuboids ---------------------------------------------------------------------------------
  S->cu_state.resize(10 * (size_t)c.n_cu);
  S->t_cu_state.resize(10 * (size_t)c.n_cu);
  S->cu_flags.assign(c.n_cu, PPO_CU_FIXROLLPITCH | PPO_CU_FIXHEIGHT);  // Optimizer.cc:2167-2168
  static const double sgn[3][8] = {{1, 1, -1, -1, 1, 1, -1, -1}, {1, -1, -1, 1, 1, -1, -1, 1}, {-1, -1, -1, -1, 1, 1, 1, 1}};
  struct CamObs {
    int kf, cu;
    double bbox[4], corners[16], local[10];  // local: the cuboid in the camera frame [t q scale] (EdgeSE3Cuboid measurement)
  };
  Rng rng3(c.seed * 7919u + 17u);  // own stream: the content of the other edge kinds does not depend on cuboid_3d
  std::vector<CamObs> cobs;
  S->pce_rowptr.push_back(0);
  // MapPlane holds ONE asso_cuboid_id (MapPlane.h:67): distinct planes for distinct cuboids while planes last
  const int cpe_plane0 = c.n_pl > 0 ? rng.below(c.n_pl) : 0;
  for (int i = 0; i < c.n_cu; i++) {
    double yaw = rng.uni(-M_PI, M_PI);
    double ctr[3] = {rng.uni(-1, 1), rng.uni(-0.5, 0.5), rng.uni(-1, 1)};
    double sc[3] = {rng.uni(0.2, 0.8), rng.uni(0.2, 0.8), rng.uni(0.2, 0.8)};
    auto write_state = [](double *o, const double t[3], double yw, const double s[3]) {
      double qz = std::sin(0.5 * yw), qw = std::cos(0.5 * yw);
      if (qw < 0) qz = -qz, qw = -qw;
      o[0] = t[0]; o[1] = t[1]; o[2] = t[2];
      o[3] = 0; o[4] = 0; o[5] = qz; o[6] = qw;
      o[7] = s[0]; o[8] = s[1]; o[9] = s[2];
    };
    write_state(&S->t_cu_state[10 * (size_t)i], ctr, yaw, sc);
    double ti[3] = {ctr[0] + 0.05 * rng.normal(), ctr[1], ctr[2] + 0.05 * rng.normal()};
    double si[3] = {sc[0] * (1 + 0.05 * rng.normal()), sc[1] * (1 + 0.05 * rng.normal()), sc[2] * (1 + 0.05 * rng.normal())};
    write_state(&S->cu_state[10 * (size_t)i], ti, yaw + 3 * deg * rng.normal(), si);
    double cy = std::cos(yaw), sy = std::sin(yaw);
    double cw[3][8];
    for (int k = 0; k < 8; k++) {
      double lx = sc[0] * sgn[0][k], ly = sc[1] * sgn[1][k], lz = sc[2] * sgn[2][k];
      cw[0][k] = cy * lx - sy * ly + ctr[0];
      cw[1][k] = sy * lx + cy * ly + ctr[1];
      cw[2][k] = lz + ctr[2];
    }
    int span = std::max(2, n_loc / 5), start = rng.below(n_kf);
    for (int s = 0; s < span && s < n_kf; s++) {
      int kf = loop[(start + s) % n_kf];
      CamObs o;
      o.kf = kf;
      o.cu = i;
      bool ok = true;
      double mn[2] = {1e30, 1e30}, mx[2] = {-1e30, -1e30};
      for (int k = 0; k < 8; k++) {
        double X[3] = {cw[0][k], cw[1][k], cw[2][k]}, p[3];
        cam_point(Ttrue[kf], X, p);
        if (p[2] < 0.3) ok = false;
        double u = FX * p[0] / p[2] + CX, v = FY * p[1] / p[2] + CY;
        o.corners[2 * k] = u;
        o.corners[2 * k + 1] = v;
        mn[0] = std::min(mn[0], u); mx[0] = std::max(mx[0], u);
        mn[1] = std::min(mn[1], v); mx[1] = std::max(mx[1], v);
      }
      if (!ok) continue;
      o.bbox[0] = (mn[0] + mx[0]) / 2 + 2 * rng.normal();
      o.bbox[1] = (mn[1] + mx[1]) / 2 + 2 * rng.normal();
      o.bbox[2] = (mx[0] - mn[0]) + 2 * rng.normal();
      o.bbox[3] = (mx[1] - mn[1]) + 2 * rng.normal();
      for (int k = 0; k < 16; k++) o.corners[k] += 2 * rng.normal();
      // cv::Rect bbox_2d margin test (Optimizer.cc:2456-2460), object_boundary_margin = 5
      int rx = (int)std::floor(o.bbox[0] - o.bbox[2] / 2), ry = (int)std::floor(o.bbox[1] - o.bbox[3] / 2);
      int rw = (int)std::ceil(o.bbox[2]), rh = (int)std::ceil(o.bbox[3]);
      if (!(rx > 5 && ry > 5 && rx + rw < IMG_W - 5 && ry + rh < IMG_H - 5)) continue;
      if (c.cuboid_3d) {  // the detector's 3-D cuboid in the camera frame: noisy, and with an arbitrary choice of the front face
        const int quarter = rng3.below(4);  // measured yaw off by a multiple of 90 degrees, x / y half sizes swapped accordingly
        const double yw = yaw + 2 * deg * rng3.normal() + quarter * (M_PI / 2.0);
        const double cyw = std::cos(yw), syw = std::sin(yw);
        const Pose &T = Ttrue[kf];
        float Tco[16] = {0};
        const double Rwo[3][3] = {{cyw, -syw, 0}, {syw, cyw, 0}, {0, 0, 1}};
        for (int r = 0; r < 3; r++) {
          for (int q = 0; q < 3; q++) {
            double a = 0;
            for (int m = 0; m < 3; m++) a += T.R.m[r][m] * Rwo[m][q];
            Tco[4 * r + q] = (float)a;
          }
          Tco[4 * r + 3] = (float)(T.R.m[r][0] * ctr[0] + T.R.m[r][1] * ctr[1] + T.R.m[r][2] * ctr[2] + T.t[r] + 0.02 * rng3.normal());
        }
        Tco[15] = 1;
        double p7[7];
        ppo::tcw_float_to_pose7(Tco, p7);
        o.local[0] = p7[4], o.local[1] = p7[5], o.local[2] = p7[6];
        o.local[3] = p7[0], o.local[4] = p7[1], o.local[5] = p7[2], o.local[6] = p7[3];
        const bool sw = quarter & 1;
        o.local[7] = (sw ? sc[1] : sc[0]) * (1 + 0.03 * rng3.normal());
        o.local[8] = (sw ? sc[0] : sc[1]) * (1 + 0.03 * rng3.normal());
        o.local[9] = sc[2] * (1 + 0.03 * rng3.normal());
      }
      cobs.push_back(o);
    }
    if (c.pt_obj_3d) {
      int npts = 30;
      for (int j = 0; j < npts; j++) {
        double l[3] = {rng.uni(-1, 1) * sc[0], rng.uni(-1, 1) * sc[1], rng.uni(-1, 1) * sc[2]};
        if (rng.uni() < 0.10) {
          int ax = rng.below(3);
          l[ax] = (rng.uni() < 0.5 ? -1 : 1) * (sc[ax] + rng.uni(0, 1));
        }
        double X[3] = {cy * l[0] - sy * l[1] + ctr[0], sy * l[0] + cy * l[1] + ctr[1], l[2] + ctr[2]};
        for (int a = 0; a < 3; a++) S->pce_pts.push_back((double)(float)X[a]);
      }
      S->pce_cuboid.push_back(i);
      S->pce_rowptr.push_back((int32_t)(S->pce_pts.size() / 3));
    }
    if (c.cuboid_plane && c.n_pl > 0 && i < c.n_pl) {
      S->cpe_cuboid.push_back(i);
      S->cpe_plane.push_back((cpe_plane0 + i) % c.n_pl);
      S->cpe_meas.push_back(0.01 * rng.normal());
      S->cpe_meas.push_back(0.01 * rng.normal());
      S->cpe_meas.push_back(0.02 * rng.normal());
      S->cpe_info.push_back(3282.8 / 4.0);  // cuboid_plane_angle_info = 2 (Parameters.cc:70, Optimizer.cc:2664-2665)
      S->cpe_info.push_back(3282.8 / 4.0);
      S->cpe_info.push_back(100.0 * 100.0);
    }
  }
  const double cam_info = (1.0 * 0.7) * (1.0 * 0.7);  // (ba_weight * meas_quality)^2, Optimizer.cc:2462-2465
  const double se3_info = (1.0 * 0.75) * (1.0 * 0.75);  // (ba_weight_SE3 * meas_quality)^2, Optimizer.cc:1786-1790
  for (int pass = 0; pass < 3; pass++) {
    if (pass == 0 && !c.cuboid_2d) continue;
    if (pass == 1 && !c.corners_2d) continue;
    if (pass == 2 && !c.cuboid_3d) continue;
    for (const CamObs &o : cobs) {
      S->cbe_kf.push_back(o.kf);
      S->cbe_cuboid.push_back(o.cu);
      S->cbe_kind.push_back(pass == 0 ? PPO_CUBOID_BBOX : (pass == 1 ? PPO_CUBOID_CORNER : PPO_CUBOID_SE3));
      double m[16] = {0};
      if (pass == 0) std::memcpy(m, o.bbox, sizeof o.bbox);
      else if (pass == 1) std::memcpy(m, o.corners, sizeof o.corners);
      else std::memcpy(m, o.local, sizeof o.local);
      S->cbe_meas.insert(S->cbe_meas.end(), m, m + 16);
      S->cbe_info.push_back(pass == 2 ? se3_info : cam_info);
    }
  }

  // ---- publish ---------------------------------------------------------------------------------
  ppo_ba_graph &g = S->g;
  std::memset(&g, 0, sizeof g);
  g.n_kf = n_kf;
  g.kf_pose = S->kf_pose.data();
  g.kf_fixed = S->kf_fixed.data();
  g.kf_intr = S->kf_intr.data();
  g.n_pt = c.n_pt;
  g.pt_xyz = S->pt_xyz.data();
  g.pt_fixed = nullptr;
  g.n_pl = c.n_pl;
  g.pl_coef = S->pl_coef.data();
  g.n_cu = c.n_cu;
  g.cu_state = S->cu_state.data();
  g.cu_flags = S->cu_flags.data();
  g.pt_rowptr = S->pt_rowptr.data();
  g.n_pe = (int32_t)S->pe_kf.size();
  g.pe_kf = S->pe_kf.data();
  g.pe_obs = S->pe_obs.data();
  g.pe_invsigma2 = S->pe_invsigma2.data();
  g.n_ple = (int32_t)S->ple_plane.size();
  g.ple_plane = S->ple_plane.data();
  g.ple_kf = S->ple_kf.data();
  g.ple_kind = S->ple_kind.data();
  g.ple_meas = S->ple_meas.data();
  g.ple_info = S->ple_info.data();
  g.n_cbe = (int32_t)S->cbe_kf.size();
  g.cbe_kf = S->cbe_kf.data();
  g.cbe_cuboid = S->cbe_cuboid.data();
  g.cbe_kind = S->cbe_kind.data();
  g.cbe_meas = S->cbe_meas.data();
  g.cbe_info = S->cbe_info.data();
  g.n_pce = (int32_t)S->pce_cuboid.size();
  g.pce_cuboid = S->pce_cuboid.data();
  g.pce_rowptr = S->pce_rowptr.data();
  g.pce_pts = S->pce_pts.data();
  g.n_cpe = (int32_t)S->cpe_cuboid.size();
  g.cpe_cuboid = S->cpe_cuboid.data();
  g.cpe_plane = S->cpe_plane.data();
  g.cpe_meas = S->cpe_meas.data();
  g.cpe_info = S->cpe_info.data();
  return S;
}

const ppo_ba_graph *ppo_synth_graph(const ppo_synth *s) { return &s->g; }

void ppo_synth_truth(const ppo_synth *s, ppo_ba_state *out) {
  if (out->kf_pose) std::copy(s->t_kf_pose.begin(), s->t_kf_pose.end(), out->kf_pose);
  if (out->pt_xyz) std::copy(s->t_pt_xyz.begin(), s->t_pt_xyz.end(), out->pt_xyz);
  if (out->pl_coef) std::copy(s->t_pl_coef.begin(), s->t_pl_coef.end(), out->pl_coef);
  if (out->cu_state) std::copy(s->t_cu_state.begin(), s->t_cu_state.end(), out->cu_state);
}

void ppo_synth_destroy(ppo_synth *s) { delete s; }

}  // extern "C"
