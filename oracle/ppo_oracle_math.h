// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
// CPU restatement (plain C++, double, no dependencies) of the geometry the reference's local
// bundle adjustment uses.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may use anything under oracle/.  PARITY UNPINNED BY THE REFERENCE:
// the reference ships no tests / golden vectors for this path and cannot be compiled here
// (needs Eigen3, un-vendored); correctness of this restatement is pinned by
// formula-level known-answer tests (tests/test_oracle_kat.py) and an independent numpy
// transliteration (tests/np_ref.py).
//
// Each function cites the reference lines it follows (paths relative to /root/reference).
// The un-vendored dependency is Eigen3 (>= 3.1.0, CMakeLists.txt:39); its published algorithms
// for Quaterniond(Matrix3d), toRotationMatrix, quaternion*vector, AngleAxis and LDLT are restated.
#pragma once
#include <cmath>
#include <cstring>

namespace ppo_oracle {

struct V3 {
  double v[3];
  double &operator[](int i) { return v[i]; }
  double operator[](int i) const { return v[i]; }
};
struct M3 {
  double m[3][3];
  double &operator()(int i, int j) { return m[i][j]; }
  double operator()(int i, int j) const { return m[i][j]; }
};
struct Quat {
  double x, y, z, w;
};
struct SE3 {  // g2o::SE3Quat: _r, _t   (Thirdparty/g2o/g2o/types/se3quat.h:46-47)
  Quat r;
  V3 t;
};
struct Plane {  // g2o::Plane3D::_coeffs (include/G2O_Plane3D.h:127)
  double c[4];
};
struct Cuboid {  // g2o::cuboid: pose (object->world) + half scale (include/g2o_cuboid.h:33-34)
  SE3 pose;
  V3 scale;
};

inline V3 v3(double a, double b, double c) { return V3{{a, b, c}}; }
inline V3 operator+(const V3 &a, const V3 &b) { return v3(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
inline V3 operator-(const V3 &a, const V3 &b) { return v3(a[0] - b[0], a[1] - b[1], a[2] - b[2]); }
inline V3 operator*(double s, const V3 &a) { return v3(s * a[0], s * a[1], s * a[2]); }
inline double dot(const V3 &a, const V3 &b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline V3 cross(const V3 &a, const V3 &b) {
  return v3(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
}
inline double norm(const V3 &a) { return std::sqrt(dot(a, a)); }
inline M3 m3_identity() {
  M3 r;
  std::memset(&r, 0, sizeof r);
  r(0, 0) = r(1, 1) = r(2, 2) = 1.0;
  return r;
}
inline M3 operator*(const M3 &a, const M3 &b) {
  M3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r(i, j) = a(i, 0) * b(0, j) + a(i, 1) * b(1, j) + a(i, 2) * b(2, j);
  return r;
}
inline V3 operator*(const M3 &a, const V3 &b) {
  return v3(a(0, 0) * b[0] + a(0, 1) * b[1] + a(0, 2) * b[2], a(1, 0) * b[0] + a(1, 1) * b[1] + a(1, 2) * b[2],
            a(2, 0) * b[0] + a(2, 1) * b[1] + a(2, 2) * b[2]);
}
inline M3 transpose(const M3 &a) {
  M3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r(i, j) = a(j, i);
  return r;
}
inline M3 add_scaled(const M3 &a, double s, const M3 &b) {
  M3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r(i, j) = a(i, j) + s * b(i, j);
  return r;
}
// skew(): Thirdparty/g2o/g2o/types/se3_ops.hpp:27-38
inline M3 skew(const V3 &v) {
  M3 m;
  std::memset(&m, 0, sizeof m);
  m(0, 1) = -v[2];
  m(0, 2) = v[1];
  m(1, 2) = -v[0];
  m(1, 0) = v[2];
  m(2, 0) = -v[1];
  m(2, 1) = v[0];
  return m;
}

// ---- Eigen::Quaterniond restated -----------------------------------------------------------
// Quaterniond(Matrix3d): Eigen/src/Geometry/Quaternion.h, quaternionbase_assign_impl<.,3,3>
inline Quat quat_from_matrix(const M3 &m) {
  double q[4];  // x y z w
  double t = m(0, 0) + m(1, 1) + m(2, 2);
  if (t > 0.0) {
    t = std::sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (m(2, 1) - m(1, 2)) * t;
    q[1] = (m(0, 2) - m(2, 0)) * t;
    q[2] = (m(1, 0) - m(0, 1)) * t;
  } else {
    int i = 0;
    if (m(1, 1) > m(0, 0)) i = 1;
    if (m(2, 2) > m(i, i)) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (m(k, j) - m(j, k)) * t;
    q[j] = (m(j, i) + m(i, j)) * t;
    q[k] = (m(k, i) + m(i, k)) * t;
  }
  return Quat{q[0], q[1], q[2], q[3]};
}
// QuaternionBase::toRotationMatrix
inline M3 quat_to_matrix(const Quat &q) {
  M3 r;
  const double tx = 2.0 * q.x, ty = 2.0 * q.y, tz = 2.0 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  r(0, 0) = 1.0 - (tyy + tzz);
  r(0, 1) = txy - twz;
  r(0, 2) = txz + twy;
  r(1, 0) = txy + twz;
  r(1, 1) = 1.0 - (txx + tzz);
  r(1, 2) = tyz - twx;
  r(2, 0) = txz - twy;
  r(2, 1) = tyz + twx;
  r(2, 2) = 1.0 - (txx + tyy);
  return r;
}
// QuaternionBase::_transformVector
inline V3 quat_rotate(const Quat &q, const V3 &v) {
  V3 u = v3(q.x, q.y, q.z);
  V3 uv = cross(u, v);
  uv = uv + uv;
  return v + q.w * uv + cross(u, uv);
}
// quaternion product (internal::quat_product)
inline Quat quat_mul(const Quat &a, const Quat &b) {
  Quat r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}
inline Quat quat_conj(const Quat &q) { return Quat{-q.x, -q.y, -q.z, q.w}; }
// SE3Quat::normalizeRotation  (se3quat.h:331-336): flip to w>=0, then normalise
inline void normalize_rotation(Quat &q) {
  if (q.w < 0) {
    q.x = -q.x;
    q.y = -q.y;
    q.z = -q.z;
    q.w = -q.w;
  }
  double n = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  q.x /= n;
  q.y /= n;
  q.z /= n;
  q.w /= n;
}
// Eigen::AngleAxisd::toRotationMatrix
inline M3 angle_axis_matrix(double angle, const V3 &axis) {
  M3 res;
  V3 sin_axis = std::sin(angle) * axis;
  double c = std::cos(angle);
  V3 cos1_axis = (1.0 - c) * axis;
  double tmp;
  tmp = cos1_axis[0] * axis[1];
  res(0, 1) = tmp - sin_axis[2];
  res(1, 0) = tmp + sin_axis[2];
  tmp = cos1_axis[0] * axis[2];
  res(0, 2) = tmp + sin_axis[1];
  res(2, 0) = tmp - sin_axis[1];
  tmp = cos1_axis[1] * axis[2];
  res(1, 2) = tmp - sin_axis[0];
  res(2, 1) = tmp + sin_axis[0];
  res(0, 0) = cos1_axis[0] * axis[0] + c;
  res(1, 1) = cos1_axis[1] * axis[1] + c;
  res(2, 2) = cos1_axis[2] * axis[2] + c;
  return res;
}
inline Quat quat_from_angle_axis(double angle, const V3 &axis) {
  double s = std::sin(0.5 * angle);
  return Quat{s * axis[0], s * axis[1], s * axis[2], std::cos(0.5 * angle)};
}

// ---- g2o::SE3Quat ---------------------------------------------------------------------------
// SE3Quat(R,t): se3quat.h:58-60 ; Converter::toSE3Quat: src/Converter.cc:37-47
inline SE3 se3_from_Rt(const M3 &R, const V3 &t) {
  SE3 s;
  s.r = quat_from_matrix(R);
  s.t = t;
  normalize_rotation(s.r);
  return s;
}
inline SE3 se3_from_qt(const Quat &q, const V3 &t) {  // se3quat.h:62-64
  SE3 s{q, t};
  normalize_rotation(s.r);
  return s;
}
inline SE3 se3_identity() { return SE3{Quat{0, 0, 0, 1}, v3(0, 0, 0)}; }
// operator*: se3quat.h:107-113
inline SE3 se3_mul(const SE3 &a, const SE3 &b) {
  SE3 r = a;
  r.t = r.t + quat_rotate(a.r, b.t);
  r.r = quat_mul(a.r, b.r);
  normalize_rotation(r.r);
  return r;
}
// inverse: se3quat.h:126-131
inline SE3 se3_inverse(const SE3 &a) {
  SE3 r;
  r.r = quat_conj(a.r);
  r.t = quat_rotate(r.r, -1.0 * a.t);
  return r;
}
// map: se3quat.h:268-271
inline V3 se3_map(const SE3 &T, const V3 &p) { return quat_rotate(T.r, p) + T.t; }
// exp: se3quat.h:274-308 (small-angle branch theta < 1e-5: R = I + W + W^2, V = R)
inline SE3 se3_exp(const double u[6]) {
  V3 omega = v3(u[0], u[1], u[2]);
  V3 upsilon = v3(u[3], u[4], u[5]);
  double theta = norm(omega);
  M3 Omega = skew(omega);
  M3 R, V;
  M3 I = m3_identity();
  if (theta < 0.00001) {
    R = add_scaled(add_scaled(I, 1.0, Omega), 1.0, Omega * Omega);
    V = R;
  } else {
    M3 Omega2 = Omega * Omega;
    R = add_scaled(add_scaled(I, std::sin(theta) / theta, Omega), (1 - std::cos(theta)) / (theta * theta), Omega2);
    V = add_scaled(add_scaled(I, (1 - std::cos(theta)) / (theta * theta), Omega),
                   (theta - std::sin(theta)) / (std::pow(theta, 3)), Omega2);
  }
  return se3_from_qt(quat_from_matrix(R), V * upsilon);
}
// VertexSE3Expmap::oplusImpl: types_six_dof_expmap.h:88-91   estimate <- exp(update) * estimate
inline SE3 se3_oplus(const SE3 &est, const double u[6]) { return se3_mul(se3_exp(u), est); }

// ---- g2o::Plane3D (include/G2O_Plane3D.h) -----------------------------------------------------
// normalize: :120-125
inline void plane_normalize(Plane &p) {
  double n = std::sqrt(p.c[0] * p.c[0] + p.c[1] * p.c[1] + p.c[2] * p.c[2]);
  double inv = 1. / n;
  for (int i = 0; i < 4; i++) p.c[i] = p.c[i] * inv;
  if (p.c[3] < 0.0)
    for (int i = 0; i < 4; i++) p.c[i] = -p.c[i];
}
inline Plane plane_from_vector(const double c[4]) {  // fromVector :45-48
  Plane p;
  for (int i = 0; i < 4; i++) p.c[i] = c[i];
  plane_normalize(p);
  return p;
}
inline V3 plane_normal(const Plane &p) { return v3(p.c[0], p.c[1], p.c[2]); }
inline double plane_distance(const Plane &p) { return -p.c[3]; }                               // :58-60
inline double azimuth(const V3 &v) { return std::atan2(v[1], v[0]); }                           // :50-52
inline double elevation(const V3 &v) { return std::atan2(v[2], std::sqrt(v[0] * v[0] + v[1] * v[1])); }  // :54-56
// rotation(v) = (AngleAxis(az, Z) * AngleAxis(-el, Y)).toRotationMatrix()   :66-72
inline M3 plane_rotation(const V3 &v) {
  Quat qa = quat_from_angle_axis(azimuth(v), v3(0, 0, 1));
  Quat qe = quat_from_angle_axis(-elevation(v), v3(0, 1, 0));
  return quat_to_matrix(quat_mul(qa, qe));
}
// oplus: :74-87
inline void plane_oplus(Plane &p, const double v[3]) {
  double az = v[0], el = v[1];
  double s = std::sin(el), c = std::cos(el);
  V3 n = v3(c * std::cos(az), c * std::sin(az), s);
  M3 R = plane_rotation(plane_normal(p));
  double d = plane_distance(p) + v[2];
  V3 rn = R * n;
  p.c[0] = rn[0];
  p.c[1] = rn[1];
  p.c[2] = rn[2];
  p.c[3] = -d;
  plane_normalize(p);
}
// ominus: :89-95   (this = p, argument = q)
inline void plane_ominus(const Plane &p, const Plane &q, double out[3]) {
  M3 R = transpose(plane_rotation(plane_normal(p)));
  V3 n = R * plane_normal(q);
  double d = plane_distance(p) - plane_distance(q);
  out[0] = azimuth(n);
  out[1] = elevation(n);
  out[2] = d;
}
// ominus_ver: :97-106
inline void plane_ominus_ver(const Plane &p, const Plane &q, double out[2]) {
  V3 v = cross(plane_normal(p), plane_normal(q));
  double vn = norm(v);
  V3 axis = v3(v[0] / vn, v[1] / vn, v[2] / vn);
  V3 b = angle_axis_matrix(M_PI / 2, axis) * plane_normal(p);
  M3 R = transpose(plane_rotation(b));
  V3 n = R * plane_normal(q);
  out[0] = azimuth(n);
  out[1] = elevation(n);
}
// ominus_par: :108-117
inline void plane_ominus_par(const Plane &p, const Plane &q, double out[2]) {
  V3 nor = plane_normal(p);
  if (dot(plane_normal(q), nor) < 0) nor = -1.0 * nor;
  M3 R = transpose(plane_rotation(nor));
  V3 n = R * plane_normal(q);
  out[0] = azimuth(n);
  out[1] = elevation(n);
}
// operator*(Isometry3D, Plane3D): :131-140 ; Isometry from SE3Quat: se3quat.h:341-346
inline Plane plane_transform(const SE3 &T, const Plane &pl) {
  M3 R = quat_to_matrix(T.r);
  V3 n2 = R * plane_normal(pl);
  double v2[4] = {n2[0], n2[1], n2[2], pl.c[3] - dot(T.t, n2)};
  if (v2[3] < 0.0)
    for (int i = 0; i < 4; i++) v2[i] = -v2[i];
  return plane_from_vector(v2);
}

// ---- g2o::cuboid (include/g2o_cuboid.h, src/g2o_cuboid.cc) -------------------------------------
// exptwist_norollpitch: g2o_cuboid.cc:6-36
inline SE3 exptwist_norollpitch(const double u[6]) {
  V3 omega = v3(u[0], u[1], u[2]);
  V3 upsilon = v3(u[3], u[4], u[5]);
  double theta = norm(omega);
  M3 Omega = skew(omega);
  M3 R;
  std::memset(&R, 0, sizeof R);
  R(0, 0) = std::cos(omega[2]);
  R(0, 1) = -std::sin(omega[2]);
  R(1, 0) = std::sin(omega[2]);
  R(1, 1) = std::cos(omega[2]);
  R(2, 2) = 1;
  M3 V;
  if (theta < 0.00001) {
    V = R;
  } else {
    M3 Omega2 = Omega * Omega;
    V = add_scaled(add_scaled(m3_identity(), (1 - std::cos(theta)) / (theta * theta), Omega),
                   (theta - std::sin(theta)) / (std::pow(theta, 3)), Omega2);
  }
  return se3_from_qt(quat_from_matrix(R), V * upsilon);
}
// VertexCuboid::oplusImpl: g2o_cuboid.cc:39-67 (whether_fixrotation is never set by the BA)
inline Cuboid cuboid_oplus(const Cuboid &est, unsigned flags, const double u[9]) {
  Cuboid nc;
  if (flags & 1u) {  // whether_fixrollpitch
    double u2[6] = {0, 0, u[2], u[3], u[4], u[5]};
    nc.pose = se3_mul(est.pose, exptwist_norollpitch(u2));
  } else {
    nc.pose = se3_mul(est.pose, se3_exp(u));
  }
  if (flags & 2u)  // whether_fixheight: keep the previous translation y
    nc.pose.t = v3(nc.pose.t[0], est.pose.t[1], nc.pose.t[2]);
  nc.scale = v3(est.scale[0] + u[6], est.scale[1] + u[7], est.scale[2] + u[8]);
  return nc;
}
// compute3D_BoxCorner: g2o_cuboid.h:198-207 (similarityTransform :173-179)
inline void cuboid_corners(const Cuboid &c, double out[3][8]) {
  static const double sgn[3][8] = {{1, 1, -1, -1, 1, 1, -1, -1}, {1, -1, -1, 1, 1, -1, -1, 1}, {-1, -1, -1, -1, 1, 1, 1, 1}};
  M3 R = quat_to_matrix(c.pose.r);
  for (int k = 0; k < 8; k++) {
    // [R*diag(scale) | t] * [corner;1], then homo_to_real divides by the homogeneous 1
    for (int i = 0; i < 3; i++) {
      double acc = 0;
      for (int j = 0; j < 3; j++) acc += (R(i, j) * c.scale[j]) * sgn[j][k];
      acc += c.pose.t[i] * 1.0;
      out[i][k] = acc / 1.0;
    }
  }
}
// projectOntoImage: g2o_cuboid.h:210-215 ; K = [fx 0 cx; 0 fy cy; 0 0 1]
inline void cuboid_project(const Cuboid &c, const SE3 &Tcw, const double K[9], double out[2][8]) {
  double cw[3][8];
  cuboid_corners(c, cw);
  M3 R = quat_to_matrix(Tcw.r);
  for (int k = 0; k < 8; k++) {
    double pc[3];
    for (int i = 0; i < 3; i++) pc[i] = R(i, 0) * cw[0][k] + R(i, 1) * cw[1][k] + R(i, 2) * cw[2][k] + Tcw.t[i];
    // homo_to_real (4 -> 3) divides by 1; K * p; homo_to_real (3 -> 2)
    double h[3];
    for (int i = 0; i < 3; i++) h[i] = K[3 * i + 0] * pc[0] + K[3 * i + 1] * pc[1] + K[3 * i + 2] * pc[2];
    out[0][k] = h[0] / h[2];
    out[1][k] = h[1] / h[2];
  }
}
// projectOntoImageBbox: g2o_cuboid.h:218-234  -> [cx cy w h]
inline void cuboid_project_bbox(const Cuboid &c, const SE3 &Tcw, const double K[9], double out[4]) {
  double p[2][8];
  cuboid_project(c, Tcw, K, p);
  double mn[2], mx[2];
  for (int r = 0; r < 2; r++) {
    mn[r] = mx[r] = p[r][0];
    for (int k = 1; k < 8; k++) {
      if (p[r][k] < mn[r]) mn[r] = p[r][k];
      if (p[r][k] > mx[r]) mx[r] = p[r][k];
    }
  }
  out[0] = (mx[0] + mn[0]) / 2;
  out[1] = (mx[1] + mn[1]) / 2;
  out[2] = mx[0] - mn[0];
  out[3] = mx[1] - mn[1];
}
// point_boundary_error: g2o_cuboid.h:237-255
inline V3 cuboid_point_boundary_error(const Cuboid &c, const V3 &pt, double max_outside_margin_ratio) {
  V3 lp = se3_map(se3_inverse(c.pose), pt);
  V3 e;
  for (int i = 0; i < 3; i++) {
    double a = std::fabs(lp[i]);
    if (a < c.scale[i])
      e[i] = 0;
    else if (a < (max_outside_margin_ratio + 1) * c.scale[i])
      e[i] = a - c.scale[i];
    else
      e[i] = max_outside_margin_ratio * c.scale[i];
  }
  return e;
}
// toMinimalVector: g2o_cuboid.h:145-163
// SE3Quat::log  Thirdparty/g2o/g2o/types/se3quat.h:229-264 (deltaR: se3_ops.hpp:40-47)
inline void se3_log(const SE3 &T, double out[6]) {
  const M3 R = quat_to_matrix(T.r);
  const double d = 0.5 * (R.m[0][0] + R.m[1][1] + R.m[2][2] - 1);
  const V3 dR = v3(R.m[2][1] - R.m[1][2], R.m[0][2] - R.m[2][0], R.m[1][0] - R.m[0][1]);
  V3 omega;
  M3 V_inv;
  if (d > 0.99999) {
    omega = 0.5 * dR;
    const M3 Om = skew(omega);
    V_inv = add_scaled(add_scaled(m3_identity(), -0.5, Om), 1. / 12., Om * Om);
  } else {
    const double theta = std::acos(d);
    omega = (theta / (2 * std::sqrt(1 - d * d))) * dR;
    const M3 Om = skew(omega);
    V_inv = add_scaled(add_scaled(m3_identity(), -0.5, Om), (1 - theta / (2 * std::tan(theta / 2))) / (theta * theta), Om * Om);
  }
  const V3 ups = V_inv * T.t;
  for (int i = 0; i < 3; i++) out[i] = omega[i], out[3 + i] = ups[i];
}
// cuboid::rotate_cuboid  include/g2o_cuboid.h:112-122
inline Cuboid cuboid_rotate(const Cuboid &c, double yaw_angle) {
  Cuboid res;
  SE3 rot = se3_from_qt(Quat{0, 0, std::sin(yaw_angle * 0.5), std::cos(yaw_angle * 0.5)}, v3(0, 0, 0));
  res.pose = se3_mul(c.pose, rot);
  res.scale = c.scale;
  if ((yaw_angle == M_PI / 2.0) || (yaw_angle == -M_PI / 2.0) || (yaw_angle == 3 * M_PI / 2.0)) std::swap(res.scale[0], res.scale[1]);
  return res;
}
// cuboid::cube_log_error :72-79
inline void cuboid_log_error(const Cuboid &self, const Cuboid &newone, double res[9]) {
  const SE3 pose_diff = se3_mul(se3_inverse(newone.pose), self.pose);
  se3_log(pose_diff, res);
  for (int i = 0; i < 3; i++) res[6 + i] = self.scale[i] - newone.scale[i];
}
// cuboid::min_log_error :82-109 (whether_rotate_cubes = true): the error of the yaw rotation (-90, 0, 90, 180 degrees) of smallest norm
inline void cuboid_min_log_error(const Cuboid &self, const Cuboid &newone, double res[9]) {
  const double angles[4] = {-1, 0, 1, 2};
  double best = 0;
  for (int i = 0; i < 4; i++) {
    double e[9], n2 = 0;
    cuboid_log_error(self, cuboid_rotate(newone, angles[i] * M_PI / 2.0), e);
    for (int k = 0; k < 9; k++) n2 += e[k] * e[k];
    const double n = std::sqrt(n2);
    if (i == 0 || n < best) {  // Eigen minCoeff: the first of equal minima
      best = n;
      for (int k = 0; k < 9; k++) res[k] = e[k];
    }
  }
}
// EdgeSE3Cuboid::computeError :330-340: the measured cuboid (camera frame) moved to the world frame with Twc = Tcw^-1 (cuboid::transform_from :125-131)
inline void cuboid_se3_error(const SE3 &Tcw, const Cuboid &global_cube, const Cuboid &meas, double err[9]) {
  Cuboid esti;
  esti.pose = se3_mul(se3_inverse(Tcw), meas.pose);
  esti.scale = meas.scale;
  cuboid_min_log_error(global_cube, esti, err);
}
inline void cuboid_to_minimal(const Cuboid &c, double v[9]) {
  const Quat &q = c.pose.r;
  v[0] = c.pose.t[0];
  v[1] = c.pose.t[1];
  v[2] = c.pose.t[2];
  v[3] = std::atan2(2 * (q.w * q.x + q.y * q.z), 1 - 2 * (q.x * q.x + q.y * q.y));
  v[4] = std::asin(2 * (q.w * q.y - q.z * q.x));
  v[5] = std::atan2(2 * (q.w * q.z + q.x * q.y), 1 - 2 * (q.y * q.y + q.z * q.z));
  v[6] = c.scale[0];
  v[7] = c.scale[1];
  v[8] = c.scale[2];
}

// ---- RobustKernelHuber::robustify: core/robust_kernel_impl.cpp:76-90 ---------------------------
inline void huber(double e, double delta, double rho[3]) {
  // "float dsqr;" in the reference's (ORB-SLAM2-modified) RobustKernelHuber, core/robust_kernel_impl.h:84: delta^2 is rounded to
  // float32 at setDelta() (robust_kernel_impl.cpp:64-68) and that rounded value is what the inlier test and rho(e) use
  double dsqr = (double)(float)(delta * delta);
  if (e <= dsqr) {
    rho[0] = e;
    rho[1] = 1.;
    rho[2] = 0.;
  } else {
    double sqrte = std::sqrt(e);
    rho[0] = 2 * sqrte * delta - dsqr;
    rho[1] = delta / sqrte;
    rho[2] = -0.5 * rho[1] / e;
  }
}

}  // namespace ppo_oracle
