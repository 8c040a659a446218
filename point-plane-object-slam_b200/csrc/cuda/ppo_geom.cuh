// Device-side geometry of the local BA: SE3 (quaternion + translation), Plane3D, 9-DoF cuboid and
// the residual functions of the eight edge types.  Written for the GPU (flat double arrays in
// registers, no structs of matrices); semantics follow the reference lines cited at each function
// (paths relative to the reference repository).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#define PPO_D __device__ __forceinline__

namespace ppo {

// ---------------------------------------------------------------------------------------------
// quaternions are [x y z w]; rotation matrices row-major double[9]
// ---------------------------------------------------------------------------------------------
PPO_D void quat_to_R(const double q[4], double R[9]) {  // Eigen QuaternionBase::toRotationMatrix
  const double tx = 2.0 * q[0], ty = 2.0 * q[1], tz = 2.0 * q[2];
  const double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
  const double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
  const double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
  R[0] = 1.0 - (tyy + tzz); R[1] = txy - twz;         R[2] = txz + twy;
  R[3] = txy + twz;         R[4] = 1.0 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;         R[7] = tyz + twx;         R[8] = 1.0 - (txx + tyy);
}
PPO_D void quat_from_R(const double m[9], double q[4]) {  // Eigen Quaterniond(Matrix3d)
  double t = m[0] + m[4] + m[8];
  if (t > 0.0) {
    t = sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (m[7] - m[5]) * t;
    q[1] = (m[2] - m[6]) * t;
    q[2] = (m[3] - m[1]) * t;
  } else {
    int i = 0;
    if (m[4] > m[0]) i = 1;
    if (m[8] > m[4 * i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(m[4 * i] - m[4 * j] - m[4 * k] + 1.0);
    double qq[3];
    qq[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (m[3 * k + j] - m[3 * j + k]) * t;
    qq[j] = (m[3 * j + i] + m[3 * i + j]) * t;
    qq[k] = (m[3 * k + i] + m[3 * i + k]) * t;
    q[0] = qq[0]; q[1] = qq[1]; q[2] = qq[2];
  }
}
PPO_D void quat_normalize_pos(double q[4]) {  // SE3Quat::normalizeRotation, se3quat.h:331-336
  double s = (q[3] < 0) ? -1.0 : 1.0;
  const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  s /= n;
  q[0] *= s; q[1] *= s; q[2] *= s; q[3] *= s;
}
PPO_D void quat_mul(const double a[4], const double b[4], double r[4]) {
  r[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  r[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  r[1] = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
  r[2] = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
}
PPO_D void quat_rot(const double q[4], const double v[3], double o[3]) {  // v + w*2(u x v) + u x 2(u x v)
  const double ux = 2 * (q[1] * v[2] - q[2] * v[1]);
  const double uy = 2 * (q[2] * v[0] - q[0] * v[2]);
  const double uz = 2 * (q[0] * v[1] - q[1] * v[0]);
  o[0] = v[0] + q[3] * ux + (q[1] * uz - q[2] * uy);
  o[1] = v[1] + q[3] * uy + (q[2] * ux - q[0] * uz);
  o[2] = v[2] + q[3] * uz + (q[0] * uy - q[1] * ux);
}
PPO_D void mat3_vec(const double R[9], const double v[3], double o[3]) {
  o[0] = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
  o[1] = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
  o[2] = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
}
PPO_D void mat3T_vec(const double R[9], const double v[3], double o[3]) {
  o[0] = R[0] * v[0] + R[3] * v[1] + R[6] * v[2];
  o[1] = R[1] * v[0] + R[4] * v[1] + R[7] * v[2];
  o[2] = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
}

// ---------------------------------------------------------------------------------------------
// SE3: pose7 = [qx qy qz qw tx ty tz]
// ---------------------------------------------------------------------------------------------
// rotation part of SE3Quat::exp (se3quat.h:274-308) and the V matrix applied to upsilon.
// w = omega, returns R (row-major) and t = V * upsilon.
PPO_D void so3_exp_V(const double w[3], const double ups[3], double R[9], double t[3], bool yaw_only_R) {
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  const double th = sqrt(th2);
  // Omega and Omega^2
  const double O[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
  double O2[9];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) O2[3 * i + j] = O[3 * i] * O[j] + O[3 * i + 1] * O[3 + j] + O[3 * i + 2] * O[6 + j];
  double V[9];
  if (yaw_only_R) {  // exptwist_norollpitch, src/g2o_cuboid.cc:6-36: R = Rz(omega_z)
    double s, c;
    sincos(w[2], &s, &c);
    R[0] = c; R[1] = -s; R[2] = 0; R[3] = s; R[4] = c; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
  }
  if (th < 0.00001) {
    if (!yaw_only_R) {
#pragma unroll
      for (int i = 0; i < 9; i++) R[i] = ((i % 4) == 0 ? 1.0 : 0.0) + O[i] + O2[i];
    }
#pragma unroll
    for (int i = 0; i < 9; i++) V[i] = R[i];
  } else {
    double s, c;
    sincos(th, &s, &c);
    const double a = s / th, b = (1 - c) / th2, d = (th - s) / (th2 * th);
    if (!yaw_only_R) {
#pragma unroll
      for (int i = 0; i < 9; i++) R[i] = ((i % 4) == 0 ? 1.0 : 0.0) + a * O[i] + b * O2[i];
    }
#pragma unroll
    for (int i = 0; i < 9; i++) V[i] = ((i % 4) == 0 ? 1.0 : 0.0) + b * O[i] + d * O2[i];
  }
  mat3_vec(V, ups, t);
}
// VertexSE3Expmap::oplusImpl (types_six_dof_expmap.h:88-91): est <- exp(u) * est
PPO_D void se3_oplus(const double p[7], const double u[6], double o[7]) {
  double R[9], t[3], qe[4];
  so3_exp_V(u, u + 3, R, t, false);
  quat_from_R(R, qe);
  quat_normalize_pos(qe);
  double rt[3];
  quat_rot(qe, p + 4, rt);
  quat_mul(qe, p, o);
  quat_normalize_pos(o);
  o[4] = t[0] + rt[0];
  o[5] = t[1] + rt[1];
  o[6] = t[2] + rt[2];
}
// Rt12 = [R row-major (9) | t (3)] cache of a pose
PPO_D void pose_to_Rt(const double p[7], double Rt[12]) {
  quat_to_R(p, Rt);
  Rt[9] = p[4]; Rt[10] = p[5]; Rt[11] = p[6];
}

// ---------------------------------------------------------------------------------------------
// Plane3D (include/G2O_Plane3D.h)
// ---------------------------------------------------------------------------------------------
PPO_D void plane_normalize(double c[4]) {  // :120-125
  const double inv = 1.0 / sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
  c[0] *= inv; c[1] *= inv; c[2] *= inv; c[3] *= inv;
  if (c[3] < 0.0) { c[0] = -c[0]; c[1] = -c[1]; c[2] = -c[2]; c[3] = -c[3]; }
}
// rotation(v) = Rz(azimuth) * Ry(-elevation) built, as Eigen does, from the product of the two
// half-angle quaternions (:66-72)
PPO_D void plane_rotation(const double v[3], double R[9]) {
  const double az = atan2(v[1], v[0]);
  const double el = atan2(v[2], sqrt(v[0] * v[0] + v[1] * v[1]));
  double sa, ca, se, ce;
  sincos(0.5 * az, &sa, &ca);
  sincos(-0.5 * el, &se, &ce);
  const double qa[4] = {0, 0, sa, ca}, qe[4] = {0, se, 0, ce};
  double q[4];
  quat_mul(qa, qe, q);
  quat_to_R(q, R);
}
PPO_D void plane_oplus(const double c[4], const double v[3], double o[4]) {  // :74-87
  double s, co, sa, ca;
  sincos(v[1], &s, &co);
  sincos(v[0], &sa, &ca);
  const double n[3] = {co * ca, co * sa, s};
  double R[9];
  plane_rotation(c, R);
  const double d = -c[3] + v[2];
  mat3_vec(R, n, o);
  o[3] = -d;
  plane_normalize(o);
}
// T * plane (:131-140) with T given as Rt12
PPO_D void plane_transform(const double Rt[12], const double c[4], double o[4]) {
  mat3_vec(Rt, c, o);
  o[3] = c[3] - (Rt[9] * o[0] + Rt[10] * o[1] + Rt[11] * o[2]);
  if (o[3] < 0.0) { o[0] = -o[0]; o[1] = -o[1]; o[2] = -o[2]; o[3] = -o[3]; }
  plane_normalize(o);
}
PPO_D void az_el(const double R[9], const double m[3], double *az, double *el) {  // of R^T m
  double n[3];
  mat3T_vec(R, m, n);
  *az = atan2(n[1], n[0]);
  *el = atan2(n[2], sqrt(n[0] * n[0] + n[1] * n[1]));
}
// residual of EdgePlane / EdgeVerticalPlane / EdgeParallelPlane (:181-193, :220-232, :279-291):
// local = Tcw * plane ; err = local.ominus{,_ver,_par}(meas)
PPO_D int plane_edge_error(int kind, const double pl[4], const double Rt[12], const double meas[4], double err[3]) {
  double l[4], R[9];
  plane_transform(Rt, pl, l);
  err[2] = 0.0;
  if (kind == 0) {  // ominus :89-95
    plane_rotation(l, R);
    az_el(R, meas, &err[0], &err[1]);
    err[2] = (-l[3]) - (-meas[3]);
    return 3;
  }
  if (kind == 1) {  // ominus_ver :97-106: rotate the local normal by 90 deg about n x m
    double v[3] = {l[1] * meas[2] - l[2] * meas[1], l[2] * meas[0] - l[0] * meas[2], l[0] * meas[1] - l[1] * meas[0]};
    const double vn = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    v[0] /= vn; v[1] /= vn; v[2] /= vn;
    // AngleAxis(pi/2, v).toRotationMatrix() * n
    double s, c;
    sincos(M_PI / 2, &s, &c);
    const double c1 = 1.0 - c;
    double A[9];
    A[0] = c1 * v[0] * v[0] + c;        A[1] = c1 * v[0] * v[1] - s * v[2]; A[2] = c1 * v[0] * v[2] + s * v[1];
    A[3] = c1 * v[0] * v[1] + s * v[2]; A[4] = c1 * v[1] * v[1] + c;        A[5] = c1 * v[1] * v[2] - s * v[0];
    A[6] = c1 * v[0] * v[2] - s * v[1]; A[7] = c1 * v[1] * v[2] + s * v[0]; A[8] = c1 * v[2] * v[2] + c;
    double b[3];
    mat3_vec(A, l, b);
    plane_rotation(b, R);
    az_el(R, meas, &err[0], &err[1]);
    return 2;
  }
  // ominus_par :108-117
  double nor[3] = {l[0], l[1], l[2]};
  if (meas[0] * nor[0] + meas[1] * nor[1] + meas[2] * nor[2] < 0) { nor[0] = -nor[0]; nor[1] = -nor[1]; nor[2] = -nor[2]; }
  plane_rotation(nor, R);
  az_el(R, meas, &err[0], &err[1]);
  return 2;
}

// ---------------------------------------------------------------------------------------------
// cuboid: c10 = [tx ty tz qx qy qz qw sx sy sz]   (include/g2o_cuboid.h, src/g2o_cuboid.cc)
// ---------------------------------------------------------------------------------------------
// VertexCuboid::oplusImpl src/g2o_cuboid.cc:39-67
PPO_D void cuboid_oplus(const double c[10], unsigned flags, const double u[9], double o[10]) {
  double R[9], t[3], qd[4];
  if (flags & 1u) {
    const double w[3] = {0.0, 0.0, u[2]};
    so3_exp_V(w, u + 3, R, t, true);
  } else {
    so3_exp_V(u, u + 3, R, t, false);
  }
  quat_from_R(R, qd);
  quat_normalize_pos(qd);
  double rt[3];
  quat_rot(c + 3, t, rt);  // pose * delta: t' = t + R * t_delta ; q' = q * q_delta
  quat_mul(c + 3, qd, o + 3);
  quat_normalize_pos(o + 3);
  o[0] = c[0] + rt[0];
  o[1] = (flags & 2u) ? c[1] : c[1] + rt[1];
  o[2] = c[2] + rt[2];
  o[7] = c[7] + u[6];
  o[8] = c[8] + u[7];
  o[9] = c[9] + u[8];
}
// corner k of compute3D_BoxCorner (g2o_cuboid.h:198-207) in the camera frame, projected with K
// (projectOntoImage :210-215).  Rc = cuboid rotation, Rt = camera pose cache.
PPO_D void cuboid_corner_px(const double Rc[9], const double c[10], const double Rt[12], const double fx, const double fy,
                            const double cx, const double cy, int k, double *u, double *v) {
  const double sx = ((k & 3) == 0 || (k & 3) == 1) ? 1.0 : -1.0;  // 1 1 -1 -1 ...
  const double sy = ((k & 3) == 0 || (k & 3) == 3) ? 1.0 : -1.0;  // 1 -1 -1 1 ...
  const double sz = (k < 4) ? -1.0 : 1.0;
  const double l[3] = {c[7] * sx, c[8] * sy, c[9] * sz};
  double w[3], p[3];
  mat3_vec(Rc, l, w);
  w[0] += c[0]; w[1] += c[1]; w[2] += c[2];
  mat3_vec(Rt, w, p);
  p[0] += Rt[9]; p[1] += Rt[10]; p[2] += Rt[11];
  // K * p then divide by the third homogeneous coordinate (= z)
  *u = (fx * p[0] + cx * p[2]) / p[2];
  *v = (fy * p[1] + cy * p[2]) / p[2];
}
// EdgeSE3CuboidProj (4-D, g2o_cuboid.cc:70-80) / EdgeSE3CuboidCornerProj (16-D, :103-120)
PPO_D int cuboid_cam_error(int kind, const double Rt[12], const double c[10], const float intr[5], const double *meas,
                           double err[16]) {
  double Rc[9];
  quat_to_R(c + 3, Rc);
  const double fx = intr[0], fy = intr[1], cx = intr[2], cy = intr[3];
  if (kind == 0) {
    double mnu = 1e300, mnv = 1e300, mxu = -1e300, mxv = -1e300;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      double u, v;
      cuboid_corner_px(Rc, c, Rt, fx, fy, cx, cy, k, &u, &v);
      mnu = fmin(mnu, u); mxu = fmax(mxu, u);
      mnv = fmin(mnv, v); mxv = fmax(mxv, v);
    }
    err[0] = (mxu + mnu) / 2 - meas[0];
    err[1] = (mxv + mnv) / 2 - meas[1];
    err[2] = (mxu - mnu) - meas[2];
    err[3] = (mxv - mnv) - meas[3];
    return 4;
  }
#pragma unroll
  for (int k = 0; k < 8; k++) {
    double u, v;
    cuboid_corner_px(Rc, c, Rt, fx, fy, cx, cy, k, &u, &v);
    err[2 * k] = u - meas[2 * k];
    err[2 * k + 1] = v - meas[2 * k + 1];
  }
  return 16;
}
// ---- EdgeSE3Cuboid (include/g2o_cuboid.h:322-340): 9-D error between the cuboid vertex and the measured cuboid moved to the world ----
// SE3Quat product / inverse with normalizeRotation (se3quat.h:110-134); poses as (q[4] = x y z w, t[3])
PPO_D void se3q_mul(const double qa[4], const double ta[3], const double qb[4], const double tb[3], double q[4], double t[3]) {
  double r[3];
  quat_rot(qa, tb, r);
  t[0] = ta[0] + r[0], t[1] = ta[1] + r[1], t[2] = ta[2] + r[2];
  quat_mul(qa, qb, q);
  quat_normalize_pos(q);
}
PPO_D void se3q_inv(const double q[4], const double t[3], double qi[4], double ti[3]) {
  qi[0] = -q[0], qi[1] = -q[1], qi[2] = -q[2], qi[3] = q[3];
  const double nt[3] = {-t[0], -t[1], -t[2]};
  quat_rot(qi, nt, ti);
}
// SE3Quat::log (se3quat.h:229-264)
PPO_D void se3q_log(const double q[4], const double t[3], double out[6]) {
  double R[9];
  quat_to_R(q, R);
  const double d = 0.5 * (R[0] + R[4] + R[8] - 1);
  const double dR[3] = {R[7] - R[5], R[2] - R[6], R[3] - R[1]};
  double w[3], k2;  // V_inv = I - 0.5 Omega + k2 Omega^2
  if (d > 0.99999) {
    w[0] = 0.5 * dR[0], w[1] = 0.5 * dR[1], w[2] = 0.5 * dR[2];
    k2 = 1. / 12.;
  } else {
    const double theta = acos(d), f = theta / (2 * sqrt(1 - d * d));
    w[0] = f * dR[0], w[1] = f * dR[1], w[2] = f * dR[2];
    k2 = (1 - theta / (2 * tan(theta / 2))) / (theta * theta);
  }
  // Omega = skew(w); Omega^2 as a matrix product, entry by entry like Eigen's 3 x 3 product
  const double O[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
  double V[9];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const double o2 = O[3 * i] * O[j] + O[3 * i + 1] * O[3 + j] + O[3 * i + 2] * O[6 + j];
      V[3 * i + j] = ((i == j ? 1.0 : 0.0) + -0.5 * O[3 * i + j]) + k2 * o2;
    }
  out[0] = w[0], out[1] = w[1], out[2] = w[2];
  mat3_vec(V, t, out + 3);
}
// cam = Tcw as [qx qy qz qw tx ty tz]; c, meas = cuboids as [t(3) q(4) scale(3)] (meas in the camera frame)
PPO_D void cuboid_se3_error(const double cam[7], const double c[10], const double *meas, double err[9]) {
  double qwc[4], twc[3], qe[4], te[3];
  se3q_inv(cam, cam + 4, qwc, twc);                 // Twc
  se3q_mul(qwc, twc, meas + 3, meas, qe, te);       // transform_from: Twc * meas.pose
  double best = 0;
#pragma unroll 1
  for (int i = 0; i < 4; i++) {                     // min_log_error: yaw -90, 0, 90, 180 degrees of the measured cuboid
    const double yaw = (double)(i - 1) * 3.14159265358979323846 / 2.0;
    double qz[4] = {0, 0, sin(yaw * 0.5), cos(yaw * 0.5)};
    quat_normalize_pos(qz);
    const double tz[3] = {0, 0, 0};
    double qr[4], tr[3], qi[4], ti[3], qd[4], td[3], e[9];
    se3q_mul(qe, te, qz, tz, qr, tr);               // rotate_cuboid
    const bool swap = (i == 0 || i == 2);
    const double s0 = swap ? meas[8] : meas[7], s1 = swap ? meas[7] : meas[8];
    se3q_inv(qr, tr, qi, ti);
    se3q_mul(qi, ti, c + 3, c, qd, td);             // cube_log_error: newone.pose^-1 * this.pose
    se3q_log(qd, td, e);
    e[6] = c[7] - s0, e[7] = c[8] - s1, e[8] = c[9] - meas[9];
    double n2 = 0;
#pragma unroll
    for (int k = 0; k < 9; k++) n2 += e[k] * e[k];
    const double n = sqrt(n2);
    if (i == 0 || n < best) {
      best = n;
#pragma unroll
      for (int k = 0; k < 9; k++) err[k] = e[k];
    }
  }
}
// EdgePointCuboidOnlyObject::computeError (g2o_cuboid.cc:132-160) with point_boundary_error
// (g2o_cuboid.h:237-255); prior_object_half_size is never set by the BA.
PPO_D void point_cuboid_error(const double c[10], const double *pts, int n, double ratio, double prior_w, double err[3]) {
  double qc[4] = {-c[3], -c[4], -c[5], c[6]};
  double acc[3] = {0, 0, 0};
  for (int i = 0; i < n; i++) {
    const double d[3] = {pts[3 * i] - c[0], pts[3 * i + 1] - c[1], pts[3 * i + 2] - c[2]};
    double lp[3];
    quat_rot(qc, d, lp);  // pose^-1 * p = R^T (p - t)
#pragma unroll
    for (int a = 0; a < 3; a++) {
      const double v = fabs(lp[a]), s = c[7 + a];
      double e;
      if (v < s) e = 0;
      else if (v < (ratio + 1) * s) e = v - s;
      else e = ratio * s;
      acc[a] += fabs(e);
    }
  }
#pragma unroll
  for (int a = 0; a < 3; a++) {
    double m = n > 0 ? acc[a] / n : acc[a];
    err[a] = 1.0 * (m / c[7 + a]) + prior_w * c[7 + a];
  }
}

// ---------------------------------------------------------------------------------------------
// reprojection edges
// ---------------------------------------------------------------------------------------------
// camera-frame point
PPO_D void cam_point(const double Rt[12], const double X[3], double p[3]) {
  p[0] = Rt[0] * X[0] + Rt[1] * X[1] + Rt[2] * X[2] + Rt[9];
  p[1] = Rt[3] * X[0] + Rt[4] * X[1] + Rt[5] * X[2] + Rt[10];
  p[2] = Rt[6] * X[0] + Rt[7] * X[1] + Rt[8] * X[2] + Rt[11];
}
// EdgeSE3ProjectXYZ / EdgeStereoSE3ProjectXYZ::computeError (types_six_dof_expmap.h:174-179,206-211;
// .cpp:172-189).  Stereo: invz and bf*invz are float.
PPO_D int point_edge_error(const double p[3], const float intr[5], float ou, float ov, float our, double err[3]) {
  const double fx = intr[0], fy = intr[1], cx = intr[2], cy = intr[3];
  if (our < 0.f) {
    err[0] = (double)ou - ((p[0] / p[2]) * fx + cx);
    err[1] = (double)ov - ((p[1] / p[2]) * fy + cy);
    err[2] = 0.0;
    return 2;
  }
  const float invz = (float)(1.0 / p[2]);
  const double r0 = p[0] * (double)invz * fx + cx;
  const double r1 = p[1] * (double)invz * fy + cy;
  const double r2 = r0 - (double)__fmul_rn(intr[4], invz);
  err[0] = (double)ou - r0;
  err[1] = (double)ov - r1;
  err[2] = (double)our - r2;
  return 3;
}
// Huber weight rho'(e) (robust_kernel_impl.cpp:76-90); returns rho(e) through *rho0
// (the reference keeps delta^2 in a float member, core/robust_kernel_impl.h:84: the rounded value enters the inlier test and rho)
PPO_D double huber_w(double e, double delta, double *rho0) {
  const double dsqr = (double)(float)(delta * delta);
  if (e <= dsqr) {
    *rho0 = e;
    return 1.0;
  }
  const double sq = sqrt(e);
  *rho0 = 2 * sq * delta - dsqr;
  return delta / sq;
}

}  // namespace ppo
