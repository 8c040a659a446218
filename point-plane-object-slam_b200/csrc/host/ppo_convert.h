// Host-side converters between the reference's float32 map state and the engine's f64 graph:
// the equivalents of ORB_SLAM2::Converter used by the BA (src/Converter.cc:37-53,63-90,110-116,
// 172-190).  Used by the Optimizer shim (host/ppo_optimizer_shim.*) and the synthetic map
// generator.  Plain C++, no dependencies.
#pragma once
#include <cmath>

namespace ppo {

// Converter::toSE3Quat (Converter.cc:37-47): float 4x4 Tcw (row-major) -> [qx qy qz qw tx ty tz].
// The float rotation is not exactly orthonormal; g2o builds the quaternion with Eigen's
// Quaterniond(Matrix3d) (trace / largest-diagonal branches) and then normalises with w >= 0
// (se3quat.h:58-60,331-336).
inline void tcw_float_to_pose7(const float T[16], double out[7]) {
  double m[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) m[i][j] = (double)T[4 * i + j];
  double q[4];
  double tr = m[0][0] + m[1][1] + m[2][2];
  if (tr > 0.0) {
    double s = std::sqrt(tr + 1.0);
    q[3] = 0.5 * s;
    s = 0.5 / s;
    q[0] = (m[2][1] - m[1][2]) * s;
    q[1] = (m[0][2] - m[2][0]) * s;
    q[2] = (m[1][0] - m[0][1]) * s;
  } else {
    int i = 0;
    if (m[1][1] > m[0][0]) i = 1;
    if (m[2][2] > m[i][i]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    double s = std::sqrt(m[i][i] - m[j][j] - m[k][k] + 1.0);
    q[i] = 0.5 * s;
    s = 0.5 / s;
    q[3] = (m[k][j] - m[j][k]) * s;
    q[j] = (m[j][i] + m[i][j]) * s;
    q[k] = (m[k][i] + m[i][k]) * s;
  }
  if (q[3] < 0)
    for (int i = 0; i < 4; i++) q[i] = -q[i];
  double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int i = 0; i < 4; i++) out[i] = q[i] / n;
  out[4] = (double)T[3];
  out[5] = (double)T[7];
  out[6] = (double)T[11];
}

// Converter::toCvMat(SE3Quat) (Converter.cc:49-53,63-71): pose7 -> float 4x4 row-major
inline void pose7_to_tcw_float(const double p[7], float T[16]) {
  const double x = p[0], y = p[1], z = p[2], w = p[3];
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y,
               tzz = tz * z;
  T[0] = (float)(1 - (tyy + tzz));
  T[1] = (float)(txy - twz);
  T[2] = (float)(txz + twy);
  T[3] = (float)p[4];
  T[4] = (float)(txy + twz);
  T[5] = (float)(1 - (txx + tzz));
  T[6] = (float)(tyz - twx);
  T[7] = (float)p[5];
  T[8] = (float)(txz - twy);
  T[9] = (float)(tyz + twx);
  T[10] = (float)(1 - (txx + tyy));
  T[11] = (float)p[6];
  T[12] = T[13] = T[14] = 0.f;
  T[15] = 1.f;
}

// Converter::toPlane3D (Converter.cc:172-181): float 4x1 -> Plane3D coefficients (sign flip to
// d >= 0, then Plane3D::fromVector normalises, G2O_Plane3D.h:45-48,120-125)
inline void plane_float_to_coef(const float c[4], double out[4]) {
  double v[4] = {(double)c[0], (double)c[1], (double)c[2], (double)c[3]};
  if (c[3] < 0.0f)
    for (int i = 0; i < 4; i++) v[i] = -v[i];
  double n = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  double inv = 1. / n;
  for (int i = 0; i < 4; i++) out[i] = v[i] * inv;
  if (out[3] < 0.0)
    for (int i = 0; i < 4; i++) out[i] = -out[i];
}

// (double)(float)sqrt(th): how Optimizer.cc forms every Huber delta ("const float th = sqrt(..)")
inline double huber_delta(double chi2_threshold) { return (double)(float)std::sqrt(chi2_threshold); }

}  // namespace ppo
