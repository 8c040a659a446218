"""Independent numpy transliteration (rotation-matrix based, written from the formulas in
SURVEY.md Appendix A, not from oracle/) of the reference's vertex/edge maths.  Second opinion for
the oracle's known-answer tests."""
import numpy as np


def skew(w):
    return np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0.0]])


def quat_to_R(q):  # q = [x y z w]
    x, y, z, w = q
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def pose_to_Rt(p):
    return quat_to_R(p[:4]), np.asarray(p[4:7], float)


def se3_exp(u):
    """se3quat.h:274-308: returns (R, t)."""
    w, v = np.asarray(u[:3], float), np.asarray(u[3:], float)
    th = np.linalg.norm(w)
    W = skew(w)
    if th < 1e-5:
        R = np.eye(3) + W + W @ W
        V = R
    else:
        R = np.eye(3) + np.sin(th) / th * W + (1 - np.cos(th)) / th**2 * W @ W
        V = np.eye(3) + (1 - np.cos(th)) / th**2 * W + (th - np.sin(th)) / th**3 * W @ W
    return R, V @ v


def project(p, intr, obs):
    """mono / stereo reprojection residual, types_six_dof_expmap.cpp:172-189."""
    fx, fy, cx, cy, bf = [float(v) for v in intr]
    u = fx * p[0] / p[2] + cx
    v = fy * p[1] / p[2] + cy
    if obs[2] < 0:
        return np.array([obs[0] - u, obs[1] - v])
    return np.array([obs[0] - u, obs[1] - v, obs[2] - (u - bf / p[2])])


def plane_normalize(c):
    c = np.asarray(c, float) / np.linalg.norm(c[:3])
    return -c if c[3] < 0 else c


def az(v):
    return np.arctan2(v[1], v[0])


def el(v):
    return np.arctan2(v[2], np.hypot(v[0], v[1]))


def Rz(a):
    return np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1.0]])


def Ry(a):
    return np.array([[np.cos(a), 0, np.sin(a)], [0, 1.0, 0], [-np.sin(a), 0, np.cos(a)]])


def plane_rot(n):  # G2O_Plane3D.h:66-72
    return Rz(az(n)) @ Ry(-el(n))


def plane_oplus(c, v):
    c = plane_normalize(c)
    n = np.array([np.cos(v[1]) * np.cos(v[0]), np.cos(v[1]) * np.sin(v[0]), np.sin(v[1])])
    d = -c[3] + v[2]
    return plane_normalize(np.r_[plane_rot(c[:3]) @ n, -d])


def plane_ominus(a, b):
    a, b = plane_normalize(a), plane_normalize(b)
    n = plane_rot(a[:3]).T @ b[:3]
    return np.array([az(n), el(n), -a[3] + b[3]])


def axis_angle_R(axis, ang):
    K = skew(axis)
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K


def plane_ominus_ver(a, b):
    a, b = plane_normalize(a), plane_normalize(b)
    v = np.cross(a[:3], b[:3])
    bb = axis_angle_R(v / np.linalg.norm(v), np.pi / 2) @ a[:3]
    n = plane_rot(bb).T @ b[:3]
    return np.array([az(n), el(n)])


def plane_ominus_par(a, b):
    a, b = plane_normalize(a), plane_normalize(b)
    nor = a[:3] if b[:3] @ a[:3] >= 0 else -a[:3]
    n = plane_rot(nor).T @ b[:3]
    return np.array([az(n), el(n)])


def plane_transform(pose, c):
    R, t = pose_to_Rt(pose)
    c = plane_normalize(c)
    n2 = R @ c[:3]
    d2 = c[3] - t @ n2
    v = np.r_[n2, d2]
    return plane_normalize(-v if d2 < 0 else v)


SGN = np.array([[1, 1, -1, -1, 1, 1, -1, -1], [1, -1, -1, 1, 1, -1, -1, 1], [-1, -1, -1, -1, 1, 1, 1, 1.0]])


def cuboid_corners(c):
    t, q, s = np.asarray(c[:3]), c[3:7], np.asarray(c[7:10])
    return quat_to_R(q) @ (s[:, None] * SGN) + t[:, None]


def cuboid_project(c, pose, intr):
    R, t = pose_to_Rt(pose)
    pc = R @ cuboid_corners(c) + t[:, None]
    fx, fy, cx, cy = [float(v) for v in intr[:4]]
    return np.stack([fx * pc[0] / pc[2] + cx, fy * pc[1] / pc[2] + cy])


def cuboid_bbox(c, pose, intr):
    p = cuboid_project(c, pose, intr)
    mn, mx = p.min(axis=1), p.max(axis=1)
    return np.r_[(mn + mx) / 2, mx - mn]


def cuboid_oplus_yaw(c, u):
    """fixrollpitch + fixheight update (g2o_cuboid.cc:6-67); returns (R, t, scale)."""
    R0, t0, s0 = quat_to_R(c[3:7]), np.asarray(c[:3], float), np.asarray(c[7:10], float)
    th = abs(u[2])
    Rd = Rz(u[2])
    W = skew([0, 0, u[2]])
    V = Rd if th < 1e-5 else np.eye(3) + (1 - np.cos(th)) / th**2 * W + (th - np.sin(th)) / th**3 * W @ W
    t = t0 + R0 @ (V @ np.asarray(u[3:6], float))
    t[1] = t0[1]
    return R0 @ Rd, t, s0 + np.asarray(u[6:9], float)


def point_cuboid_error(c, pts, ratio=1.0, prior=0.2):
    R, t, s = quat_to_R(c[3:7]), np.asarray(c[:3]), np.asarray(c[7:10])
    lp = np.abs((np.asarray(pts) - t) @ R)  # R^T (p - t)
    e = np.where(lp < s, 0.0, np.where(lp < (ratio + 1) * s, lp - s, ratio * s))
    return e.mean(axis=0) / s + prior * s


def huber(e, delta):
    dsqr = float(np.float32(delta * delta))  # "float dsqr" of the reference's RobustKernelHuber (core/robust_kernel_impl.h:84)
    if e <= dsqr:
        return e, 1.0
    return 2 * np.sqrt(e) * delta - dsqr, delta / np.sqrt(e)
