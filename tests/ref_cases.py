"""Seeded inputs shared by tests/golden/make_ref_golden.py (run on the compiled reference) and tests/test_ref_pin.py (run on the
oracle): every function of the formula-level C API is evaluated on the same random arguments through a library + symbol prefix."""
import ctypes as C

import numpy as np

from ppo_pkg import ppo

A = ppo.abi

WINDOWS = {
    "points_only": (dict(index=0, n_kf=6, n_fixed=2, n_pt=160), True),
    "mixed_bbox": (dict(index=1, n_kf=8, n_fixed=2, n_pt=220, n_pl=4, n_cu=3), False),
    "mixed_corners": (dict(index=1, n_kf=8, n_fixed=2, n_pt=220, n_pl=4, n_cu=3, corners_2d=1, cuboid_2d=0), False),
    # the graph of LocalBACameraPointCuboids2D (Optimizer.cc:1252-1992): no planes, bbox edges + the 9-D EdgeSE3Cuboid
    "cuboids_se3": (dict(index=1, n_kf=8, n_fixed=2, n_pt=220, n_pl=0, n_cu=3, cuboid_2d=1, cuboid_3d=1, plane_3d=0, cuboid_plane=0), False),
}
INTR = np.array([517.306408, 516.469215, 318.643040, 255.313989, 40.0], np.float32)


def _d(n):
    return (C.c_double * n)()


def _a(x, ct=C.c_double):
    x = np.asarray(x, dtype=np.float64 if ct is C.c_double else np.float32).ravel()
    return (ct * len(x))(*x)


def _pose(f, rng, scale=1.0):
    o = _d(7)
    f("se3_exp")(_a(rng.normal(size=6) * scale), o)
    return np.array(o)


def _cuboid(rng):
    yaw = rng.uniform(-np.pi, np.pi)
    return np.r_[rng.uniform(-2, 2, 3), 0.0, 0.0, np.sin(yaw / 2), np.cos(yaw / 2), rng.uniform(0.2, 0.8, 3)]


def function_vectors(lib, prefix, n=24, seed=20261017):
    """name -> array of outputs (inputs are regenerated from the seed, so only outputs are stored / compared)."""
    f = lambda name: getattr(lib, prefix + name)
    for name in ("point_edge", "plane_edge", "cuboid_cam_edge"):
        f(name).restype = C.c_int
    rng = np.random.default_rng(seed)
    out = {k: [] for k in ("se3_exp", "se3_oplus", "se3_map", "se3_matrix", "se3_from_Rt", "plane_normalize", "plane_oplus", "plane_ominus", "plane_ominus_ver",
                           "plane_ominus_par", "plane_transform", "cuboid_oplus", "cuboid_corners", "cuboid_project_corners", "cuboid_project_bbox",
                           "cuboid_point_error", "cuboid_to_minimal", "huber", "point_edge_mono", "point_edge_stereo", "plane_edge", "cuboid_cam_bbox",
                           "cuboid_cam_corner", "cuboid_cam_se3")}
    rng3 = np.random.default_rng(seed + 3)  # own stream for the vectors added later: the earlier ones keep their inputs
    for it in range(n):
        # SE3: exponential (incl. the small-angle branch at |w| < 1e-5), oplus, map, matrix, construction from a float32 rotation
        u = rng.normal(size=6) * (1e-6 if it % 6 == 0 else (3.0 if it % 6 == 1 else 0.4))
        o = _d(7)
        f("se3_exp")(_a(u), o)
        out["se3_exp"].append(np.array(o))
        pose = _pose(f, rng)
        o = _d(7)
        f("se3_oplus")(_a(pose), _a(rng.normal(size=6) * 0.1), o)
        out["se3_oplus"].append(np.array(o))
        o = _d(3)
        f("se3_map")(_a(pose), _a(rng.normal(size=3)), o)
        out["se3_map"].append(np.array(o))
        Rm = _d(9)
        f("se3_matrix")(_a(pose), Rm)
        out["se3_matrix"].append(np.array(Rm))
        R32 = np.array(Rm).astype(np.float32).astype(np.float64)  # Converter::toSE3Quat input: float32-rounded rotation
        o = _d(7)
        f("se3_from_Rt")(_a(R32), _a(rng.normal(size=3)), o)
        out["se3_from_Rt"].append(np.array(o))
        # planes
        c = np.r_[rng.normal(size=3), rng.uniform(-4, 4)]
        o = _d(4)
        f("plane_normalize")(_a(c), o)
        out["plane_normalize"].append(np.array(o))
        o = _d(4)
        f("plane_oplus")(_a(c), _a(np.r_[rng.normal(size=2) * 0.2, rng.normal() * 0.1]), o)
        out["plane_oplus"].append(np.array(o))
        c2 = c + np.r_[rng.normal(size=3) * 0.05, 0.02]
        cv = np.r_[np.cross(c[:3], rng.normal(size=3)), 1.5] + np.r_[rng.normal(size=3) * 0.02, 0]
        for kind, key, other in ((0, "plane_ominus", c2), (1, "plane_ominus_ver", cv), (2, "plane_ominus_par", -c2 if it % 2 else c2)):
            o = _d(3)
            f("plane_ominus")(kind, _a(c), _a(other), o)
            out[key].append(np.array(o))
        o = _d(4)
        f("plane_transform")(_a(pose), _a(c), o)
        out["plane_transform"].append(np.array(o))
        # cuboids
        cu = _cuboid(rng)
        o = _d(10)
        upd = np.r_[rng.normal(size=6) * 0.05, rng.normal(size=3) * 0.02]
        if it % 5 == 0:
            upd[2] = 1e-7  # yaw below the 1e-5 branch of exptwist_norollpitch
        f("cuboid_oplus")(_a(cu), 3, _a(upd), o)
        out["cuboid_oplus"].append(np.array(o))
        o = _d(24)
        f("cuboid_corners")(_a(cu), o)
        out["cuboid_corners"].append(np.array(o))
        cam = np.r_[0.0, 0.0, 0.0, 1.0, -cu[0] + rng.normal() * 0.2, -cu[1] + rng.normal() * 0.2, -cu[2] + 4.0]  # looks at the cuboid from 4 m
        co, bb = _d(16), _d(4)
        f("cuboid_project")(_a(cu), _a(cam), _a(INTR, C.c_float), co, bb)
        out["cuboid_project_corners"].append(np.array(co))
        out["cuboid_project_bbox"].append(np.array(bb))
        pts = cu[:3] + rng.normal(size=(30, 3)) * 0.6
        o = _d(3)
        f("cuboid_point_error")(_a(cu), _a(pts), 30, C.c_double(1.0), C.c_double(0.2), o)
        out["cuboid_point_error"].append(np.array(o))
        o = _d(9)
        f("cuboid_to_minimal")(_a(cu), o)
        out["cuboid_to_minimal"].append(np.array(o))
        o = _d(3)
        f("huber")(C.c_double(rng.uniform(0, 20)), C.c_double(np.float32(np.sqrt(5.991))), o)
        out["huber"].append(np.array(o))
        # point edges: residual + analytic Jacobians
        X = np.array([rng.normal() * 0.5, rng.normal() * 0.5, 3.0 + rng.uniform(0, 3)])
        ident = np.array([0, 0, 0, 1, 0, 0, 0.0])
        campose = _pose(f, rng, 0.05)
        for key, ur in (("point_edge_mono", -1.0), ("point_edge_stereo", 300.0 + rng.normal())):
            err, Jp, Jk = _d(3), _d(9), _d(18)
            D = f("point_edge")(_a(campose), _a(X), _a(INTR, C.c_float), _a([320 + rng.normal() * 50, 250 + rng.normal() * 50, ur], C.c_float), err, Jp, Jk)
            out[key].append(np.r_[D, np.array(err), np.array(Jp), np.array(Jk)])
        # plane edges (all three kinds) and camera-cuboid edges (both kinds)
        for kind, other in ((0, c2), (1, cv), (2, c2)):
            err = _d(3)
            D = f("plane_edge")(kind, _a(c), _a(campose), _a(other), err)
            out["plane_edge"].append(np.r_[D, np.array(err)])
        err = _d(16)
        f("cuboid_cam_edge")(0, _a(cam), _a(cu), _a(INTR, C.c_float), _a(np.array(bb) + rng.normal(size=4) * 2), err)
        out["cuboid_cam_bbox"].append(np.array(err)[:4])
        err = _d(16)
        f("cuboid_cam_edge")(1, _a(cam), _a(cu), _a(INTR, C.c_float), _a(np.array(co) + rng.normal(size=16) * 2), err)
        out["cuboid_cam_corner"].append(np.array(err))
        # EdgeSE3Cuboid: the measured cuboid in the camera frame = the true one seen from `campose3`, with a random choice of the front
        # face (yaw off by k * 90 degrees, x / y half sizes swapped for odd k) and noise; it % 4 == 3: a gross yaw error near the 45-degree tie
        campose3 = _pose(f, rng3, 0.3)
        cu3 = _cuboid(rng3)
        k4 = int(rng3.integers(0, 4))
        yaw_off = k4 * np.pi / 2 + rng3.normal() * 0.05 + (np.pi / 4 - 0.01 if it % 4 == 3 else 0.0)
        qz = np.array([0.0, 0.0, np.sin(yaw_off / 2), np.cos(yaw_off / 2)])
        qa = cu3[3:7]
        qm = np.array([qa[3] * qz[0] + qa[0] * qz[3] + qa[1] * qz[2] - qa[2] * qz[1], qa[3] * qz[1] + qa[1] * qz[3] + qa[2] * qz[0] - qa[0] * qz[2],
                       qa[3] * qz[2] + qa[2] * qz[3] + qa[0] * qz[1] - qa[1] * qz[0], qa[3] * qz[3] - qa[0] * qz[0] - qa[1] * qz[1] - qa[2] * qz[2]])
        sc = cu3[7:10] * (1 + rng3.normal(size=3) * 0.05)
        if k4 % 2:
            sc = sc[[1, 0, 2]]
        world_meas = np.r_[cu3[:3] + rng3.normal(size=3) * 0.05, qm, sc]
        # into the camera frame: Tcw * pose  (rotation by the pose quaternion through se3_map of the centre; orientation q_c * q)
        ctr = _d(3)
        f("se3_map")(_a(campose3), _a(world_meas[:3]), ctr)
        qc = campose3[:4]
        ql = np.array([qc[3] * qm[0] + qc[0] * qm[3] + qc[1] * qm[2] - qc[2] * qm[1], qc[3] * qm[1] + qc[1] * qm[3] + qc[2] * qm[0] - qc[0] * qm[2],
                       qc[3] * qm[2] + qc[2] * qm[3] + qc[0] * qm[1] - qc[1] * qm[0], qc[3] * qm[3] - qc[0] * qm[0] - qc[1] * qm[1] - qc[2] * qm[2]])
        meas3 = np.r_[np.array(ctr), ql / np.linalg.norm(ql), sc, np.zeros(6)]
        err = _d(16)
        D = f("cuboid_cam_edge")(2, _a(campose3), _a(cu3), _a(INTR, C.c_float), _a(meas3), err)
        out["cuboid_cam_se3"].append(np.r_[D, np.array(err)[:9]])
        _ = ident
    return {k: np.array(v) for k, v in out.items()}


def make_graph(name):
    kw, pts_only = WINDOWS[name]
    kw = dict(kw)
    idx = kw.pop("index")
    return ppo.synth.make_graph(ppo.synth.config(idx, **kw)), pts_only


def run_window(name, make_handle, params):
    """optimize(5) -> re-levelling -> optimize(10) of one golden window through any ppo_ba-shaped handle."""
    g, pts_only = make_graph(name)
    if pts_only:
        params.solver = A.SOLVER_6_3
    h = make_handle(params)
    h.set_graph(g)
    res = h.local_ba()
    st = h.get_state()
    out = {"kf_pose": st.kf_pose, "pt_xyz": st.pt_xyz, "pl_coef": st.pl_coef, "cu_state": st.cu_state,
           "iterations": np.array([res.round1.iterations, res.round2.iterations], np.int64),
           "chi2": np.array([res.round1.chi2_initial, res.round2.chi2_final]),
           "outliers": np.array([res.n_outlier_point_edges, res.n_outlier_plane_edges, res.n_outlier_cuboid_edges], np.int64)}
    for kind in range(A.PPO_EDGE_KINDS if hasattr(A, "PPO_EDGE_KINDS") else 5):
        chi2, dpos, norm = h.edge_chi2(kind)
        out["edge_chi2_%d" % kind] = chi2
        out["edge_flags_%d" % kind] = h.get_edge_flags(kind)
    return out
