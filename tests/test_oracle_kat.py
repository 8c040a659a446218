"""Known-answer tests pinning the CPU oracle (oracle/) — the reference ships no tests or golden
vectors for this path, so the restatement is pinned by (1) closed-form
answers written from the cited reference lines, (2) an independent numpy transliteration
(tests/np_ref.py), (3) analytic-vs-numeric Jacobian agreement, (4) a dense numpy re-derivation of
the whole normal-equation / Schur / LM-step algebra on a tiny mixed graph."""
import ctypes as C

import numpy as np
import pytest

import np_ref as R


def _d(n):
    return (C.c_double * n)()


def _arr(x, ct=C.c_double):
    x = np.asarray(x, dtype=np.float64 if ct is C.c_double else np.float32).ravel()
    return (ct * len(x))(*x)


@pytest.fixture(scope="module")
def L(oracle_mod):
    return oracle_mod.lib()


def se3_exp(L, u):
    o = _d(7)
    L.ppo_oracle_se3_exp(_arr(u), o)
    return np.array(o)


def test_exp_identity_and_small_angle_branch(L):
    # se3quat.h:274-308
    assert np.allclose(se3_exp(L, [0] * 6), [0, 0, 0, 1, 0, 0, 0], atol=0)
    for th in (0.999e-5, 1.001e-5, 0.3, 2.5):
        u = np.r_[th * np.array([0.6, -0.48, 0.64]), 0.1, -0.2, 0.3]
        p = se3_exp(L, u)
        Rn, tn = R.se3_exp(u)
        assert np.allclose(R.quat_to_R(p[:4]), Rn, atol=2e-10)  # I+W+W^2 is orthonormal to O(theta^2)=1e-10
        assert np.allclose(p[4:], tn, atol=1e-12)
        assert p[3] >= 0 and abs(np.linalg.norm(p[:4]) - 1) < 1e-15


def test_se3_from_float_rotation_branches(L):
    # Converter.cc:37-47 + Eigen Quaterniond(Matrix3d): trace>0 and the three largest-diagonal branches
    rng = np.random.default_rng(1)
    for axis, ang in [([1, 0, 0], 0.3), ([1, 0, 0], 3.0), ([0, 1, 0], 3.0), ([0, 0, 1], 3.1), ([0.5, 0.5, 0.7071], 2.9)]:
        a = np.array(axis) / np.linalg.norm(axis)
        Rm = R.axis_angle_R(a, ang).astype(np.float32).astype(np.float64)
        t = rng.normal(size=3)
        o = _d(7)
        L.ppo_oracle_se3_from_Rt(_arr(Rm), _arr(t), o)
        p = np.array(o)
        assert p[3] >= 0
        assert np.allclose(R.quat_to_R(p[:4]), Rm, atol=3e-7)  # float32-rounded input
        assert np.allclose(p[4:], t)


def test_se3_oplus_and_map(L):
    rng = np.random.default_rng(2)
    for _ in range(5):
        pose = se3_exp(L, rng.normal(size=6))
        u = rng.normal(size=6) * 0.1
        o = _d(7)
        L.ppo_oracle_se3_oplus(_arr(pose), _arr(u), o)
        Rd, td = R.se3_exp(u)
        R0, t0 = R.pose_to_Rt(pose)
        assert np.allclose(R.quat_to_R(np.array(o)[:4]), Rd @ R0, atol=1e-13)
        assert np.allclose(np.array(o)[4:], Rd @ t0 + td, atol=1e-13)
        x = rng.normal(size=3)
        m = _d(3)
        L.ppo_oracle_se3_map(_arr(pose), _arr(x), m)
        assert np.allclose(np.array(m), R0 @ x + t0, atol=1e-14)


def test_plane_normalize_sign_rule(L):
    # G2O_Plane3D.h:120-125: unit normal, coeffs(3) >= 0
    o = _d(4)
    L.ppo_oracle_plane_normalize(_arr([0, 0, 2, -4]), o)
    assert np.allclose(np.array(o), [0, 0, -1, 2])
    L.ppo_oracle_plane_normalize(_arr([3, 0, 4, 10]), o)
    assert np.allclose(np.array(o), [0.6, 0, 0.8, 2])


def test_plane_ominus_oplus(L):
    rng = np.random.default_rng(3)
    for _ in range(20):
        a = np.r_[rng.normal(size=3), rng.uniform(0.5, 4)]
        o3 = _d(3)
        L.ppo_oracle_plane_ominus(0, _arr(a), _arr(a), o3)
        assert np.allclose(np.array(o3), 0, atol=1e-15)  # ominus(self) = 0
        v = rng.normal(size=3) * 0.05
        o4 = _d(4)
        L.ppo_oracle_plane_oplus(_arr(a), _arr(v), o4)
        assert np.allclose(np.array(o4), R.plane_oplus(a, v), atol=1e-14)
        L.ppo_oracle_plane_ominus(0, _arr(a), o4, o3)  # a.ominus(a (+) v) = (az, el, -dd)
        assert np.allclose(np.array(o3), [v[0], v[1], -v[2]], atol=1e-12)
        b = np.r_[rng.normal(size=3), rng.uniform(0.5, 4)]
        for kind, f in ((0, R.plane_ominus), (1, R.plane_ominus_ver), (2, R.plane_ominus_par)):
            L.ppo_oracle_plane_ominus(kind, _arr(a), _arr(b), o3)
            ref = f(a, b)
            assert np.allclose(np.array(o3)[:len(ref)], ref, atol=1e-13)


def test_plane_ver_par_zero_cases(L):
    o3 = _d(3)
    L.ppo_oracle_plane_ominus(1, _arr([1, 0, 0, 1]), _arr([0, 1, 0, 2]), o3)  # perpendicular
    assert np.allclose(np.array(o3)[:2], 0, atol=1e-15)
    L.ppo_oracle_plane_ominus(2, _arr([0, 0, 1, 1]), _arr([0, 0, 1, 3]), o3)  # parallel
    assert np.allclose(np.array(o3)[:2], 0, atol=1e-15)
    L.ppo_oracle_plane_ominus(2, _arr([0.6, 0, 0.8, 1]), _arr([-0.6, 0, -0.8, 3]), o3)  # anti-parallel
    assert np.allclose(np.array(o3)[:2], 0, atol=1e-15)


def test_plane_transform(L):
    rng = np.random.default_rng(4)
    for _ in range(10):
        pose = se3_exp(L, rng.normal(size=6))
        c = np.r_[rng.normal(size=3), rng.uniform(0.5, 4)]
        o = _d(4)
        L.ppo_oracle_plane_transform(_arr(pose), _arr(c), o)
        assert np.allclose(np.array(o), R.plane_transform(pose, c), atol=1e-14)
        # a world point on the plane stays on the transformed plane
        cn = R.plane_normalize(c)
        X = -cn[3] * cn[:3] + np.cross(cn[:3], rng.normal(size=3))
        Rm, t = R.pose_to_Rt(pose)
        assert abs(np.array(o)[:3] @ (Rm @ X + t) + np.array(o)[3]) < 1e-12


def test_cuboid_corners_order_and_bbox(L):
    # g2o_cuboid.h:198-207: corner sign table; unit cube at identity
    c = [0, 0, 0, 0, 0, 0, 1, 1, 1, 1]
    o = _d(24)
    L.ppo_oracle_cuboid_corners(_arr(c), o)
    assert np.array_equal(np.array(o).reshape(3, 8), R.SGN)
    # camera looking down +z from z=-5: K = [500 0 320; 0 500 240]
    pose = [0, 0, 0, 1, 0, 0, 5]
    intr = _arr([500, 500, 320, 240, 40], C.c_float)
    corners, bbox = _d(16), _d(4)
    L.ppo_oracle_cuboid_project(_arr(c), _arr(pose), intr, corners, bbox)
    # nearest face at depth 4: half-extent 500/4 = 125 px
    assert np.allclose(np.array(bbox), [320, 240, 250, 250])
    assert np.allclose(np.array(corners).reshape(8, 2).T, R.cuboid_project(np.array(c, float), np.array(pose, float), [500, 500, 320, 240]))


def test_cuboid_oplus_yaw_only_fixheight(L):
    rng = np.random.default_rng(5)
    for th in (0.0, 1e-9, 0.2, -1.3):
        yaw0 = rng.uniform(-3, 3)
        c = np.r_[rng.normal(size=3), 0, 0, np.sin(yaw0 / 2), np.cos(yaw0 / 2), rng.uniform(0.2, 0.8, 3)]
        if c[6] < 0:
            c[3:7] *= -1
        u = np.r_[0.3, -0.2, th, rng.normal(size=3) * 0.1, rng.normal(size=3) * 0.01]  # roll/pitch entries ignored
        o = _d(10)
        L.ppo_oracle_cuboid_oplus(_arr(c), 3, _arr(u), o)
        o = np.array(o)
        Rn, tn, sn = R.cuboid_oplus_yaw(c, u)
        assert np.allclose(R.quat_to_R(o[3:7]), Rn, atol=1e-13)
        assert np.allclose(o[:3], tn, atol=1e-13) and o[1] == c[1]  # height (translation y) kept exactly
        assert np.allclose(o[7:], sn)


def test_point_cuboid_error_known_values(L):
    # g2o_cuboid.h:237-255, g2o_cuboid.cc:132-160: scale (1, 2, 0.5)
    c = [0, 0, 0, 0, 0, 0, 1, 1, 2, 0.5]
    pts = np.array([[0.5, 0.5, 0.1],    # inside -> 0
                    [1.5, 0.0, 0.0],    # x outside by 0.5 (< 2*s)
                    [0.0, 5.0, 0.0],    # y beyond 2*s -> capped at s = 2
                    [0.0, 0.0, -0.75]])  # z outside by 0.25
    o = _d(3)
    L.ppo_oracle_cuboid_point_error(_arr(c), _arr(pts), 4, C.c_double(1.0), C.c_double(0.2), o)
    expect = np.array([0.5 / 4 / 1 + 0.2 * 1, 2.0 / 4 / 2 + 0.2 * 2, 0.25 / 4 / 0.5 + 0.2 * 0.5])
    assert np.allclose(np.array(o), expect, atol=1e-15)
    assert np.allclose(np.array(o), R.point_cuboid_error(np.array(c, float), pts))


def test_huber_at_threshold(L):
    # robust_kernel_impl.cpp:76-90
    d = float(np.float32(np.sqrt(5.991)))
    rho = _d(3)
    L.ppo_oracle_huber.argtypes = [C.c_double, C.c_double, C.c_void_p]
    dsqr = float(np.float32(d * d))  # the reference stores delta^2 in a float member (robust_kernel_impl.h:84); found by tests/test_ref_pin.py
    L.ppo_oracle_huber(dsqr, d, rho)
    assert rho[0] == dsqr and rho[1] == 1.0
    L.ppo_oracle_huber(dsqr * (1 + 1e-12), d, rho)
    assert rho[1] < 1.0 and abs(rho[0] - d * d) < 1e-6
    L.ppo_oracle_huber(100.0, 2.0, rho)
    assert np.allclose(list(rho), [2 * 10 * 2 - 4, 0.2, -0.5 * 0.2 / 100])


def test_point_edge_residual_float_quirk_and_jacobians(L):
    rng = np.random.default_rng(6)
    intr = np.array([517.306408, 516.469215, 318.643040, 255.313989, 40.0], np.float32)
    delta = 1e-6
    for stereo in (False, True):
        for _ in range(5):
            pose = se3_exp(L, rng.normal(size=6) * 0.3)
            X = rng.normal(size=3) + [0, 0, 4]
            Rm, t = R.pose_to_Rt(pose)
            pm = _d(3)
            L.ppo_oracle_se3_map(_arr(pose), _arr(X), pm)
            p = np.array(pm)  # the oracle's own camera-frame point (quaternion rotation), so the float quirk is bit-testable
            obs = np.array([300.5, 200.25, 290.0 if stereo else -1.0], np.float32)
            err, Jpt, Jkf = _d(3), _d(9), _d(18)
            D = L.ppo_oracle_point_edge(_arr(pose), _arr(X), _arr(intr, C.c_float), _arr(obs, C.c_float), err, Jpt, Jkf)
            assert D == (3 if stereo else 2)
            e = np.array(err)[:D]
            if stereo:  # types_six_dof_expmap.cpp:182-189: invz and bf*invz in float
                invz = np.float32(1.0 / p[2])
                r0 = p[0] * float(invz) * float(intr[0]) + float(intr[2])
                r1 = p[1] * float(invz) * float(intr[1]) + float(intr[3])
                r2 = r0 - float(np.float32(intr[4]) * invz)
                assert np.array_equal(e, [float(obs[0]) - r0, float(obs[1]) - r1, float(obs[2]) - r2])
                assert np.allclose(e, R.project(p, intr, obs), atol=5e-4)  # float 1/z costs ~1e-5 px
            else:
                assert np.allclose(e, R.project(p, intr, obs), atol=1e-11)
            # analytic Jacobians vs central differences of the exact (double) projection
            Jn_pt = np.zeros((D, 3))
            Jn_kf = np.zeros((D, 6))
            for d in range(3):
                dx = np.zeros(3); dx[d] = delta
                Jn_pt[:, d] = (R.project(Rm @ (X + dx) + t, intr, obs) - R.project(Rm @ (X - dx) + t, intr, obs)) / (2 * delta)
            for d in range(6):
                du = np.zeros(6); du[d] = delta
                Rp, tp = R.se3_exp(du); Rq, tq = R.se3_exp(-du)
                Jn_kf[:, d] = (R.project(Rp @ p + tp, intr, obs) - R.project(Rq @ p + tq, intr, obs)) / (2 * delta)
            assert np.allclose(np.array(Jpt)[:3 * D].reshape(D, 3), Jn_pt, rtol=1e-6, atol=1e-4)
            assert np.allclose(np.array(Jkf)[:6 * D].reshape(D, 6), Jn_kf, rtol=1e-6, atol=1e-4)


def test_dense_ldlt(L):
    rng = np.random.default_rng(7)
    n = 57
    M = rng.normal(size=(n, n))
    Aspd = M @ M.T + n * np.eye(n)
    rhs = rng.normal(size=n)
    sol = _d(n)
    assert L.ppo_oracle_dense_solve(n, _arr(np.triu(Aspd)), _arr(rhs), sol) == 1
    assert np.allclose(np.array(sol), np.linalg.solve(Aspd, rhs), rtol=1e-10)
    Aind = Aspd - 3 * n * np.eye(n)  # not positive definite -> isPositive() false (linear_solver_dense.h:108-112)
    assert L.ppo_oracle_dense_solve(n, _arr(np.triu(Aind)), _arr(rhs), sol) == 0


def test_linear_solver_eigen_flavour_solves_indefinite_systems(L):
    """LinearSolverEigen (solvers/linear_solver_eigen.h:94-124) = Eigen::SimplicialLDLT: LDL^T without pivoting that fails only on
    an exactly zero pivot; LinearSolverDense rejects the same system (isPositive() false).  Flavours: 0 = PPO_SOLVER_DENSE_X, 1 = PPO_SOLVER_6_3."""
    rng = np.random.default_rng(11)
    n = 83
    M = rng.normal(size=(n, n))
    Aspd = M @ M.T + n * np.eye(n)
    Aind = Aspd - 1.7 * n * np.eye(n)
    assert (np.linalg.eigvalsh(Aind) < 0).any() and (np.linalg.eigvalsh(Aind) > 0).any()
    rhs = rng.normal(size=n)
    sol = _d(n)
    for A_, want in ((Aspd, 1), (Aind, 1)):
        assert L.ppo_oracle_dense_solve_flavour(1, n, _arr(np.triu(A_)), _arr(rhs), sol) == want
        x = np.array(sol)
        assert np.abs(A_ @ x - rhs).max() <= 1e-9 * (np.abs(A_).max() * np.abs(x).max())
    assert L.ppo_oracle_dense_solve_flavour(0, n, _arr(np.triu(Aind)), _arr(rhs), sol) == 0
    Azero = Aspd.copy()
    Azero[0, :] = Azero[:, 0] = 0.0  # first pivot exactly zero: "failure, D(k,k) is zero"
    assert L.ppo_oracle_dense_solve_flavour(1, n, _arr(np.triu(Azero)), _arr(rhs), sol) == 0


# ----- whole-system algebra re-derived densely in numpy -------------------------------------------
def _np_system(g, P, A, flags=None, rec=None):
    """Builds the full (un-reduced) Gauss-Newton system of a tiny graph with np_ref residuals and
    central differences (step 1e-6), ordering [free KFs | cuboids | planes | points].
    flags: optional {edge kind: uint8 array of EF_LEVEL1 | EF_ROBUST} (default: every edge active and robust);
    rec: optional dict that receives per-edge (chi2, ||error||, depth positive) lists per kind."""
    a = g.a
    n_kf, n_pt, n_pl, n_cu = g.c.n_kf, g.c.n_pt, g.c.n_pl, g.c.n_cu
    free_kf = [i for i in range(n_kf) if not a["kf_fixed"][i]]
    off = {}
    o = 0
    for i in free_kf:
        off[("kf", i)] = o; o += 6
    for i in range(n_cu):
        off[("cu", i)] = o; o += 9
    n_p = o
    for i in range(n_pl):
        off[("pl", i)] = o; o += 3
    for i in range(n_pt):
        off[("pt", i)] = o; o += 3
    N = o
    H = np.zeros((N, N)); b = np.zeros(N)
    chi_tot = 0.0

    def kf_oplus(p, u):
        Rd, td = R.se3_exp(u); R0, t0 = R.pose_to_Rt(p)
        return Rd @ R0, Rd @ t0 + td

    def add(res_fn, verts, W, delta_h, kind=None, idx=None, extra=None):
        # verts: list of (key, dim, oplus_fn(u) -> replacement value), res_fn(values dict) -> residual
        nonlocal chi_tot
        fl = A.EF_ROBUST if flags is None or kind is None else int(flags[kind][idx])
        if fl & A.EF_LEVEL1:
            return
        r0 = res_fn({})
        chi = float(r0 @ (W * r0))
        if rec is not None and kind is not None:
            rec.setdefault(kind, {})[idx] = (chi, float(np.linalg.norm(r0)), extra() if extra else True)
        w = 1.0
        rho0 = chi
        if delta_h is not None and (fl & A.EF_ROBUST):
            rho0, w = R.huber(chi, delta_h)
        chi_tot += rho0
        Js = []
        for key, dim, _ in verts:
            J = np.zeros((len(r0), dim))
            for d in range(dim):
                u = np.zeros(dim); u[d] = 1e-6
                J[:, d] = (res_fn({key: u}) - res_fn({key: -u})) / 2e-6
            Js.append(J)
        for (k1, d1, _), J1 in zip(verts, Js):
            if k1 not in off:
                continue
            o1 = off[k1]
            b[o1:o1 + d1] += -w * J1.T @ (W * r0)
            for (k2, d2, _), J2 in zip(verts, Js):
                if k2 not in off:
                    continue
                o2 = off[k2]
                H[o1:o1 + d1, o2:o2 + d2] += w * J1.T @ (W[:, None] * J2)

    rp = a["pt_rowptr"]
    for p in range(n_pt):
        for e in range(rp[p], rp[p + 1]):
            k = int(a["pe_kf"][e]); obs = a["pe_obs"][e]; intr = a["kf_intr"][k]
            D = 2 if obs[2] < 0 else 3

            def res(pert, k=k, p=p, obs=obs, intr=intr):
                Rk, tk = kf_oplus(a["kf_pose"][k], pert.get(("kf", k), np.zeros(6)))
                X = a["pt_xyz"][p] + pert.get(("pt", p), np.zeros(3))
                return R.project(Rk @ X + tk, intr, obs)
            W = np.full(D, float(a["pe_invsigma2"][e]))

            def depth_pos(k=k, p=p):  # isDepthPositive(): z of the point in the camera frame
                Rk, tk = R.pose_to_Rt(a["kf_pose"][k])
                return bool((Rk @ a["pt_xyz"][p] + tk)[2] > 0)
            add(res, [(("kf", k), 6, None), (("pt", p), 3, None)], W, P.huber_mono if D == 2 else P.huber_stereo, A.EDGE_POINT, e, depth_pos)
    for e in range(g.c.n_ple):
        k = int(a["ple_kf"][e]); pl = int(a["ple_plane"][e]); kind = int(a["ple_kind"][e]); meas = a["ple_meas"][e]
        D = 3 if kind == 0 else 2

        def res(pert, k=k, pl=pl, kind=kind, meas=meas):
            Rk, tk = kf_oplus(a["kf_pose"][k], pert.get(("kf", k), np.zeros(6)))
            c = R.plane_oplus(a["pl_coef"][pl], pert[("pl", pl)]) if ("pl", pl) in pert else R.plane_normalize(a["pl_coef"][pl])
            n2 = Rk @ c[:3]; d2 = c[3] - tk @ n2
            loc = R.plane_normalize(np.r_[n2, d2] if d2 >= 0 else -np.r_[n2, d2])
            return (R.plane_ominus, R.plane_ominus_ver, R.plane_ominus_par)[kind](loc, meas)
        add(res, [(("pl", pl), 3, None), (("kf", k), 6, None)], a["ple_info"][e][:D], P.huber_plane if kind == 0 else P.huber_vp_plane, A.EDGE_PLANE, e)
    for e in range(g.c.n_cbe):
        k = int(a["cbe_kf"][e]); cu = int(a["cbe_cuboid"][e]); kind = int(a["cbe_kind"][e])
        D = 4 if kind == 0 else 16
        meas = a["cbe_meas"][e][:D]; intr = a["kf_intr"][k]

        def res(pert, k=k, cu=cu, kind=kind, meas=meas, intr=intr):
            Rk, tk = kf_oplus(a["kf_pose"][k], pert.get(("kf", k), np.zeros(6)))
            Rc, tc, sc = R.cuboid_oplus_yaw(a["cu_state"][cu], pert.get(("cu", cu), np.zeros(9)))
            pc = Rk @ (Rc @ (sc[:, None] * R.SGN) + tc[:, None]) + tk[:, None]
            uv = np.stack([float(intr[0]) * pc[0] / pc[2] + float(intr[2]), float(intr[1]) * pc[1] / pc[2] + float(intr[3])])
            if kind == 0:
                mn, mx = uv.min(axis=1), uv.max(axis=1)
                return np.r_[(mn + mx) / 2, mx - mn] - meas
            return uv.T.ravel() - meas
        add(res, [(("kf", k), 6, None), (("cu", cu), 9, None)], np.full(D, a["cbe_info"][e]), P.huber_bbox if kind == 0 else P.huber_corner, A.EDGE_CUBOID_CAM, e)
    for e in range(g.c.n_pce):
        cu = int(a["pce_cuboid"][e]); pts = a["pce_pts"][a["pce_rowptr"][e]:a["pce_rowptr"][e + 1]]

        def res(pert, cu=cu, pts=pts):
            Rc, tc, sc = R.cuboid_oplus_yaw(a["cu_state"][cu], pert.get(("cu", cu), np.zeros(9)))
            lp = np.abs((pts - tc) @ Rc)
            er = np.where(lp < sc, 0.0, np.where(lp < 2 * sc, lp - sc, sc))
            return er.mean(axis=0) / sc + 0.2 * sc
        add(res, [(("cu", cu), 9, None)], np.ones(3), None, A.EDGE_POINT_CUBOID, e)
    for e in range(g.c.n_cpe):
        fl = A.EF_ROBUST if flags is None else int(flags[A.EDGE_CUBOID_PLANE][e])
        if fl & A.EF_LEVEL1:
            continue
        r0 = a["cpe_meas"][e]
        chi = float(r0 @ (a["cpe_info"][e] * r0))
        if rec is not None:
            rec.setdefault(A.EDGE_CUBOID_PLANE, {})[e] = (chi, float(np.linalg.norm(r0)), True)
        chi_tot += R.huber(chi, P.huber_cuboid_plane)[0] if (fl & A.EF_ROBUST) else chi
    return H, b, n_p, N, chi_tot


def test_system_schur_and_lm_step_against_dense_numpy(ppo, oracle_mod):
    A = ppo.abi
    cfg = ppo.synth.config(1, n_kf=5, n_fixed=2, n_pt=40, n_pl=3, n_cu=2, corners_2d=1)
    g = ppo.synth.make_graph(cfg)
    assert g.c.n_ple > 0 and g.c.n_cbe > 0 and g.c.n_pce > 0 and g.c.n_cpe > 0
    o = oracle_mod.Oracle()
    o.set_graph(g)
    lin = o.debug_linearize()
    H, b, n_p, N, chi = _np_system(g, o.params, A)
    assert lin["n_p"] == n_p and lin["n_p"] + 3 * lin["n_l"] == N
    assert np.isclose(lin["chi2"], chi, rtol=1e-7)  # stereo float quirk only
    scale = np.abs(H).max()
    # pose block (upper part as g2o stores it)
    assert np.allclose(np.triu(lin["Hpp"]), np.triu(H[:n_p, :n_p]), rtol=2e-5, atol=2e-6 * scale)
    Hll = np.stack([H[n_p + 3 * l:n_p + 3 * l + 3, n_p + 3 * l:n_p + 3 * l + 3].ravel() for l in range(lin["n_l"])])
    assert np.allclose(lin["Hll"], Hll, rtol=2e-5, atol=2e-6 * scale)
    assert np.allclose(lin["b"], b, rtol=2e-5, atol=2e-6 * np.abs(b).max())
    # damped step: the Schur route must equal the dense solve of the full system
    lam = 1e-5 * np.abs(np.diag(H)).max()
    sol = o.debug_solve(lam, lin["n_p"], lin["n_l"])
    assert sol["ok"] == 1
    x_ref = np.linalg.solve(H + lam * np.eye(N), b)
    assert np.allclose(sol["x"], x_ref, rtol=1e-4, atol=1e-6 * np.abs(x_ref).max())
    # Schur complement itself
    Hpl = H[:n_p, n_p:]
    Hll_full = H[n_p:, n_p:] + lam * np.eye(N - n_p)
    S_ref = H[:n_p, :n_p] + lam * np.eye(n_p) - Hpl @ np.linalg.solve(Hll_full, Hpl.T)
    assert np.allclose(np.triu(sol["Hschur"]), np.triu(S_ref), rtol=1e-4, atol=1e-6 * np.abs(S_ref).max())


def test_lm_schedule_converges_and_matches_reference_rules(ppo, oracle_mod):
    g, truth = ppo.synth.make_graph(ppo.synth.config(0), with_truth=True)
    o = oracle_mod.Oracle()
    o.set_graph(g)
    res = o.local_ba()
    assert res.round1.iterations == 5 and res.round1.n_pose_dim == 54  # 9 free KFs x 6
    tr = res.round1.trace_list()
    assert all(t["accepted"] for t in tr) and tr[-1]["chi2_after"] < 0.2 * tr[0]["chi2_before"]
    # lambda follows levenberg.cpp:134-141 on accepted steps: x max(1/3, min(1-(2rho-1)^3, 2/3))
    for a_, b_ in zip(tr[:-1], tr[1:]):
        f = max(1 / 3, min(1 - (2 * b_["rho"] - 1) ** 3, 2 / 3))
        assert np.isclose(b_["lam"], a_["lam"] * f, rtol=1e-12)
    s = o.get_state()
    assert np.abs(s.kf_pose - truth.kf_pose).max() < 5e-3
    assert 0.02 < res.n_outlier_point_edges / g.c.n_pe < 0.12  # 3 % gross + chi2 tail
    # stop flag set on entry: nothing happens (Optimizer.cc:2723-2725)
    o.reset()
    stop = np.ones(1, np.uint8)
    res2 = o.local_ba(stop)
    assert res2.skipped == 1 and np.allclose(o.get_state().kf_pose, g["kf_pose"], rtol=0, atol=1e-15)


# ----- the whole LM loop re-implemented in numpy on the dense, un-reduced system -----------------------------------
def _R_to_quat(Rm):
    """rotation matrix -> [x y z w], w >= 0 (branch on the trace; any valid branch gives the same rotation)."""
    tr = np.trace(Rm)
    if tr > 0:
        s = 2 * np.sqrt(tr + 1)
        q = np.array([(Rm[2, 1] - Rm[1, 2]) / s, (Rm[0, 2] - Rm[2, 0]) / s, (Rm[1, 0] - Rm[0, 1]) / s, s / 4])
    else:
        i = int(np.argmax(np.diag(Rm)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = 2 * np.sqrt(1 + Rm[i, i] - Rm[j, j] - Rm[k, k])
        q = np.zeros(4)
        q[i] = s / 4
        q[j] = (Rm[j, i] + Rm[i, j]) / s
        q[k] = (Rm[k, i] + Rm[i, k]) / s
        q[3] = (Rm[k, j] - Rm[j, k]) / s
    q /= np.linalg.norm(q)
    return -q if q[3] < 0 else q


def _np_apply(A, g, x, n_p):
    """x (ordering of _np_system) applied with the vertices' oplus; returns a new GraphArrays."""
    a = {k: v.copy() for k, v in g.a.items()}
    o = 0
    for i in range(g.c.n_kf):
        if a["kf_fixed"][i]:
            continue
        Rd, td = R.se3_exp(x[o:o + 6]); R0, t0 = R.pose_to_Rt(a["kf_pose"][i])
        a["kf_pose"][i] = np.r_[_R_to_quat(Rd @ R0), Rd @ t0 + td]
        o += 6
    for i in range(g.c.n_cu):
        Rc, tc, sc = R.cuboid_oplus_yaw(a["cu_state"][i], x[o:o + 9])
        a["cu_state"][i] = np.r_[tc, _R_to_quat(Rc), sc]
        o += 9
    assert o == n_p
    for i in range(g.c.n_pl):
        a["pl_coef"][i] = R.plane_oplus(a["pl_coef"][i], x[o:o + 3])
        o += 3
    for i in range(g.c.n_pt):
        a["pt_xyz"][i] = a["pt_xyz"][i] + x[o:o + 3]
        o += 3
    return A.GraphArrays(**a)


def test_lm_iterations_against_numpy_reimplementation(ppo, oracle_mod):
    """optimize(3) of the oracle against a numpy Levenberg-Marquardt written from levenberg.cpp:61-180 on the dense,
    un-reduced Gauss-Newton system of np_ref (no Schur complement, no block structure, rotation matrices, its own
    central differences): lambda_0 = tau max diag, rho = (chi - chi') / (x.(lambda x + b) + 1e-3), accept factor
    max(1/3, min(1 - (2 rho - 1)^3, 2/3)), reject factor ni (doubling)."""
    A = ppo.abi
    g = ppo.synth.make_graph(ppo.synth.config(1, n_kf=5, n_fixed=2, n_pt=40, n_pl=3, n_cu=2, corners_2d=1))
    o = oracle_mod.Oracle()
    o.set_graph(g)
    stats = o.optimize(3)
    tr = stats.trace_list()
    P = o.params
    cur = g
    lam, ni = None, 2.0
    for it in range(3):
        H, b, n_p, N, chi = _np_system(cur, P, A)
        if it == 0:
            lam = P.lm_tau * np.abs(np.diag(H)).max()
        assert np.isclose(chi, tr[it]["chi2_before"], rtol=1e-6)
        trials = 0
        while True:
            x = np.linalg.solve(H + lam * np.eye(N), b)
            trial = _np_apply(A, cur, x, n_p)
            chi_new = _np_system(trial, P, A)[4]
            rho = (chi - chi_new) / (float(x @ (lam * x + b)) + 1e-3)
            trials += 1
            if rho > 0 and np.isfinite(chi_new):
                lam *= max(1 / 3, min(1 - (2 * rho - 1) ** 3, 2 / 3))
                ni = 2.0
                cur, chi = trial, chi_new
                break
            lam *= ni
            ni *= 2
            assert trials < 10
        assert trials == tr[it]["trials"] and tr[it]["accepted"] == 1
        assert np.isclose(chi, tr[it]["chi2_after"], rtol=1e-5)
        assert np.isclose(lam, tr[it]["lam"], rtol=1e-3)
    s = o.get_state()
    assert np.abs(s.kf_pose - cur["kf_pose"]).max() < 1e-6
    assert np.abs(s.pt_xyz - cur["pt_xyz"]).max() < 1e-5
    assert np.abs(s.pl_coef - cur["pl_coef"]).max() < 1e-6
    assert np.abs(s.cu_state - cur["cu_state"]).max() < 1e-5


def _np_lm(A, g, P, iters, flags, trace):
    """levenberg.cpp:61-180 on the dense system; checks every iteration against the oracle's trace; returns the final graph."""
    cur, lam, ni = g, None, 2.0
    for it in range(iters):
        H, b, n_p, N, chi = _np_system(cur, P, A, flags)
        if it == 0:
            lam = P.lm_tau * np.abs(np.diag(H)).max()
        assert np.isclose(chi, trace[it]["chi2_before"], rtol=1e-6), it
        trials = 0
        while True:
            x = np.linalg.solve(H + lam * np.eye(N), b)
            trial = _np_apply(A, cur, x, n_p)
            chi_new = _np_system(trial, P, A, flags)[4]
            rho = (chi - chi_new) / (float(x @ (lam * x + b)) + 1e-3)
            trials += 1
            if rho > 0 and np.isfinite(chi_new):
                lam *= max(1 / 3, min(1 - (2 * rho - 1) ** 3, 2 / 3))
                ni = 2.0
                cur, chi = trial, chi_new
                break
            lam *= ni
            ni *= 2
            assert trials < 10
        assert trials == trace[it]["trials"] and trace[it]["accepted"] == 1, it
        assert np.isclose(chi, trace[it]["chi2_after"], rtol=2e-5), it
        assert np.isclose(lam, trace[it]["lam"], rtol=5e-3), it
    return cur


@pytest.mark.parametrize("inject", [False, True])
def test_full_schedule_against_numpy_reimplementation(ppo, oracle_mod, inject):
    """The complete LocalBACameraPlaneCuboids schedule, optimize(5) -> outlier pass -> optimize(10), in numpy: the
    re-levelling rules are written from Optimizer.cc:2736-2833 (points: chi2 > 5.991 / 7.815 or negative depth -> level 1,
    robust kernel off for all; bbox / corner edges: ||error|| > 80 / 10 -> level 1, kernel kept; plane / ver / par edges:
    chi2 > 500 / 200 -> level 1, kernel off; cuboid-plane: ||error|| > 500 -> level 1; point-cuboid: untouched)."""
    A = ppo.abi
    g = ppo.synth.make_graph(ppo.synth.config(1, n_kf=5, n_fixed=2, n_pt=40, n_pl=3, n_cu=2, corners_2d=1))
    if inject:  # one gross plane observation and one gross cuboid-corner observation, so that those rules fire too
        a = {k: v.copy() for k, v in g.a.items()}
        e = int(np.argmax(a["ple_kind"] == 0))
        a["ple_meas"][e] = R.plane_normalize(a["ple_meas"][e] + np.array([0.9, -0.7, 0.4, 1.5]))
        a["cbe_meas"][int(np.argmax(a["cbe_kind"] == 1))][:16] += 60.0
        g = A.GraphArrays(**a)
    o = oracle_mod.Oracle()
    o.set_graph(g)
    res = o.local_ba()
    if inject:
        assert res.n_outlier_plane_edges >= 1 and res.n_outlier_cuboid_edges >= 1
    P = o.params
    n_edges = {A.EDGE_POINT: g.c.n_pe, A.EDGE_PLANE: g.c.n_ple, A.EDGE_CUBOID_CAM: g.c.n_cbe, A.EDGE_POINT_CUBOID: g.c.n_pce,
               A.EDGE_CUBOID_PLANE: g.c.n_cpe}
    flags = {k: np.full(n, A.EF_ROBUST, np.uint8) for k, n in n_edges.items()}
    flags[A.EDGE_POINT_CUBOID][:] = 0  # EdgePointCuboidOnlyObject never gets a robust kernel (Optimizer.cc:2630-2650)
    r1 = res.round1.trace_list()
    assert all(t["accepted"] for t in r1)  # otherwise the per-edge errors would be those of a rejected trial
    cur = _np_lm(A, g, P, res.round1.iterations, flags, r1)
    rec = {}
    _np_system(cur, P, A, flags, rec)
    for e, (chi, _, dpos) in rec[A.EDGE_POINT].items():
        mono = cur["pe_obs"][e][2] < 0
        flags[A.EDGE_POINT][e] = A.EF_LEVEL1 if (chi > (5.991 if mono else 7.815) or not dpos) else 0
    for e, (_, nrm, _) in rec[A.EDGE_CUBOID_CAM].items():
        th = 80.0 if cur["cbe_kind"][e] == 0 else 10.0
        flags[A.EDGE_CUBOID_CAM][e] = A.EF_ROBUST | (A.EF_LEVEL1 if nrm > th else 0)
    for e, (chi, _, _) in rec[A.EDGE_PLANE].items():
        th = 500.0 if cur["ple_kind"][e] == 0 else 200.0
        flags[A.EDGE_PLANE][e] = A.EF_LEVEL1 if chi > th else 0
    for e, (_, nrm, _) in rec.get(A.EDGE_CUBOID_PLANE, {}).items():
        flags[A.EDGE_CUBOID_PLANE][e] = A.EF_ROBUST | (A.EF_LEVEL1 if nrm > 500.0 else 0)
    for kind in (A.EDGE_POINT, A.EDGE_PLANE, A.EDGE_CUBOID_CAM):
        assert np.array_equal(flags[kind], o.get_edge_flags(kind)), kind
    assert int((flags[A.EDGE_POINT] & A.EF_LEVEL1).sum()) == res.n_outlier_point_edges
    r2 = res.round2.trace_list()
    assert all(t["accepted"] for t in r2)
    cur = _np_lm(A, cur, P, res.round2.iterations, flags, r2)
    s = o.get_state()
    assert np.abs(s.kf_pose - cur["kf_pose"]).max() < 5e-6
    assert np.abs(s.pt_xyz - cur["pt_xyz"]).max() < 5e-5
    assert np.abs(s.pl_coef - cur["pl_coef"]).max() < 5e-6
    assert np.abs(s.cu_state - cur["cu_state"]).max() < 5e-5


def test_levelled_out_edges_keep_their_round1_error(ppo, oracle_mod):
    """level-1 edges are inactive in round 2, computeActiveErrors never refreshes them, so the erase test
    of Optimizer.cc:2856-2881 reads their ROUND-1 chi2; active edges are refreshed.  isDepthPositive uses the final
    estimates.  Points whose every edge was levelled out keep their round-1 position (dropped from the index mapping)."""
    A = ppo.abi
    g = ppo.synth.make_graph(ppo.synth.config(0, n_kf=8, n_fixed=2, n_pt=400))
    o = oracle_mod.Oracle()
    o.params.solver = A.SOLVER_6_3
    o.set_graph(g)
    o.optimize(5)
    chi_r1, _, _ = o.edge_chi2(A.EDGE_POINT)
    pts_r1 = o.get_state().pt_xyz.copy()
    n_out = o.outlier_pass()
    flags = o.get_edge_flags(A.EDGE_POINT)
    lvl1 = (flags & A.EF_LEVEL1) != 0
    assert lvl1.sum() == n_out[0] > 0 and not (flags & A.EF_ROBUST).any()  # kernels dropped on every point edge (:2752,2768)
    mono = g["pe_obs"][:, 2] < 0
    assert np.array_equal(lvl1, (chi_r1 > np.where(mono, 5.991, 7.815)) | (o.edge_chi2(A.EDGE_POINT)[1] == 0))
    o.optimize(10)
    chi_r2, _, _ = o.edge_chi2(A.EDGE_POINT)
    assert np.array_equal(chi_r2[lvl1], chi_r1[lvl1])                 # stale: exactly the round-1 values
    assert (chi_r2[~lvl1] != chi_r1[~lvl1]).mean() > 0.99             # refreshed
    # a point that lost all its edges is no longer a vertex of round 2: it keeps its round-1 estimate
    rp = g["pt_rowptr"]
    dropped = [p for p in range(g.c.n_pt) if lvl1[rp[p]:rp[p + 1]].all()]
    pts_r2 = o.get_state().pt_xyz
    for p in dropped:
        assert np.array_equal(pts_r2[p], pts_r1[p])
    moved = np.abs(pts_r2 - pts_r1).max(axis=1) > 0
    assert moved.sum() >= g.c.n_pt - len(dropped) - 5


def test_threaded_oracle_variant_agrees_with_the_single_threaded_one(ppo, oracle_mod):
    """The OpenMP variant (reported next to the reference-configuration baseline) must give the single-threaded
    oracle's answer up to summation order: same schedule and outlier sets, estimates within 1e-6."""
    g = ppo.synth.make_graph(ppo.synth.config(1, n_kf=16, n_pt=3000, n_pl=8, n_cu=4))
    runs = []
    for threads in (1, 4):
        o = oracle_mod.Oracle()
        o.set_threads(threads)
        o.set_graph(g)
        r = o.local_ba()
        runs.append((r, o.get_state(), [o.get_edge_flags(k) for k in range(ppo.abi.EDGE_KINDS)]))
    (ra, sa, fa), (rb, sb, fb) = runs
    assert (ra.round1.iterations, ra.round2.iterations) == (rb.round1.iterations, rb.round2.iterations)
    assert all(np.array_equal(x, y) for x, y in zip(fa, fb))
    assert np.isclose(ra.round2.chi2_final, rb.round2.chi2_final, rtol=1e-7)
    assert np.abs(sa.kf_pose - sb.kf_pose).max() < 1e-7 and np.abs(sa.pt_xyz - sb.pt_xyz).max() < 1e-4
    assert np.abs(sa.pl_coef - sb.pl_coef).max() < 1e-6 and np.abs(sa.cu_state - sb.cu_state).max() < 1e-4  # cuboids: numeric Jacobians
