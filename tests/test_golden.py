"""Committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py): seeded inputs plus the oracle's
outputs.  CPU tests pin the oracle to them; the GPU test checks the CUDA engine (through the C-ABI) against the same
files, so that parity does not depend on the oracle binary built on the day of the test."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = ["points_only", "mixed_bbox", "mixed_corners"]
TOL = 1e-4  # north_star tolerance for the CUDA engine


def load(ppo, name):
    z = np.load(os.path.join(HERE, "golden", name + ".npz"))
    g = ppo.abi.GraphArrays(**{k[3:]: z[k] for k in z.files if k.startswith("in_")})
    out = {k[4:]: z[k] for k in z.files if k.startswith("out_")}
    return g, out


def run(handle, g):
    handle.set_graph(g)
    lin = handle.debug_linearize()
    res = handle.local_ba()
    return lin, res, handle.get_state()


def check(ppo, h, g, out, state_tol, chi_tol, edge_tol):
    A = ppo.abi
    lin, res, st = run(h, g)
    assert [lin["n_p"], lin["n_l"]] == out["lin_dims"].tolist()
    assert np.isclose(lin["chi2"], float(out["lin_chi2"]), rtol=chi_tol)
    scale = lambda a: max(np.abs(a).max(), 1e-300)
    assert np.abs(np.triu(lin["Hpp"]) - np.triu(out["lin_Hpp"])).max() <= 1e-6 * scale(out["lin_Hpp"])
    assert np.abs(lin["Hll"] - out["lin_Hll"]).max() <= 1e-6 * scale(out["lin_Hll"])
    assert np.abs(lin["b"] - out["lin_b"]).max() <= 1e-6 * scale(out["lin_b"])
    for rname, r in (("r1", res.round1), ("r2", res.round2)):
        want = out[rname + "_summary"].tolist()
        assert [r.iterations, r.terminated, r.n_pose_dim, r.n_landmarks, r.n_active_edges] == want
        tr = r.trace_list()
        gt = out[rname + "_trace"]
        assert len(tr) == len(gt)
        for t, w in zip(tr, gt):
            assert (t["trials"], t["accepted"]) == (int(w[3]), int(w[4]))
            assert np.isclose(t["chi2_before"], w[0], rtol=chi_tol) and np.isclose(t["chi2_after"], w[1], rtol=chi_tol)
            assert np.isclose(t["lam"], w[2], rtol=2e-2)
        assert np.isclose(r.chi2_final, float(out[rname + "_chi2_final"]), rtol=chi_tol)
    assert [res.n_outlier_point_edges, res.n_outlier_plane_edges, res.n_outlier_cuboid_edges] == out["outliers"].tolist()
    for name in ("kf_pose", "pt_xyz", "pl_coef", "cu_state"):
        a, b = getattr(st, name), out[name]
        assert a.shape == b.shape
        if a.size:
            assert np.abs(a - b).max() <= state_tol * max(1.0, np.abs(b).max()), name
    for kind in range(A.EDGE_KINDS):
        chi2, depth, _ = h.edge_chi2(kind)
        want = out[f"edge{kind}_chi2"]
        assert chi2.shape == want.shape
        if chi2.size:
            rt = max(edge_tol, 2e-2) if (kind == A.EDGE_CUBOID_CAM and edge_tol > 1e-6) else edge_tol
            assert np.allclose(chi2, want, rtol=rt, atol=1e-5 * max(1.0, np.abs(want).max())), kind
            assert np.array_equal(depth, out[f"edge{kind}_depth"])
            assert np.array_equal(h.get_edge_flags(kind), out[f"edge{kind}_flags"])


@pytest.mark.parametrize("name", CASES)
def test_golden_inputs_are_the_seeded_windows(ppo, name):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    g, _ = mg.make_graph(name)
    want, _ = load(ppo, name)
    assert set(g.a) == set(want.a)
    for k in g.a:
        assert np.array_equal(g.a[k], want.a[k]), k


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_golden(ppo, oracle_mod, name):
    g, out = load(ppo, name)
    p = oracle_mod.default_params()
    if name == "points_only":
        p.solver = ppo.abi.SOLVER_6_3
    check(ppo, oracle_mod.Oracle(p), g, out, state_tol=1e-9, chi_tol=1e-9, edge_tol=1e-7)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_engine_matches_golden(ppo, name):
    g, out = load(ppo, name)
    p = ppo.default_params()
    if name == "points_only":
        p.solver = ppo.abi.SOLVER_6_3
    check(ppo, ppo.LocalBA(p), g, out, state_tol=TOL, chi_tol=1e-6, edge_tol=1e-3)
