"""Generates the golden input/output fixtures in this directory.

What they are: small seeded synthetic windows (inputs as flat ppo_ba_graph arrays) together with the outputs of the
CPU oracle (oracle/ppo_oracle.cpp) on them: final estimates, per-edge chi2, outlier flags, the LM trace of both
optimisation rounds and the linearised blocks at the initial estimate.  What they are NOT: outputs of the reference
binary.  The reference (g2o + Eigen + OpenCV + PCL) cannot be compiled in this image and ships no tests or golden
vectors of its own, so parity stays "unpinned": these files pin the ORACLE (any later change to
it that moves a number shows up in tests/test_golden.py) and give the GPU tests a fixed, reference-independent target.

Run from the repo root:   python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from ppo_pkg import ppo  # noqa: E402
import oracle_lib  # noqa: E402

A = ppo.abi

# name -> (synth config overrides, points-only solver?)
CASES = {
    "points_only": (dict(index=0, n_kf=6, n_fixed=2, n_pt=160), True),
    "mixed_bbox": (dict(index=1, n_kf=8, n_fixed=2, n_pt=220, n_pl=4, n_cu=3), False),
    "mixed_corners": (dict(index=1, n_kf=8, n_fixed=2, n_pt=220, n_pl=4, n_cu=3, corners_2d=1, cuboid_2d=0), False),
}


def make_graph(name):
    kw, pts_only = CASES[name]
    kw = dict(kw)
    idx = kw.pop("index")
    return ppo.synth.make_graph(ppo.synth.config(idx, **kw)), pts_only


def run_oracle(g, pts_only):
    p = oracle_lib.default_params()
    if pts_only:
        p.solver = A.SOLVER_6_3
    o = oracle_lib.Oracle(p)
    o.set_graph(g)
    lin = o.debug_linearize()
    res = o.local_ba()
    st = o.get_state()
    out = {
        "lin_chi2": np.float64(lin["chi2"]), "lin_Hpp": lin["Hpp"], "lin_Hll": lin["Hll"], "lin_b": lin["b"],
        "lin_dims": np.array([lin["n_p"], lin["n_l"]], np.int64),
        "kf_pose": st.kf_pose, "pt_xyz": st.pt_xyz, "pl_coef": st.pl_coef, "cu_state": st.cu_state,
        "outliers": np.array([res.n_outlier_point_edges, res.n_outlier_plane_edges, res.n_outlier_cuboid_edges], np.int64),
    }
    for rname, r in (("r1", res.round1), ("r2", res.round2)):
        tr = r.trace_list()
        out[rname + "_trace"] = np.array([[t["chi2_before"], t["chi2_after"], t["lam"], t["trials"], t["accepted"]] for t in tr], np.float64).reshape(-1, 5)
        out[rname + "_summary"] = np.array([r.iterations, r.terminated, r.n_pose_dim, r.n_landmarks, r.n_active_edges], np.int64)
        out[rname + "_chi2_final"] = np.float64(r.chi2_final)
    for kind in range(A.EDGE_KINDS):
        chi2, depth, norm = o.edge_chi2(kind)
        out[f"edge{kind}_chi2"] = chi2
        out[f"edge{kind}_depth"] = depth
        out[f"edge{kind}_flags"] = o.get_edge_flags(kind)
    return out


def main():
    for name in CASES:
        g, pts_only = make_graph(name)
        out = run_oracle(g, pts_only)
        arrays = {"in_" + k: v for k, v in g.a.items()}
        arrays.update({"out_" + k: v for k, v in out.items()})
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **arrays)
        print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB, round-2 chi2 {float(out['r2_chi2_final']):.6f}, outliers {out['outliers'].tolist()}")


if __name__ == "__main__":
    main()
