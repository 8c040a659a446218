"""First-contact diagnostics on the GPU box: engine vs oracle at every level, printing errors instead of asserting."""
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from ppo_pkg import ppo  # noqa: E402
import oracle_lib  # noqa: E402

A = ppo.abi


def rel(a, b):
    d = np.abs(np.asarray(a) - np.asarray(b))
    return float(d.max() / max(1e-300, np.abs(b).max())) if d.size else 0.0


def level(cfg, name):
    print(f"==== {name}", flush=True)
    g = ppo.synth.make_graph(cfg)
    print({k: getattr(g.c, k) for k in A._COUNTS})
    o = oracle_lib.Oracle()
    e = ppo.LocalBA()
    o.set_graph(g)
    e.set_graph(g)
    lo, le = o.debug_linearize(), e.debug_linearize()
    print("dims", (lo["n_p"], lo["n_l"]), (le["n_p"], le["n_l"]), "chi2", lo["chi2"], le["chi2"], "rel", abs(lo["chi2"] - le["chi2"]) / lo["chi2"])
    if (lo["n_p"], lo["n_l"]) == (le["n_p"], le["n_l"]):
        print("Hpp rel", rel(np.triu(le["Hpp"]), np.triu(lo["Hpp"])), "Hll rel", rel(le["Hll"], lo["Hll"]), "b rel", rel(le["b"], lo["b"]))
        n_p = lo["n_p"]
        dH = np.abs(np.triu(le["Hpp"]) - np.triu(lo["Hpp"]))
        if dH.size:
            i, j = np.unravel_index(dH.argmax(), dH.shape)
            print("  worst Hpp entry", (i, j), le["Hpp"][i, j], lo["Hpp"][i, j])
        db = np.abs(le["b"] - lo["b"])
        k = db.argmax()
        print("  worst b entry", k, "(pose part)" if k < n_p else "(landmark %d)" % ((k - n_p) // 3), le["b"][k], lo["b"][k])
        lam = 1e-5 * max(np.abs(np.diag(lo["Hpp"])).max() if n_p else 0, np.abs(lo["Hll"][:, [0, 4, 8]]).max())
        so, se = o.debug_solve(lam, lo["n_p"], lo["n_l"]), e.debug_solve(lam, le["n_p"], le["n_l"])
        print("solve ok", so["ok"], se["ok"], "Hschur rel", rel(np.triu(se["Hschur"]), np.triu(so["Hschur"])), "bschur rel", rel(se["bschur"], so["bschur"]),
              "xp rel", rel(se["x"][:n_p], so["x"][:n_p]), "xl rel", rel(se["x"][n_p:], so["x"][n_p:]))
    o.reset(); e.reset()
    t = time.time(); ro = o.local_ba(); to = time.time() - t
    t = time.time(); re_ = e.local_ba(); te = time.time() - t
    print(f"local_ba: oracle {to:.3f}s engine {te:.3f}s (device ms {re_.round1.ms_total + re_.round2.ms_total:.2f})")
    for nm, a, b in (("round1", ro.round1, re_.round1), ("round2", ro.round2, re_.round2)):
        print(nm, "iters", a.iterations, b.iterations, "term", a.terminated, b.terminated, "np", a.n_pose_dim, b.n_pose_dim, "nl", a.n_landmarks, b.n_landmarks,
              "edges", a.n_active_edges, b.n_active_edges, "chi2", a.chi2_final, b.chi2_final)
        for x, y in zip(a.trace_list(), b.trace_list()):
            print("   o", x)
            print("   e", y)
    print("outliers", (ro.n_outlier_point_edges, ro.n_outlier_plane_edges, ro.n_outlier_cuboid_edges), (re_.n_outlier_point_edges, re_.n_outlier_plane_edges, re_.n_outlier_cuboid_edges))
    so, se = o.get_state(), e.get_state()
    for nm in ("kf_pose", "pt_xyz", "pl_coef", "cu_state"):
        a, b = getattr(se, nm), getattr(so, nm)
        if a.size:
            print(" state", nm, "max abs diff", float(np.abs(a - b).max()), "rel", float((np.abs(a - b) / np.maximum(np.abs(b), 1)).max()))
    e.close()


if __name__ == "__main__":
    cases = [
        (ppo.synth.config(0), "config0 points-only"),
        (ppo.synth.config(1, n_kf=8, n_fixed=2, n_pt=300, n_pl=4, n_cu=3, corners_2d=1), "tiny mixed"),
        (ppo.synth.config(1), "config1"),
    ]
    for cfg, name in cases:
        try:
            level(cfg, name)
        except Exception:
            traceback.print_exc()
