"""Multi-GPU partitioning of the local BA.

* independent windows: window w -> rank w % world (no data-path collective);
* one large window: the LANDMARKS are partitioned -- points (with their CSR edge rows) into `world` contiguous
  slices of balanced edge count, planes (with their plane edges and the cuboid-plane edges that name them) into
  `world` contiguous slices of balanced Schur cost (a plane seen from k key-frames contributes k (k + 1) / 2
  block pairs to the reduced system); key-frames, cuboids and the edges among them (camera-cuboid,
  point-cuboid) are replicated -- rank 0 owns their contribution inside the engine (ppo_ba_set_shard).
"""
import numpy as np

from . import _abi as A


def windows_for_rank(n_windows, rank, world):
    return list(range(rank, n_windows, world))


def point_slices(pt_rowptr, world):
    """Contiguous point ranges [p0, p1) per rank with (almost) equal numbers of edges."""
    rp = np.asarray(pt_rowptr, dtype=np.int64)
    n_pt, n_pe = len(rp) - 1, int(rp[-1])
    cuts = [0]
    for r in range(1, world):
        target = n_pe * r // world
        cuts.append(int(np.searchsorted(rp, target, side="left")))
    cuts.append(n_pt)
    cuts = np.maximum.accumulate(np.clip(cuts, 0, n_pt))
    return [(int(cuts[r]), int(cuts[r + 1])) for r in range(world)]


def plane_slices(ple_plane, ple_kf, n_pl, world):
    """Contiguous plane ranges [q0, q1) per rank with (almost) equal numbers of Schur block pairs."""
    cost = np.zeros(n_pl, np.int64)
    if n_pl and len(ple_plane):
        u = np.unique(np.stack([np.asarray(ple_plane, np.int64), np.asarray(ple_kf, np.int64)], 1), axis=0)
        k = np.bincount(u[:, 0], minlength=n_pl).astype(np.int64)
        cost = k * (k + 1) // 2
    cum = np.concatenate([[0], np.cumsum(cost)])
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(cum, cum[-1] * r // world, side="left")))
    cuts.append(n_pl)
    cuts = np.maximum.accumulate(np.clip(cuts, 0, n_pl))
    return [(int(cuts[r]), int(cuts[r + 1])) for r in range(world)]


def shard_graph(g, rank, world):
    """GraphArrays holding this rank's slice of the points and of the planes, and a replica of everything else."""
    p0, p1 = point_slices(g["pt_rowptr"], world)[rank]
    rp = g["pt_rowptr"].astype(np.int64)
    e0, e1 = int(rp[p0]), int(rp[p1])
    a = {k: v for k, v in g.a.items()}
    a = dict(a)
    a["pt_xyz"] = g["pt_xyz"][p0:p1]
    if "pt_fixed" in g.a:
        a["pt_fixed"] = g["pt_fixed"][p0:p1]
    a["pt_rowptr"] = (rp[p0:p1 + 1] - e0).astype(np.int32)
    a["pe_kf"] = g["pe_kf"][e0:e1]
    a["pe_obs"] = g["pe_obs"][e0:e1]
    a["pe_invsigma2"] = g["pe_invsigma2"][e0:e1]
    n_pl = len(g["pl_coef"]) if "pl_coef" in g.a else 0
    if n_pl:
        has_ple = "ple_plane" in g.a and len(g["ple_plane"])
        q0, q1 = plane_slices(g["ple_plane"] if has_ple else [], g["ple_kf"] if has_ple else [], n_pl, world)[rank]
        a["pl_coef"] = g["pl_coef"][q0:q1]
        if "ple_plane" in g.a:
            m = (g["ple_plane"] >= q0) & (g["ple_plane"] < q1)
            for k in ("ple_plane", "ple_kf", "ple_kind", "ple_meas", "ple_info"):
                a[k] = g[k][m]
            a["ple_plane"] = a["ple_plane"] - q0
        if "cpe_plane" in g.a:
            m = (g["cpe_plane"] >= q0) & (g["cpe_plane"] < q1)
            for k in ("cpe_cuboid", "cpe_plane", "cpe_meas", "cpe_info"):
                a[k] = g[k][m]
            a["cpe_plane"] = a["cpe_plane"] - q0
    return A.GraphArrays(**{k: np.ascontiguousarray(v) for k, v in a.items()}), (p0, p1), (e0, e1)


def plane_range(g, rank, world):
    """[q0, q1) of the planes shard_graph gives to `rank`."""
    n_pl = len(g["pl_coef"]) if "pl_coef" in g.a else 0
    if not n_pl:
        return (0, 0)
    has_ple = "ple_plane" in g.a and len(g["ple_plane"])
    return plane_slices(g["ple_plane"] if has_ple else [], g["ple_kf"] if has_ple else [], n_pl, world)[rank]
