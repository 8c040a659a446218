"""Multi-GPU partitioning of the local BA.

* independent windows: window w -> rank w % world (no data-path collective);
* one large window: points (with their CSR edge rows) are cut into `world` contiguous slices of balanced
  edge count; key-frames, cuboids, planes and every non-point edge are replicated — rank 0 owns their
  contribution inside the engine (ppo_ba_set_shard).
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


def shard_graph(g, rank, world):
    """GraphArrays holding this rank's slice of the points and a replica of everything else."""
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
    return A.GraphArrays(**{k: np.ascontiguousarray(v) for k, v in a.items()}), (p0, p1), (e0, e1)
