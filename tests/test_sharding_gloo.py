"""world_size-2 gloo tests (CPU) of the host-side multi-GPU logic:
 * window -> rank dealing for independent windows,
 * landmark sharding of one window: the reduced system is additive over landmark (point and plane) shards, i.e.
   all_reduce(sum) of the per-rank [Hschur | bschur | chi2] (replicated edges counted on rank 0 only)
   equals the unsharded system — the identity the NCCL path of the engine relies on.
The per-rank linear algebra is done by the CPU oracle (tests-only); the collective is torch.distributed/gloo."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ppo_pkg import ppo
    import oracle_lib
    A = ppo.abi
    g = ppo.synth.make_graph(ppo.synth.config(1, n_kf=10, n_fixed=2, n_pt=400, n_pl=5, n_cu=2))
    # unsharded reference
    o = oracle_lib.Oracle()
    o.set_graph(g)
    full = o.debug_linearize()
    lam = 1e-5 * max(np.abs(np.diag(full["Hpp"])).max(), np.abs(full["Hll"][:, [0, 4, 8]]).max())
    full_s = o.debug_solve(lam, full["n_p"], full["n_l"])
    # this rank's shard: its points and planes (with their edges) + replicas of the key-frames, cuboids and their edges, which rank 0 owns
    gs, (p0, p1), _ = ppo.sharding.shard_graph(g, rank, world)
    if rank != 0:
        # keep the replicated edges (so every vertex stays active, as the engine's activity all-reduce guarantees)
        # but with zero information: their contribution is owned by rank 0
        a = {k: v.copy() for k, v in gs.a.items()}
        for k in ("cbe_info",):
            a[k][...] = 0.0
        for k in ("pce_cuboid", "pce_rowptr", "pce_pts"):
            a.pop(k, None)
        gs = A.GraphArrays(**a)
    osh = oracle_lib.Oracle()
    osh.set_graph(gs)
    part = osh.debug_linearize()
    n_p = full["n_p"]
    ok = part["n_p"] == n_p  # every key-frame stays active in each shard of this synthetic window
    # lambda enters the reduced system through Hpp + lambda I (once) and through Dinv (per landmark)
    ps = osh.debug_solve(lam, part["n_p"], part["n_l"])
    S = np.triu(ps["Hschur"])
    if rank != 0:
        S -= lam * np.eye(n_p)  # only the owner adds lambda to the pose diagonal
    t = torch.from_numpy(np.concatenate([S.ravel(), ps["bschur"], [part["chi2"]]]))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    red = t.numpy()
    S_red, b_red, chi_red = red[:n_p * n_p].reshape(n_p, n_p), red[n_p * n_p:n_p * n_p + n_p], red[-1]
    sc = np.abs(full_s["Hschur"]).max()
    res = dict(ok=bool(ok), dS=float(np.abs(S_red - np.triu(full_s["Hschur"])).max() / sc),
               db=float(np.abs(b_red - full_s["bschur"]).max() / np.abs(full_s["bschur"]).max()),
               dchi=float(abs(chi_red - full["chi2"]) / full["chi2"]), points=(p0, p1),
               windows=ppo.sharding.windows_for_rank(64, rank, world))
    if rank == 0:
        gathered = [None] * world
        dist.gather_object(res, gathered)
        out.put(gathered)
    else:
        dist.gather_object(res)
    dist.barrier()
    dist.destroy_process_group()


def test_window_dealing_and_schur_additivity_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(r["ok"] for r in res)
    for r in res:
        assert r["dS"] < 1e-9 and r["db"] < 1e-9 and r["dchi"] < 1e-12, r
    # contiguous, disjoint, covering point slices; windows dealt round-robin
    assert res[0]["points"][0] == 0 and res[0]["points"][1] == res[1]["points"][0] and res[1]["points"][1] == 400
    assert sorted(res[0]["windows"] + res[1]["windows"]) == list(range(64))
    assert res[0]["windows"][:3] == [0, 2, 4] and res[1]["windows"][:3] == [1, 3, 5]


def test_point_slices_balance_edges():
    sys.path.insert(0, ROOT)
    from ppo_pkg import ppo
    g = ppo.synth.make_graph(ppo.synth.config(1, n_kf=10, n_fixed=2, n_pt=999, n_pl=0, n_cu=0))
    for world in (1, 2, 4, 8):
        sl = ppo.sharding.point_slices(g["pt_rowptr"], world)
        assert sl[0][0] == 0 and sl[-1][1] == 999 and all(a[1] == b[0] for a, b in zip(sl[:-1], sl[1:]))
        rp = g["pt_rowptr"]
        edges = [int(rp[b] - rp[a]) for a, b in sl]
        assert max(edges) - min(edges) <= 12  # one point's edges at most
        parts = [ppo.sharding.shard_graph(g, r, world)[0] for r in range(world)]
        assert sum(p.c.n_pe for p in parts) == g.c.n_pe and sum(p.c.n_pt for p in parts) == 999


def test_landmark_partition_is_exact():
    """shard_graph partitions the landmarks: every point edge, plane, plane edge and cuboid-plane edge goes to exactly one rank (planes balanced
    by their Schur pair count), key-frames / cuboids / camera-cuboid / point-cuboid edges are replicated."""
    sys.path.insert(0, ROOT)
    from ppo_pkg import ppo
    g = ppo.synth.make_graph(ppo.synth.config(1, n_kf=20, n_fixed=3, n_pt=1500, n_pl=13, n_cu=4))
    for world in (1, 2, 3, 8):
        parts = [ppo.sharding.shard_graph(g, r, world)[0] for r in range(world)]
        ranges = [ppo.sharding.plane_range(g, r, world) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == g.c.n_pl and all(a[1] == b[0] for a, b in zip(ranges[:-1], ranges[1:]))
        assert sum(p.c.n_pl for p in parts) == g.c.n_pl and sum(p.c.n_ple for p in parts) == g.c.n_ple
        assert sum(p.c.n_cpe for p in parts) == g.c.n_cpe and sum(p.c.n_pe for p in parts) == g.c.n_pe
        for r, p in enumerate(parts):
            q0, q1 = ranges[r]
            assert p.c.n_pl == q1 - q0 and np.array_equal(p["pl_coef"], g["pl_coef"][q0:q1])
            if p.c.n_ple:
                assert p["ple_plane"].min() >= 0 and p["ple_plane"].max() < p.c.n_pl
                m = (g["ple_plane"] >= q0) & (g["ple_plane"] < q1)
                assert np.array_equal(p["ple_meas"], g["ple_meas"][m]) and np.array_equal(p["ple_kf"], g["ple_kf"][m])
            if p.c.n_cpe:
                assert p["cpe_plane"].min() >= 0 and p["cpe_plane"].max() < p.c.n_pl
            # replicated parts
            assert p.c.n_kf == g.c.n_kf and p.c.n_cu == g.c.n_cu and p.c.n_cbe == g.c.n_cbe and p.c.n_pce == g.c.n_pce
        if world > 1:  # balance by pair count: no rank carries more than its share plus one plane's pairs
            u = np.unique(np.stack([g["ple_plane"], g["ple_kf"]], 1), axis=0)
            k = np.bincount(u[:, 0], minlength=g.c.n_pl).astype(np.int64)
            cost = k * (k + 1) // 2
            loads = [int(cost[a:b].sum()) for a, b in ranges]
            assert max(loads) <= cost.sum() / world + cost.max()
