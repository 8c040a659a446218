"""Single large window with its landmarks (points and planes) partitioned over 2 GPUs and its reduced system factorised by both
(ppo_dense.cu: distributed solve over peer memory) must reproduce the single-GPU solve.  Needs >= 2 GPUs (skipped otherwise)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, uid, q, cfg_kw):
    sys.path.insert(0, ROOT)
    import torch  # noqa: F401  (maps libnccl.so.2 into the process)
    from ppo_pkg import ppo
    g = ppo.synth.make_graph(ppo.synth.config(1, **cfg_kw))
    gs, (p0, p1), _ = ppo.sharding.shard_graph(g, rank, world)
    comm = ppo.nccl_init(uid, rank, world, rank)
    eng = ppo.LocalBA(device=rank)
    eng.set_shard(comm, rank, world)
    eng.set_graph(gs)
    # a stop flag that only ONE rank sees raised must stop all ranks together (it joins the entry all-reduce): no iteration anywhere, no hang
    flag = np.array([1 if rank == world - 1 else 0], np.uint8)
    st0 = eng.optimize(5, flag)
    assert st0.iterations == 0, st0.iterations
    eng.reset()
    res = eng.local_ba()
    st = eng.get_state()
    q0, q1 = ppo.sharding.plane_range(g, rank, world)
    out = dict(rank=rank, p0=p0, p1=p1, q0=q0, q1=q1, kf=st.kf_pose, pt=st.pt_xyz, pl=st.pl_coef, cu=st.cu_state, chi=res.round2.chi2_final,
               it=(res.round1.iterations, res.round2.iterations), coll=eng.collective_count(), n_out=res.n_outlier_point_edges)
    if rank == 0:
        ref = ppo.LocalBA(device=0)
        ref.set_graph(g)
        r = ref.local_ba()
        s = ref.get_state()
        out["ref"] = dict(kf=s.kf_pose, pt=s.pt_xyz, pl=s.pl_coef, cu=s.cu_state, chi=r.round2.chi2_final,
                          it=(r.round1.iterations, r.round2.iterations), n_out=r.n_outlier_point_edges)
    q.put(out)
    eng.close()
    ppo.nccl_destroy(comm)


def test_sharded_window_matches_single_gpu():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, ROOT)
    from ppo_pkg import ppo
    world = 2
    uid = ppo.nccl_unique_id()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    kw = dict(n_kf=24, n_pt=6000, n_pl=12, n_cu=4)
    procs = [ctx.Process(target=_worker, args=(r, world, uid, q, kw)) for r in range(world)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=300) for _ in range(world)], key=lambda o: o["rank"])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    ref = outs[0]["ref"]
    assert outs[0]["coll"] > 0 and outs[0]["it"] == outs[1]["it"] == ref["it"]
    assert sum(o["n_out"] for o in outs) == ref["n_out"]
    for o in outs:
        assert np.isclose(o["chi"], ref["chi"], rtol=1e-6)
        assert np.abs(o["kf"] - ref["kf"]).max() < 1e-6 and np.abs(o["cu"] - ref["cu"]).max() < 1e-5
        assert np.abs(o["pl"] - ref["pl"][o["q0"]:o["q1"]]).max() < 1e-6
        assert np.abs(o["pt"] - ref["pt"][o["p0"]:o["p1"]]).max() < 1e-5
