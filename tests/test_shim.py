"""The host shim that replaces Optimizer::LocalBACameraPlaneCuboids / LocalBundleAdjustment,
exercised on a mock map with the reference's member names (csrc/host/ppo_mock_slam.h)."""
import numpy as np
import pytest


def _graph(ppo, **kw):
    return ppo.synth.make_graph(ppo.synth.config(1, n_kf=14, n_fixed=3, n_pt=900, n_pl=6, n_cu=3, **kw))


def test_flattening_reproduces_the_flat_graph_cpu(ppo):
    """With the stop flag set on entry the shim collects + flattens the window and returns before touching the GPU
    (Optimizer.cc:2723-2725): the flattened graph must equal the flat graph the mock map was built from."""
    import shim_lib
    g = _graph(ppo, corners_2d=1)
    st, counts, flat = shim_lib.run(g, stop=True)
    assert counts == [0, 0, 0, 0]  # nothing written back, nothing erased
    assert np.allclose(st.kf_pose, g["kf_pose"], atol=2e-7)  # untouched (float32 round trip of the pose)
    # points seen by no local key-frame are not part of the reference's local window: compare the others
    rp = g["pt_rowptr"]
    n_loc = int((g["kf_fixed"] == 0).sum()) + 1
    local_pts = [p for p in range(g.c.n_pt) if (g["pe_kf"][rp[p]:rp[p + 1]] < n_loc).any()]
    assert flat.c.n_pt == len(local_pts) and flat.c.n_kf == g.c.n_kf
    assert np.array_equal(flat["kf_fixed"], g["kf_fixed"]) and np.array_equal(flat["kf_intr"], g["kf_intr"])
    assert np.allclose(flat["kf_pose"], g["kf_pose"], atol=2e-7)
    # the shim orders points by discovery (key-frame by key-frame) — match them through their coordinates
    key = lambda a: [tuple(np.round(r, 6)) for r in a]
    order = {k: i for i, k in enumerate(key(flat["pt_xyz"]))}
    frp = flat["pt_rowptr"]
    for p in local_pts[::37]:
        q = order[tuple(np.round(g["pt_xyz"][p], 6))]
        assert np.array_equal(flat["pe_kf"][frp[q]:frp[q + 1]], g["pe_kf"][rp[p]:rp[p + 1]])
        assert np.array_equal(flat["pe_obs"][frp[q]:frp[q + 1]], g["pe_obs"][rp[p]:rp[p + 1]])
        assert np.allclose(flat["pe_invsigma2"][frp[q]:frp[q + 1]], g["pe_invsigma2"][rp[p]:rp[p + 1]], rtol=1e-6)
    # planes / cuboids and their edges (order of edges may differ: compare as multisets)
    assert flat.c.n_pl == g.c.n_pl and flat.c.n_cu == g.c.n_cu
    assert np.allclose(np.sort(flat["pl_coef"], axis=0), np.sort(g["pl_coef"], axis=0), atol=1e-6)
    assert np.allclose(np.sort(flat["cu_state"], axis=0), np.sort(g["cu_state"], axis=0), atol=1e-12)
    assert (flat.c.n_ple, flat.c.n_cbe, flat.c.n_pce, flat.c.n_cpe) == (g.c.n_ple, g.c.n_cbe, g.c.n_pce, g.c.n_cpe)
    assert sorted(flat["ple_kind"].tolist()) == sorted(g["ple_kind"].tolist())
    assert np.allclose(np.sort(flat["ple_info"], axis=0), np.sort(g["ple_info"], axis=0))
    assert np.allclose(np.sort(flat["cbe_meas"].ravel()), np.sort(g["cbe_meas"].ravel()))
    assert np.allclose(np.sort(flat["cbe_info"]), np.sort(g["cbe_info"]))
    assert np.allclose(np.sort(flat["cpe_info"].ravel()), np.sort(g["cpe_info"].ravel()))


def test_points_only_entry_point_flattening_cpu(ppo):
    import shim_lib
    g = ppo.synth.make_graph(ppo.synth.config(0))
    st, counts, flat = shim_lib.run(g, mixed=False, stop=True)
    assert flat.c.n_pl == 0 and flat.c.n_cu == 0 and flat.c.n_ple == 0 and flat.c.n_kf == g.c.n_kf
    assert abs(flat.c.n_pe - g.c.n_pe) < 0.02 * g.c.n_pe


@pytest.mark.gpu
@pytest.mark.parametrize("mixed", [True, False])
def test_shim_end_to_end_matches_oracle_on_its_own_flat_graph(ppo, oracle_mod, mixed):
    """LocalMapping-style call -> GPU engine -> write-back; compared with the oracle run on the graph the shim built."""
    import shim_lib
    g = _graph(ppo) if mixed else ppo.synth.make_graph(ppo.synth.config(0))
    st, counts, flat = shim_lib.run(g, mixed=mixed)
    assert shim_lib.lib().ppo_shim_last_rc() == 0
    o = oracle_mod.Oracle()
    if not mixed:
        o.params.solver = ppo.abi.SOLVER_6_3
    o.set_graph(flat)
    ro = o.local_ba()
    so = o.get_state()
    res = shim_lib.lib().ppo_shim_last_result().contents
    assert (res.round1.iterations, res.round2.iterations) == (ro.round1.iterations, ro.round2.iterations)
    assert np.isclose(res.round2.chi2_final, ro.round2.chi2_final, rtol=1e-6)
    # the map holds float32: compare at float32 resolution; key-frame slots are identical (sorted by mnId)
    assert np.abs(st.kf_pose - so.kf_pose).max() < 5e-6
    # every local key-frame got SetPose, every graph point UpdateNormalAndDepth
    n_local = int((flat["kf_fixed"] == 0).sum()) + 1
    assert counts[2] == n_local and counts[3] == flat.c.n_pt
    # erase decisions = oracle's final chi2 / depth tests (Optimizer.cc:2840-2887)
    chi2, dpos, _ = o.edge_chi2(ppo.abi.EDGE_POINT)
    mono = flat["pe_obs"][:, 2] < 0
    assert counts[0] == int(((chi2 > np.where(mono, 5.991, 7.815)) | (dpos == 0)).sum())
    if mixed:
        pchi, _, _ = o.edge_chi2(ppo.abi.EDGE_PLANE)
        assert counts[1] == int(((flat["ple_kind"] == 0) & (pchi > 500.0)).sum())
        # the shim numbers planes / cuboids in discovery order: match them to the map through their initial values
        def perm(flat_init, map_init):
            return [int(np.abs(map_init - r).sum(axis=1).argmin()) for r in flat_init]
        pp, pc = perm(flat["pl_coef"], g["pl_coef"]), perm(flat["cu_state"], g["cu_state"])
        assert sorted(pp) == list(range(g.c.n_pl)) and sorted(pc) == list(range(g.c.n_cu))
        assert np.abs(st.pl_coef[pp] - so.pl_coef).max() < 5e-6
        assert np.abs(st.cu_state[pc] - so.cu_state).max() < 1e-5
    # points: matched through the shim's ordering
    moved = np.abs(st.pt_xyz - g["pt_xyz"]).max(axis=1) > 0
    assert moved.sum() >= 0.95 * flat.c.n_pt


# ---- Optimizer::GlobalBundleAdjustemnt / BundleAdjustment ------------------------------------------------
def _global_graph(ppo):
    # points-only map; key-frame 0 (mnId 0) is the only fixed one, as in Optimizer.cc:81
    return ppo.synth.make_graph(ppo.synth.config(0, n_kf=12, n_fixed=1, n_pt=700))


def test_global_ba_flattening_cpu(ppo):
    """Without a GPU the shim still flattens the whole map, reports PPO_E_NOGPU and leaves the map untouched (no CPU
    fallback); with a GPU this test only checks the flattening."""
    import shim_lib
    g = _global_graph(ppo)
    st, counts, flat, rc = shim_lib.run_global(g, n_iterations=0, stop=True)
    assert flat.c.n_kf == g.c.n_kf and flat.c.n_pt == g.c.n_pt and flat.c.n_pe == g.c.n_pe
    assert flat.c.n_pl == 0 and flat.c.n_cu == 0 and flat.c.n_ple == 0 and flat.c.n_cbe == 0
    assert flat["kf_fixed"].tolist() == [1] + [0] * (g.c.n_kf - 1)
    assert np.allclose(flat["kf_pose"], g["kf_pose"], atol=2e-7) and np.array_equal(flat["kf_intr"], g["kf_intr"])
    assert np.allclose(flat["pt_xyz"], g["pt_xyz"].astype(np.float32), atol=0) and np.array_equal(flat["pt_rowptr"], g["pt_rowptr"])
    # observations of a point come out of a std::map keyed by KeyFrame*: same multiset per point
    rp = g["pt_rowptr"]
    for p in range(0, g.c.n_pt, 53):
        a = sorted(zip(flat["pe_kf"][rp[p]:rp[p + 1]].tolist(), map(tuple, flat["pe_obs"][rp[p]:rp[p + 1]].tolist())))
        b = sorted(zip(g["pe_kf"][rp[p]:rp[p + 1]].tolist(), map(tuple, g["pe_obs"][rp[p]:rp[p + 1]].tolist())))
        assert a == b
    import torch
    if not torch.cuda.is_available():
        assert rc == ppo.abi.PPO_E_NOGPU and counts == [0, 0, 0, 0]
        assert np.allclose(st.pt_xyz, g["pt_xyz"].astype(np.float32))


@pytest.mark.gpu
@pytest.mark.parametrize("n_loop_kf,robust", [(0, True), (7, True), (0, False)])
def test_global_ba_end_to_end_matches_oracle(ppo, oracle_mod, n_loop_kf, robust):
    """LoopClosing-style call (Optimizer.cc:46-241): one optimize(nIterations), no outlier pass, Huber delta sqrt(5.99),
    results in the map (nLoopKF == 0) or in mTcwGBA / mPosGBA."""
    import shim_lib
    g = _global_graph(ppo)
    n_it = 10
    st, counts, flat, rc = shim_lib.run_global(g, n_iterations=n_it, n_loop_kf=n_loop_kf, robust=robust)
    assert rc == 0
    p = oracle_mod.default_params()
    p.solver = ppo.abi.SOLVER_6_3
    p.huber_mono = float(np.float32(np.sqrt(5.99)))
    o = oracle_mod.Oracle(p)
    o.set_graph(flat)
    if not robust:
        o.set_edge_flags(ppo.abi.EDGE_POINT, np.zeros(flat.c.n_pe, np.uint8))
    so_stats = o.optimize(n_it)
    so = o.get_state()
    res = shim_lib.lib().ppo_shim_last_result().contents.round1
    assert res.iterations == so_stats.iterations and np.isclose(res.chi2_final, so_stats.chi2_final, rtol=1e-6)
    assert np.abs(st.kf_pose - so.kf_pose).max() < 5e-6  # the map holds float32
    assert np.abs(st.pt_xyz - so.pt_xyz).max() < 5e-5
    if n_loop_kf:
        assert counts == [g.c.n_kf, g.c.n_pt, 0, 0]  # tagged with nLoopKF, map itself untouched
    else:
        assert counts == [0, 0, g.c.n_kf, g.c.n_pt]  # SetPose on every key-frame, UpdateNormalAndDepth on every point
