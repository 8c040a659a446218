"""The host shim that replaces Optimizer::LocalBACameraPlaneCuboids / LocalBundleAdjustment,
exercised on a mock map with the reference's member names (csrc/host/ppo_mock_slam.h)."""
import numpy as np
import pytest


def _graph(ppo, **kw):
    return ppo.synth.make_graph(ppo.synth.config(1, n_kf=14, n_fixed=3, n_pt=900, n_pl=6, n_cu=3, **kw))


def _assert_flat_matches_source(g, st, counts, flat, landmarks_all_local=True):
    """The flattened graph of a stop-at-entry call must equal the flat graph the mock map was built from."""
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
    if not landmarks_all_local:  # (a plane / cuboid seen by fixed key-frames only is not part of the window: the generic graph may have some)
        return
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


def test_flattening_reproduces_the_flat_graph_cpu(ppo):
    """With the stop flag set on entry the shim collects + flattens the window and returns before touching the GPU
    (Optimizer.cc:2723-2725): the flattened graph must equal the flat graph the mock map was built from."""
    import shim_lib
    g = _graph(ppo, corners_2d=1)
    st, counts, flat = shim_lib.run(g, stop=True)
    _assert_flat_matches_source(g, st, counts, flat)


@pytest.mark.parametrize("mixed", [True, False])
def test_threaded_flattening_equals_the_serial_loops_cpu(ppo, oracle_mod, mixed):
    """Stage B spreads the point / point-edge loops over host threads once a window has more than 4096 local map points: the
    flattened graph must be bit-identical to the one the serial loops produce (stop flag set: collection + flattening only)."""
    import shim_lib
    L = shim_lib.oracle_backed_lib()
    g = ppo.synth.make_graph(ppo.synth.config(1, n_kf=24, n_fixed=5, n_pt=12000, n_pl=70, n_cu=3))
    flats = []
    try:
        for n in (1, 5, 8, 3, 8):
            L.ppo_shim_set_threads(n)
            assert L.ppo_shim_get_threads() == n
            st, counts, flat = shim_lib.run(g, mixed=mixed, stop=True, backend=L)
            if mixed:  # (also against the source: a window of this size takes the batched mirror rebuild and the threaded collection)
                _assert_flat_matches_source(g, st, counts, flat, landmarks_all_local=False)
            flats.append(flat)
    finally:
        L.ppo_shim_set_threads(1)
    ref = flats[0]
    assert ref.c.n_pt > 4096 and ref.c.n_pe > 4 * ref.c.n_pt
    for f in flats[1:]:
        assert (f.c.n_kf, f.c.n_pt, f.c.n_pe, f.c.n_ple, f.c.n_cbe, f.c.n_pce, f.c.n_cpe) == (
            ref.c.n_kf, ref.c.n_pt, ref.c.n_pe, ref.c.n_ple, ref.c.n_cbe, ref.c.n_pce, ref.c.n_cpe)
        for k in ("pt_xyz", "pt_fixed", "pt_rowptr", "pe_kf", "pe_obs", "pe_invsigma2", "kf_pose", "kf_fixed"):
            assert np.array_equal(f[k], ref[k]), k
        if mixed:
            # plane edges are flattened by ranges of planes and appended in plane order; inside a plane the order is that of a std::map keyed
            # by KeyFrame pointers (as in the reference), i.e. of the heap addresses of this run's mock map: compare plane by plane as sets
            def rows(x):
                m = np.column_stack([x["ple_plane"], x["ple_kf"], x["ple_kind"], x["ple_meas"].reshape(-1, 4), x["ple_info"].reshape(-1, 3)])
                return m[np.lexsort(m.T[::-1])]
            assert np.array_equal(f["ple_plane"], ref["ple_plane"])  # (plane order itself is fixed)
            assert np.array_equal(rows(f), rows(ref))


def test_threaded_flattening_on_one_map_is_bit_identical_cpu(ppo, oracle_mod):
    """Same mock map (same heap addresses, hence the same std::map orders), flattened with 1, 3 and 8 host threads: every array of the
    flat graph, plane / cuboid edges included, must come out identical."""
    import shim_lib
    L = shim_lib.oracle_backed_lib()
    g = ppo.synth.make_graph(ppo.synth.config(1, n_kf=24, n_fixed=5, n_pt=12000, n_pl=70, n_cu=3))
    W = shim_lib.World(g, backend=L)
    flats = []
    try:
        for n in (1, 3, 8, 1):
            L.ppo_shim_set_threads(n)
            flats.append(W.run(stop=True)[2])
    finally:
        L.ppo_shim_set_threads(1)
        W.close()
    ref = flats[0]
    assert ref.c.n_ple >= 300 and ref.c.n_pt > 4096
    for f in flats[1:]:
        assert set(f.a) == set(ref.a)
        for k in ref.a:
            assert np.array_equal(f[k], ref[k]), k


def test_threaded_write_back_equals_the_serial_one_cpu(ppo, oracle_mod):
    """The full call (flattening, oracle-backed solve, erase lists, write-back) on a window large enough for the host thread pool:
    the map after the call must be the same whether the point loops ran on one thread or on several."""
    import shim_lib
    L = shim_lib.oracle_backed_lib()
    g = ppo.synth.make_graph(ppo.synth.config(1, n_kf=10, n_fixed=2, n_pt=5000, n_pl=4, n_cu=2))
    outs = []
    try:
        for n in (1, 4):
            L.ppo_shim_set_threads(n)
            st, counts, flat = shim_lib.run(g, backend=L)
            outs.append((st, counts, flat))
    finally:
        L.ppo_shim_set_threads(1)
    (sa, ca, fa), (sb, cb, fb) = outs
    assert fa.c.n_pt > 4096
    assert ca == cb and ca[0] > 0  # same erase lists (and some observations were erased)
    for k in ("kf_pose", "pt_xyz", "pl_coef", "cu_state"):
        assert np.array_equal(getattr(sa, k), getattr(sb, k)), k


def test_points_only_entry_point_flattening_cpu(ppo):
    import shim_lib
    g = ppo.synth.make_graph(ppo.synth.config(0))
    st, counts, flat = shim_lib.run(g, mixed=False, stop=True)
    assert flat.c.n_pl == 0 and flat.c.n_cu == 0 and flat.c.n_ple == 0 and flat.c.n_kf == g.c.n_kf
    assert abs(flat.c.n_pe - g.c.n_pe) < 0.02 * g.c.n_pe


def _backend(name):
    import shim_lib
    return shim_lib.lib() if name == "engine" else shim_lib.oracle_backed_lib()


BACKENDS = [pytest.param("engine", marks=pytest.mark.gpu), "oracle"]  # "oracle": the shim's host logic, CPU only (test build)


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("mixed", [True, False])
def test_shim_end_to_end_matches_oracle_on_its_own_flat_graph(ppo, oracle_mod, mixed, backend):
    """LocalMapping-style call -> engine -> write-back; compared with the oracle run on the graph the shim built."""
    import shim_lib
    L = _backend(backend)
    g = _graph(ppo) if mixed else ppo.synth.make_graph(ppo.synth.config(0))
    st, counts, flat = shim_lib.run(g, mixed=mixed, backend=L)
    assert L.ppo_shim_last_rc() == 0
    o = oracle_mod.Oracle()
    if not mixed:
        o.params.solver = ppo.abi.SOLVER_6_3
    o.set_graph(flat)
    ro = o.local_ba()
    so = o.get_state()
    res = L.ppo_shim_last_result().contents
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


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("n_loop_kf,robust", [(0, True), (7, True), (0, False)])
def test_global_ba_end_to_end_matches_oracle(ppo, oracle_mod, n_loop_kf, robust, backend):
    """LoopClosing-style call (Optimizer.cc:46-241): one optimize(nIterations), no outlier pass, Huber delta sqrt(5.99),
    results in the map (nLoopKF == 0) or in mTcwGBA / mPosGBA."""
    import shim_lib
    g = _global_graph(ppo)
    n_it = 10
    L = _backend(backend)
    st, counts, flat, rc = shim_lib.run_global(g, n_iterations=n_it, n_loop_kf=n_loop_kf, robust=robust, backend=L)
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
    res = L.ppo_shim_last_result().contents.round1
    assert res.iterations == so_stats.iterations and np.isclose(res.chi2_final, so_stats.chi2_final, rtol=1e-6)
    assert np.abs(st.kf_pose - so.kf_pose).max() < 5e-6  # the map holds float32
    assert np.abs(st.pt_xyz - so.pt_xyz).max() < 5e-5
    if n_loop_kf:
        assert counts == [g.c.n_kf, g.c.n_pt, 0, 0]  # tagged with nLoopKF, map itself untouched
    else:
        assert counts == [0, 0, g.c.n_kf, g.c.n_pt]  # SetPose on every key-frame, UpdateNormalAndDepth on every point


# ---- Optimizer::PoseOptimization ------------------------------------------------------------------------------
@pytest.mark.parametrize("backend", BACKENDS)
def test_pose_optimization_matches_the_reference_schedule(ppo, oracle_mod, backend):
    """Tracking-style call (Optimizer.cc:247-459).  The reference's schedule is restated here in Python on top of the
    oracle's fine-grained API (4 x [restart from mTcw, optimize(10), computeError on the outliers, float chi2 test],
    robust kernels dropped after the third round) and run on the graph the shim flattened."""
    import shim_lib
    A = ppo.abi
    L = _backend(backend)
    g, truth = ppo.synth.make_graph(ppo.synth.config(0, n_kf=10, n_fixed=2, n_pt=1500), with_truth=True)
    a = {k: v.copy() for k, v in g.a.items()}
    a["pt_xyz"] = truth.pt_xyz.copy()  # a tracked frame sees an already optimised map; its own pose is the noisy prediction
    g = A.GraphArrays(**a)
    kf = 6
    pose, outl, counts, flat, rc = shim_lib.run_pose(g, kf, backend=L)
    assert rc == 0
    n = int((g["pe_kf"] == kf).sum())
    assert n > 100 and counts[2] == n and counts[1] == 1
    # the flattened problem: one free pose, every associated map point fixed, one edge each (features without a point skipped)
    assert (flat.c.n_kf, flat.c.n_pt, flat.c.n_pe) == (1, n, n) and flat["pt_fixed"].all() and not flat["kf_fixed"].any()
    sel = np.flatnonzero(g["pe_kf"] == kf)
    assert np.array_equal(flat["pe_obs"], g["pe_obs"][sel])
    o = oracle_mod.Oracle()
    o.params.solver = A.SOLVER_6_3
    o.set_graph(flat)
    flags = np.full(n, A.EF_ROBUST, np.uint8)
    mono = flat["pe_obs"][:, 2] < 0
    th = np.where(mono, np.float32(5.991), np.float32(7.815))
    n_bad = 0
    for it in range(4):
        o.reset()
        o.set_edge_flags(A.EDGE_POINT, flags)
        o.optimize(10)
        o.recompute_edge_errors(A.EDGE_POINT)
        chi2, _, _ = o.edge_chi2(A.EDGE_POINT)
        out = chi2.astype(np.float32) > th
        n_bad = int(out.sum())
        flags = ((flags & A.EF_ROBUST) | np.where(out, A.EF_LEVEL1, 0)).astype(np.uint8)
        if it == 2:
            flags &= np.uint8(~A.EF_ROBUST & 0xFF)
    want = o.get_state().kf_pose[0]
    assert np.array_equal(outl.astype(bool), out)
    assert counts[0] == n - n_bad
    assert np.abs(pose - want).max() < 5e-6  # Frame::mTcw is float32
    # and it does what it is for: the gross outliers of the synthetic window are rejected, the pose moves
    assert 0.01 < n_bad / n < 0.2
    assert np.abs(pose - truth.kf_pose[kf]).max() < 0.2 * np.abs(g["kf_pose"][kf] - truth.kf_pose[kf]).max()


def test_pose_optimization_needs_three_correspondences(ppo):
    import shim_lib
    g = ppo.synth.make_graph(ppo.synth.config(0, n_kf=4, n_fixed=1, n_pt=2))
    pose, outl, counts, flat, rc = shim_lib.run_pose(g, 1, backend=shim_lib.oracle_backed_lib())
    assert counts[0] == 0 and counts[1] == 0  # Optimizer.cc:371-372: returns 0, no SetPose


def test_bad_points_and_keyframes_are_left_out_and_untouched(ppo, oracle_mod):
    """isBad() map points (Optimizer.cc:2040-2047) and key-frames (:2008-2010, :2063-2068, :2352) never reach the graph and are
    not written back; the run equals the oracle on the graph the shim flattened.  Oracle-backed shim: CPU only."""
    import shim_lib
    L = shim_lib.oracle_backed_lib()
    g = _graph(ppo)
    bad_kf = int(np.flatnonzero(g["kf_fixed"] == 0)[3])  # a free (local) key-frame, not the one the BA is called for
    L.ppo_mock_set_options(5, bad_kf)
    try:
        st, counts, flat = shim_lib.run(g, backend=L)
    finally:
        L.ppo_mock_set_options(0, -1)
    assert L.ppo_shim_last_rc() == 0
    assert flat.c.n_kf == g.c.n_kf - 1  # the bad key-frame has no vertex
    bad_pts = np.arange(g.c.n_pt) % 5 == 4
    good_xyz = {tuple(np.round(r, 6)) for r in g["pt_xyz"][~bad_pts].astype(np.float32)}
    assert all(tuple(np.round(r, 6)) in good_xyz for r in flat["pt_xyz"].astype(np.float32))  # no bad point in the graph
    assert np.array_equal(st.pt_xyz[bad_pts], g["pt_xyz"][bad_pts].astype(np.float32))      # and none was moved
    assert np.allclose(st.kf_pose[bad_kf], g["kf_pose"][bad_kf], atol=2e-7)                 # nor the bad key-frame
    assert (flat["pe_kf"] < flat.c.n_kf).all()
    o = oracle_mod.Oracle()
    o.set_graph(flat)
    ro = o.local_ba()
    res = L.ppo_shim_last_result().contents
    assert (res.round1.iterations, res.round2.iterations) == (ro.round1.iterations, ro.round2.iterations)
    assert np.isclose(res.round2.chi2_final, ro.round2.chi2_final, rtol=1e-9)
    keep = np.arange(g.c.n_kf) != bad_kf
    assert np.abs(st.kf_pose[keep] - o.get_state().kf_pose).max() < 5e-6


@pytest.mark.parametrize("fix_camera,fix_point", [(True, False), (False, True)])
def test_fix_camera_and_fix_point_arguments(ppo, oracle_mod, fix_camera, fix_point):
    """LocalBACameraPlaneCuboids(pKF, stop, map, fixCamera, fixPoint) (Optimizer.cc:2126-2128, 2155): the flag fixes every
    local key-frame / every map point vertex; the other side is still optimised.  Oracle-backed shim: CPU only."""
    import shim_lib
    L = shim_lib.oracle_backed_lib()
    g = _graph(ppo)
    st, counts, flat = shim_lib.run(g, fix_camera=fix_camera, fix_point=fix_point, backend=L)
    assert L.ppo_shim_last_rc() == 0
    if fix_camera:
        assert flat["kf_fixed"].all()
        assert np.allclose(st.kf_pose, g["kf_pose"], atol=2e-7)  # SetPose with the unchanged estimate (float32 round trip)
        assert (np.abs(st.pt_xyz - g["pt_xyz"]).max(axis=1) > 1e-6).mean() > 0.9  # the points are still optimised
    else:
        assert flat["pt_fixed"].all()
        assert np.allclose(st.pt_xyz, g["pt_xyz"].astype(np.float32), atol=0)
        n_free = int((flat["kf_fixed"] == 0).sum())
        assert (np.abs(st.kf_pose - g["kf_pose"]).max(axis=1) > 1e-6).sum() >= n_free - 1
    o = oracle_mod.Oracle()
    o.set_graph(flat)
    ro = o.local_ba()
    res = L.ppo_shim_last_result().contents
    assert (res.round1.iterations, res.round2.iterations) == (ro.round1.iterations, ro.round2.iterations)
    assert np.abs(st.kf_pose - o.get_state().kf_pose).max() < 5e-6


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_random_windows_through_the_oracle_backed_shim(ppo, oracle_mod, seed):
    """Differential test on random window shapes: mock map -> shim (flatten, engine = oracle, write-back) against the oracle
    run directly on the graph the shim flattened."""
    import shim_lib
    L = shim_lib.oracle_backed_lib()
    rng = np.random.default_rng(100 + seed)
    cfg = dict(n_kf=int(rng.integers(6, 20)), n_fixed=int(rng.integers(1, 4)), n_pt=int(rng.integers(200, 1500)),
               n_pl=int(rng.integers(1, 8)), n_cu=int(rng.integers(1, 4)), corners_2d=int(rng.integers(0, 2)), cuboid_2d=int(rng.integers(0, 2)))
    g = ppo.synth.make_graph(ppo.synth.config(1, window=seed, **cfg))
    st, counts, flat = shim_lib.run(g, backend=L)
    assert L.ppo_shim_last_rc() == 0, cfg
    o = oracle_mod.Oracle()
    o.set_graph(flat)
    ro = o.local_ba()
    res = L.ppo_shim_last_result().contents
    assert (res.round1.iterations, res.round2.iterations) == (ro.round1.iterations, ro.round2.iterations), cfg
    assert np.isclose(res.round2.chi2_final, ro.round2.chi2_final, rtol=1e-9)
    assert np.abs(st.kf_pose - o.get_state().kf_pose).max() < 5e-6
    assert counts[3] == flat.c.n_pt


def test_observation_mirror_across_calls_cpu(ppo, oracle_mod):
    """SURVEY 8f rank 1, second half: consecutive local-BA calls on ONE map.  A later call must take the observation rows of the
    unchanged map points from the mirror (no std::map copies), rebuild exactly the rows whose MapPoint::mnObsVersion moved, and
    flatten the same graph a cold shim flattens from the same map state."""
    import shim_lib
    L = shim_lib.oracle_backed_lib()
    g = _graph(ppo)
    w = shim_lib.World(g, backend=L)
    try:
        s0 = w.mirror_stats()  # (counters run over the life of the process)
        _, c1, flat1 = w.run()
        s1 = w.mirror_stats()
        n_local = s1["rebuilt"] - s0["rebuilt"]  # cold: every local map point read from the map once
        assert s1["reused"] == s0["reused"] and n_local >= flat1.c.n_pt > 0
        # between the calls: the BA's own erasures (c1[0]) changed some rows; the tracker drops two more observations and moves a point
        rp = g["pt_rowptr"]
        touched = sum(w.erase_observation(p, int(g["pe_kf"][rp[p]])) for p in (5, 17))
        w.perturb_point(40, 0.01)  # positions are not part of a row
        # stop flag raised on entry: the shim collects and flattens, then returns without touching the map (Optimizer.cc:2723-2725)
        _, c2, flat2 = w.run(stop=True)
        s2 = w.mirror_stats()
        reused, rebuilt = s2["reused"] - s1["reused"], s2["rebuilt"] - s1["rebuilt"]
        assert c2 == c1 and reused + rebuilt == n_local  # (the counters are cumulative over the life of the map: nothing written, nothing erased)
        assert touched == 2 and 2 <= rebuilt <= 2 + c1[0]  # only the rows whose observations changed
        w.mirror_clear()
        _, _, flat3 = w.run(stop=True)  # the same map state through a cold mirror
        s3 = w.mirror_stats()
        assert s3["rebuilt"] - s2["rebuilt"] == n_local
        # and a normal call on the warm mirror still works
        _, c4, _ = w.run()
        assert c4[2] > c1[2]
    finally:
        w.close()
    assert flat2.c.n_pe == flat1.c.n_pe - c1[0] - touched  # the erased observations are gone from the graph
    assert set(flat2.a) == set(flat3.a)
    for k in flat2.a:
        assert np.array_equal(flat2[k], flat3[k]), k


# ---- Optimizer::LocalBACameraPointCuboids2D ------------------------------------------------------------------
@pytest.mark.parametrize("backend", BACKENDS)
def test_point_cuboids_2d_entry_point(ppo, oracle_mod, backend):
    """The sibling BA of the plane BA (Optimizer.cc:1252-1992): no plane vertices or edges even if the map has planes, bbox edges and one
    9-D EdgeSE3Cuboid per cuboid observation whose measurement is the LANDMARK's cuboid_local_meas (the reference reads
    pKFi->mvpMapCuboid[idx], :1779-1785); result = the oracle on the graph the shim flattened."""
    import shim_lib
    L = _backend(backend)
    A = ppo.abi
    g = ppo.synth.make_graph(ppo.synth.config(1, n_kf=14, n_fixed=3, n_pt=900, n_pl=6, n_cu=3, cuboid_3d=1))
    assert g.c.n_pl == 6 and g.c.n_ple > 0  # the map has planes; this entry point ignores them
    st, counts, flat = shim_lib.run(g, mixed=2, backend=L)
    assert L.ppo_shim_last_rc() == 0
    assert flat.c.n_pl == 0 and flat.c.n_ple == 0 and flat.c.n_cpe == 0 and flat.c.n_cu == g.c.n_cu
    kinds = flat["cbe_kind"]
    n_se3, n_bbox = int((kinds == A.CUBOID_SE3).sum()), int((kinds == A.CUBOID_BBOX).sum())
    assert n_se3 > 0 and n_se3 >= n_bbox  # every cuboid observation gets an SE3 edge; the bbox edges also pass the image-margin test
    # one measurement per landmark, information (ba_weight_SE3 * 0.75)^2
    se3 = kinds == A.CUBOID_SE3
    assert np.allclose(flat["cbe_info"][se3], 0.75 ** 2)
    for c in range(flat.c.n_cu):
        m = flat["cbe_meas"][se3 & (flat["cbe_cuboid"] == c)]
        assert len(m) == 0 or np.abs(m - m[0]).max() == 0.0
    o = oracle_mod.Oracle()
    o.set_graph(flat)
    ro = o.local_ba()
    so = o.get_state()
    res = L.ppo_shim_last_result().contents
    assert (res.round1.iterations, res.round2.iterations) == (ro.round1.iterations, ro.round2.iterations)
    assert np.isclose(res.round2.chi2_final, ro.round2.chi2_final, rtol=1e-6)
    assert np.abs(st.kf_pose - so.kf_pose).max() < 5e-6
    pc = [int(np.abs(g["cu_state"] - r).sum(axis=1).argmin()) for r in flat["cu_state"]]
    assert sorted(pc) == list(range(g.c.n_cu)) and np.abs(st.cu_state[pc] - so.cu_state).max() < 1e-5
    assert np.abs(st.pl_coef - g["pl_coef"]).max() < 1e-6  # planes untouched (float32 round trip of the map)
