#!/usr/bin/env python
"""bench.py — LM iterations/sec (and KF-windows/sec) of the local bundle adjustment on synthetic windows.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config 2] [--impl ours|reference]

A "step" is one complete local BA call (optimize(5) -> outlier pass -> optimize(10), the reference
schedule of Optimizer.cc:2727-2837) on one synthetic window per GPU.  N=1 workload: BASELINE.json
configs[2] (200 KF / 80k points / 200 planes / 50 cuboids), the window the north_star target
(>= 50 LM it/s on 1 GPU) is quoted on.  N>1: every rank solves its own window of the same size
(independent key-frame windows, no data-path collective) -> weak scaling; the same JSON line also
carries `config3` (BASELINE configs[3]: 64 config-1 windows dealt w mod N, 8 in flight per GPU) and
`strong` (configs[4]: ONE 1000-KF window, points sharded over the ranks, NCCL all-reduce of the reduced
system; checked against the single-GPU solve in the same run).
`value`  : device-resident — graph already in HBM, timed with CUDA events on the engine's stream.
`e2e`    : through the C-ABI with host buffers: ppo_ba_set_graph (H2D) + ppo_ba_local_ba + ppo_ba_get_state (D2H).
`e2e_shim`: through the reference-facing boundary itself: Optimizer::LocalBACameraPlaneCuboids on a mock map of the
           same size (window collection + flattening + engine + write-back), host time included.
`--impl reference`: the CPU oracle (single thread, the reference's own configuration: g2o OpenMP off), the SAME full call.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FP64_TFLOPS = 37.2  # measured on this pool's B200 with tools/ubench/fp64.cu (DMMA == DFMA rate)
METRIC = "LM iterations/sec"
UNIT = "LM it/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-batched", action="store_true", help="skip the extra concurrent-windows throughput measurement")
    ap.add_argument("--no-extras", action="store_true", help="N > 1: skip the config3 / strong (config 4 sharded) objects")
    ap.add_argument("--shard", action="store_true",
                    help="ONE window for the whole job, its points sharded over the ranks (NCCL all-reduce of the reduced system): strong scaling")
    return ap.parse_args()


def workload_name(ci):
    return {0: "config0: 10 KF / 2k points, points-only LocalBundleAdjustment",
            1: "config1: 50 KF / 20k points / 50 planes / 10 cuboids",
            2: "config2: 200 KF / 80k points / 200 planes / 50 cuboids",
            3: "config3 window: 50 KF / 20k points / 50 planes / 10 cuboids",
            4: "config4: 1000 KF / 400k points / 1k planes / 200 cuboids"}[ci]


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line).  Rank 0 alone
    samples, all GPUs of the job in one query per half second: eight ranks polling nvidia-smi five times a second
    measurably slow the kernel-launch path of every process."""

    def __init__(self, indices):
        super().__init__(daemon=True)
        self.indices, self.stop_flag, self.rows = list(indices), False, []

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag and self.indices:
            try:
                out = subprocess.run(["nvidia-smi", "-i", ",".join(str(i) for i in self.indices), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                for line in out.splitlines():
                    if line.strip():
                        self.rows.append([x.strip() for x in line.split(",")])
            except Exception:
                pass
            for _ in range(5):
                if self.stop_flag:
                    break
                time.sleep(0.1)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows if len(r) > 2 + i)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def oracle_lm_rate(ppo, g, threads=1, params=None):
    """Times ONE full local BA call of the CPU oracle (tests-only code, used here as the reported CPU baseline).  threads = 1 is
    the reference's configuration (single-threaded g2o); > 1 is the oracle's OpenMP variant, reported separately and labelled."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    o = oracle_lib.Oracle(params) if params is not None else oracle_lib.Oracle()
    if threads > 1:
        o.set_threads(threads)
    o.set_graph(g)
    t0 = time.perf_counter()
    r = o.local_ba()
    dt = time.perf_counter() - t0
    iters = r.round1.iterations + r.round2.iterations
    return iters / dt, iters, dt


def ref_lm_rate(ppo, g):
    """For information: the reference's own g2o code (oracle/_ref, compiled against the Eigen stand-in, so its speed is NOT
    representative of a build with real Eigen) on a small window."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ref_lib
    if not ref_lib.available(build=False):
        return None
    r = ref_lib.Ref()
    r.set_graph(g)
    t0 = time.perf_counter()
    res = r.local_ba()
    dt = time.perf_counter() - t0
    return (res.round1.iterations + res.round2.iterations) / dt


def batched_throughput(ppo, ci, params, device, n_win=8, windows=None):
    """W independent windows of the same size on ONE GPU through ppo_ba_local_ba_batch (one host thread + stream per window,
    configs[3] style): the latency-bound phases of different windows overlap on the device."""
    graphs = windows if windows is not None else [ppo.synth.make_graph(ppo.synth.config(ci, window=100 + w)) for w in range(n_win)]
    n_win = len(graphs)
    engines = [ppo.LocalBA(params, device=device) for _ in range(n_win)]
    for e, g in zip(engines, graphs):
        e.set_graph(g)
    best, iters = None, 0
    for rep in range(4):
        for e in engines:
            e.reset()
        t0 = time.perf_counter()
        res = ppo.local_ba_batch(engines)
        dt = time.perf_counter() - t0
        if rep > 0 and (best is None or dt < best):
            best = dt
        iters = sum(r.round1.iterations + r.round2.iterations for r in res)
    for e in engines:
        e.close()
    return {"windows": n_win, "lm_it_per_s": iters / best, "kf_windows_per_sec": n_win / best, "ms_per_batch": 1e3 * best,
            "note": "ppo_ba_local_ba_batch, wall clock around the whole batch, graphs resident"}


def shim_e2e(ppo, g, reps=3):
    """The reference-facing boundary itself: Optimizer::LocalBACameraPlaneCuboids (csrc/host/ppo_optimizer_shim.cpp) on ONE mock map built
    from the same window and kept across the calls, as LocalMapping calls it: stage A window collection, stage B flattening, engine,
    erase lists and write-back; wall clock of the call.  The first call finds the observation mirror cold (every map point's
    observation map is copied, as the reference does on every call); before each further call the estimates are put back, so the
    same problem (minus the observations the first call erased as outliers) is solved with the mirror warm."""
    A = ppo.abi
    path = os.path.join(A.PKG, "lib", "libppo_shim_mock.so")
    if not os.path.exists(path):
        return None
    L = C.CDLL(path)
    L.ppo_mock_world_create.argtypes = [C.POINTER(A.Graph)]
    L.ppo_mock_world_create.restype = C.c_void_p
    L.ppo_mock_world_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(A.State), C.POINTER(C.c_int32 * 4)]
    L.ppo_mock_world_restore_estimates.argtypes = [C.c_void_p, C.POINTER(A.Graph)]
    L.ppo_mock_world_restore_estimates.restype = None
    L.ppo_mock_world_destroy.argtypes = [C.c_void_p]
    L.ppo_mock_world_destroy.restype = None
    L.ppo_mock_last_call_ms.restype = C.c_double
    L.ppo_shim_last_result.restype = C.POINTER(A.Result)
    L.ppo_shim_mirror_stats.argtypes = [C.POINTER(C.c_longlong * 3)]
    L.ppo_shim_mirror_stats.restype = None
    L.ppo_shim_get_threads.restype = C.c_int
    import numpy as np
    st = A.StateArrays(g.c)
    counts = (C.c_int32 * 4)()
    flag = np.zeros(1, np.uint8)
    # a throw-away map first: the shim's engine handle, its pinned arena and the LM graphs of this window shape are created once per
    # process; the calls measured below only differ in the state of the observation mirror
    W0 = L.ppo_mock_world_create(C.byref(g.c))
    L.ppo_mock_world_run(W0, 1, 0, 0, flag.ctypes.data, C.byref(st.c), C.byref(counts))
    L.ppo_mock_world_destroy(W0)  # (clears the mirror)
    W = L.ppo_mock_world_create(C.byref(g.c))
    cold = best = None
    iters = cold_iters = 0
    stats0, stats1 = (C.c_longlong * 3)(), (C.c_longlong * 3)()
    try:
        for rep in range(reps + 1):
            if rep > 0:
                L.ppo_mock_world_restore_estimates(W, C.byref(g.c))
            L.ppo_shim_mirror_stats(C.byref(stats0))
            if L.ppo_mock_world_run(W, 1, 0, 0, flag.ctypes.data, C.byref(st.c), C.byref(counts)) != 0:
                return None
            L.ppo_shim_mirror_stats(C.byref(stats1))
            ms = float(L.ppo_mock_last_call_ms())
            r = L.ppo_shim_last_result().contents
            it = r.round1.iterations + r.round2.iterations
            dev = r.round1.ms_total + r.round2.ms_total  # device time of the two optimize() calls
            if rep == 0:
                cold, cold_iters, cold_host = ms, it, ms - dev
            elif best is None or ms < best:
                best, iters, best_host = ms, it, ms - dev
                reused, rebuilt = stats1[0] - stats0[0], stats1[1] - stats0[1]
    finally:
        L.ppo_mock_world_destroy(W)
    return {"value": iters / (best * 1e-3), "unit": UNIT, "ms_per_call": best, "lm_iterations": iters, "host_ms_per_call": best_host,
            "first_call": {"value": cold_iters / (cold * 1e-3), "ms_per_call": cold, "lm_iterations": cold_iters, "host_ms_per_call": cold_host,
                           "note": "observation mirror cold: every map point's observation map is copied, as the reference does on every call"},
            "mirror": {"rows_reused": int(reused), "rows_rebuilt": int(rebuilt)},
            "host_threads": int(L.ppo_shim_get_threads()),
            "note": "Optimizer::LocalBACameraPlaneCuboids on a mock map of the same window, map kept across calls: collection + flattening (observation "
                    "rows of unchanged map points from the shim's mirror) + H2D + solve + D2H + write-back, wall clock; host_ms = wall clock minus the "
                    "device time of the two optimize() calls; the later calls solve the window without the observations the first call erased and may "
                    "stop after fewer LM iterations"}


def run_reference(args, rank, world, out_stream):
    """`--impl reference`: the reference's CPU path on the host cores.  The real g2o code is compiled here only against an Eigen
    stand-in (oracle/_ref, slower than a real build), so the number reported is the faster, dependency-free oracle port, single-
    threaded like the reference (G2O_OPENMP off, Thirdparty/g2o/config.h:4), on the SAME full call as the GPU arm.  Rank 0 only."""
    if rank != 0:
        return
    from ppo_pkg import ppo
    g = ppo.synth.make_graph(ppo.synth.config(args.config))
    params = None
    if args.config == 0:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib
        params = oracle_lib.default_params()
        params.solver = ppo.abi.SOLVER_6_3
    # one step = one full local BA call (15 LM iterations on config 2: ~9 s on one core); the run is capped at ~3 minutes
    _, it, dt = oracle_lm_rate(ppo, g, params=params)
    n_steps = max(1, min(args.steps, int(150.0 / max(dt, 1e-3))))
    tot_it, tot_t = it, dt
    for _ in range(n_steps - 1):
        _, it, dt = oracle_lm_rate(ppo, g, params=params)
        tot_it += it
        tot_t += dt
    v = tot_it / tot_t
    sample = f"{n_steps} full local BA call(s) (optimize(5) + outlier pass + optimize(10) = {tot_it // n_steps} LM iterations each) of the same window, 1 host thread"
    extra = {}
    try:
        gs = ppo.synth.make_graph(ppo.synth.config(1, n_kf=12, n_pt=800, n_pl=6, n_cu=3))
        vr = ref_lm_rate(ppo, gs)
        if vr is not None:
            vo, _, _ = oracle_lm_rate(ppo, gs)
            extra = {"small_window_check": {"reference_g2o_on_eigen_stand_in": vr, "oracle_port": vo, "unit": UNIT,
                                            "note": "12 KF / 800 points: the unmodified reference sources (oracle/_ref) next to the port; the stand-in Eigen has no expression templates, so the port is the faster and therefore the conservative baseline"}}
    except Exception as exc:
        extra = {"small_window_check": {"error": str(exc)}}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": n_steps, "warmup": 0,
        "ms_per_step": 1e3 * tot_t / n_steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": workload_name(args.config), "schedule": "optimize(5)+outlier pass+optimize(10)", "sample": sample},
        "cpu_baseline": dict({"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample}, **extra),
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}), file=out_stream, flush=True)


def _protect_stdout():
    """Libraries (NCCL version banner, ...) may print to fd 1; the contract is ONE JSON line on stdout.
    fd 1 is pointed at stderr for the duration of the run and the JSON goes to the saved descriptor."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(saved, "w")


def state_max_err(a, b, sl=None, pl=None):
    import numpy as np
    m = 0.0
    for name in ("kf_pose", "pl_coef", "cu_state"):
        x, y = getattr(a, name), getattr(b, name)
        if name == "pl_coef" and pl is not None:
            y = y[pl[0]:pl[1]]
        if x.size:
            m = max(m, float(np.max(np.abs(x - y) / np.maximum(np.abs(y), 1.0))))
    x, y = a.pt_xyz, (b.pt_xyz if sl is None else b.pt_xyz[sl[0]:sl[1]])
    if x.size:
        m = max(m, float(np.max(np.abs(x - y) / np.maximum(np.abs(y), 1.0))))
    return m


def strong_scaling_config4(ppo, torch, dist, rank, world, local_rank, comm):
    """BASELINE configs[4]: ONE 1000-KF / 400k-point window, points sharded over the ranks; every rank takes part in the collectives.
    Rank 0 also solves the whole window alone (same run, same GPU) for the N = 1 rate and the parity of the sharded solve."""
    g_full = ppo.synth.make_graph(ppo.synth.config(4))
    gs, (p0, p1), _ = ppo.sharding.shard_graph(g_full, rank, world)
    eng = ppo.LocalBA(ppo.default_params(), device=local_rank)
    eng.set_shard(comm, rank, world)
    eng.set_graph(gs)
    eng.local_ba()  # warm-up
    eng.reset()
    torch.cuda.synchronize()
    dist.barrier()
    c0 = eng.collective_count()
    eng.mark(0)
    r = eng.local_ba()
    eng.mark(1)
    ms = eng.elapsed_ms()
    colls = eng.collective_count() - c0
    st = eng.get_state()
    iters = r.round1.iterations + r.round2.iterations
    trials = r.round1.total_trials + r.round2.total_trials
    vals = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    ms = float(vals[0])
    # per-phase CUDA events of one more call (rank 0's view; the phases end at collectives, so the ranks agree to within a trial)
    eng.reset()
    eng.set_profiling(True)
    pr = eng.local_ba()
    eng.set_profiling(False)
    phases = {k: getattr(pr.round1, k) + getattr(pr.round2, k) for k in ("ms_linearize", "ms_schur", "ms_solve", "ms_update", "ms_total")}
    dist_solve = os.environ.get("PPO_DIST_SOLVE", "1") != "0" and world <= 8
    eng.close()
    out = None
    err = torch.zeros(1, dtype=torch.float64, device="cuda")
    single = None
    if rank == 0:
        e1 = ppo.LocalBA(ppo.default_params(), device=local_rank)
        e1.set_graph(g_full)
        e1.local_ba()
        e1.reset()
        e1.mark(0)
        r1 = e1.local_ba()
        e1.mark(1)
        ms1 = e1.elapsed_ms()
        s1 = e1.get_state()
        e1.close()
        it1 = r1.round1.iterations + r1.round2.iterations
        single = (it1 / (ms1 * 1e-3), ms1, s1, it1)
        err[0] = state_max_err(st, s1, (p0, p1), ppo.sharding.plane_range(g_full, rank, world))
    # parity of the other ranks' point slices: rank 0 broadcasts its single-GPU points
    import numpy as np
    n_pt = g_full.c.n_pt
    ref_pts = torch.zeros(n_pt * 3, dtype=torch.float64, device="cuda")
    if rank == 0:
        ref_pts.copy_(torch.from_numpy(np.ascontiguousarray(single[2].pt_xyz).ravel()))
    dist.broadcast(ref_pts, src=0)
    if rank != 0:
        ref = ref_pts.cpu().numpy().reshape(-1, 3)[p0:p1]
        if ref.size:
            err[0] = float(np.max(np.abs(st.pt_xyz - ref) / np.maximum(np.abs(ref), 1.0)))
    dist.all_reduce(err, op=dist.ReduceOp.MAX)
    if rank == 0:
        n_p = r.round1.n_pose_dim
        tiles = (n_p + 63) // 64
        s_bytes = 8 * (tiles * (tiles + 3) // 2) * 64 * 68
        rate = iters / (ms * 1e-3)
        out = {"workload": workload_name(4), "partition": "landmarks partitioned over the ranks: points by edge count, planes by Schur pair count (with their edges); key-frames, cuboids and their edges replicated, owned by rank 0",
               "lm_it_per_s": rate, "ms_per_call": ms, "lm_iterations": iters, "n1_lm_it_per_s": single[0], "n1_ms_per_call": single[1],
               "speedup_vs_n1": rate / single[0], "efficiency": rate / single[0] / world, "shard_parity_max_err": float(err[0]),
               "same_iterations_as_n1": iters == single[3], "collectives_per_call": colls, "damped_trials": trials,
               "n_pose_dim": n_p, "phases_ms_per_call": phases,
               "reduced_system_bytes": s_bytes,
               "collective": ("distributed dense solve over NVLink peer memory (cudaIpc): tile column j of Hschur belongs to rank j mod N; per damped trial every owner "
                              "pulls and sums the other ranks' partial columns (k_dist_reduce, %.0f MB read per rank), the dataflow Cholesky pushes each finished panel tile "
                              "to all ranks with TMA bulk stores + system-scope version counters (k_chol_dist, %.0f MB written per rank), back-substitution replicated; "
                              "NCCL only for the Hpp blocks and three scalars per trial" % (s_bytes * (world - 1) / world / 1e6, s_bytes * (world - 1) / world / 1e6))
                             if dist_solve else "ncclAllReduce(sum, f64) of the packed lower triangle of Hschur + reduced gradient per damped trial; every rank factorises the identical system"}
    return out


def config3_windows(ppo, torch, dist, rank, world, local_rank):
    """BASELINE configs[3]: 64 independent config-1 windows dealt w mod N, all windows of a rank in flight on its GPU
    (ppo_ba_local_ba_batch), no data-path collective."""
    n_total = 64
    mine = ppo.sharding.windows_for_rank(n_total, rank, world)
    graphs = [ppo.synth.make_graph(ppo.synth.config(3, window=w)) for w in mine]
    engines = [ppo.LocalBA(ppo.default_params(), device=local_rank) for _ in mine]
    for e, g in zip(engines, graphs):
        e.set_graph(g)
    ppo.local_ba_batch(engines)  # warm-up
    best = None
    iters = 0
    for _ in range(3):
        for e in engines:
            e.reset()
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        res = ppo.local_ba_batch(engines)
        dt = time.perf_counter() - t0
        v = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        dt = float(v[0])
        best = dt if best is None or dt < best else best
        iters = sum(r.round1.iterations + r.round2.iterations for r in res)
    for e in engines:
        e.close()
    tot = torch.tensor([float(iters)], dtype=torch.float64, device="cuda")
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    if rank != 0:
        return None
    return {"workload": "configs[3]: 64 independent windows of 50 KF / 20k points / 50 planes / 10 cuboids, window w on rank w mod N", "windows": n_total,
            "windows_per_gpu": len(mine), "kf_windows_per_sec": n_total / best, "lm_it_per_s": float(tot[0]) / best, "ms_per_batch": 1e3 * best,
            "collective": "none in the data path (timing reduced with MAX over ranks)"}


def main():
    args = parse()
    out_stream = _protect_stdout()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world, out_stream)
        return
    import numpy as np  # noqa: F401
    import torch
    from ppo_pkg import ppo
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    ci = args.config
    shard = args.shard and world > 1
    # one independent window per rank (own handle, own HBM) unless sharded; every rank gets the SAME seeded window so that
    # the per-GPU work is identical at every N (with different seeds the number of rejected LM trials differs per rank
    # and max-over-ranks timing measures workload variance instead of scaling)
    g = ppo.synth.make_graph(ppo.synth.config(ci, window=0))
    params = ppo.default_params()
    if ci == 0:
        params.solver = ppo.abi.SOLVER_6_3  # points-only LocalBundleAdjustment stack (Optimizer.cc:516-522)
    eng = ppo.LocalBA(params, device=local_rank)
    comm = None
    if world > 1 and (shard or not args.no_extras):
        uid = [ppo.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        comm = ppo.nccl_init(uid[0], rank, world, local_rank)
    if shard:
        eng.set_shard(comm, rank, world)
        g_full = g
        g = ppo.sharding.shard_graph(g_full, rank, world)[0]
    eng.set_graph(g)

    def step_resident():
        eng.reset()
        eng.flush_l2()  # L2 flushed between timed steps (working set of config 2 is ~110 MB < 126 MB L2)
        r = eng.local_ba()
        return r.round1.iterations + r.round2.iterations, r

    def step_e2e():
        eng.set_graph(g)  # host buffers -> HBM inside the timed region
        r = eng.local_ba()
        eng.get_state()   # HBM -> host
        return r.round1.iterations + r.round2.iterations

    for _ in range(max(3, args.warmup)):
        step_resident()
    sampler = ClockSampler(range(world) if rank == 0 else [])
    sampler.start()
    barrier()
    l0, s0 = eng.launch_count(), eng.host_sync_count()
    eng.mark(0)
    iters = 0
    last = None
    for _ in range(args.steps):
        n, last = step_resident()
        iters += n
    eng.mark(1)
    ms = eng.elapsed_ms()
    barrier()
    launches = eng.launch_count() - l0
    host_syncs = eng.host_sync_count() - s0
    sampler.stop_flag = True
    # end-to-end through the C-ABI with host buffers: the step's inputs lie in page-locked host memory (the numpy arrays of the window,
    # registered once), and ppo_ba_set_graph copies page-locked arrays to the device from where they lie
    pinned_inputs = 0
    for v in g.a.values():
        if v.nbytes >= 65536 and eng.lib.ppo_ba_host_register(v.ctypes.data, v.nbytes) == 0:
            pinned_inputs += v.nbytes
    for _ in range(1):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    e2e_iters = 0
    n_e2e = max(1, min(args.steps, 5))
    for _ in range(n_e2e):
        e2e_iters += step_e2e()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    # roofline of the Jacobian / assembly kernel (cold L2 each launch, CUDA events on the engine's stream)
    eng.reset()
    asm_ms, asm_bytes = eng.time_assembly(20)
    eng.reset()
    sol_ms, sol_flops, sol_n = eng.time_solve(10)
    prof = None
    if rank == 0 or shard:  # sharded: every rank takes part in the collectives of the profiled call
        eng.reset()
        eng.set_profiling(True)
        pr = eng.local_ba()
        eng.set_profiling(False)
        prof = {k: getattr(pr.round1, k) + getattr(pr.round2, k) for k in ("ms_linearize", "ms_schur", "ms_solve", "ms_update", "ms_total")}
        prof["note"] = "host-driven loop with per-phase CUDA events (the timed steps run the captured graph, which has no gaps between trials)"

    batched = e2e_shim = None
    if rank == 0 and world == 1 and not args.no_batched:
        batched = batched_throughput(ppo, ci, params, local_rank)
        if ci in (1, 2, 3):
            try:
                e2e_shim = shim_e2e(ppo, g)
            except Exception as exc:  # never lose the bench line over an extra
                e2e_shim = {"error": str(exc)}
    vals = torch.tensor([ms, float(iters), e2e_s, float(e2e_iters)], dtype=torch.float64, device="cuda")
    if dist is not None:
        mx = vals.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = vals.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, e2e_s = float(mx[0]), float(mx[2])
        iters, e2e_iters = float(sm[1]), float(sm[3])
        if shard:  # every rank executes the same LM iterations of the one window
            iters, e2e_iters = iters / world, e2e_iters / world
    eng.close()
    config3 = strong = None
    if world > 1 and not shard and not args.no_extras and ci == 2:
        try:
            config3 = config3_windows(ppo, torch, dist, rank, world, local_rank)
        except Exception as exc:
            config3 = {"error": str(exc)}
        try:
            strong = strong_scaling_config4(ppo, torch, dist, rank, world, local_rank, comm)
        except Exception as exc:
            strong = {"error": str(exc)}
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        traffic = {}
        try:  # dram__bytes_read+write per launch from the committed ncu --set full capture, per workload
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json"))).get("config%d" % ci, {})
        except Exception:
            pass
        achieved = asm_bytes / (asm_ms * 1e-3) / 1e9
        value = iters / (ms * 1e-3)
        state_bytes = 8 * (7 * g.c.n_kf + 3 * g.c.n_pt + 4 * g.c.n_pl + 10 * g.c.n_cu)
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms / max(1, args.steps), "higher_is_better": True, "scaling": "strong" if shard else "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_name(ci), "windows_per_gpu": 1, "schedule": "optimize(5)+outlier pass+optimize(10)",
                       "partition": "one window, points sharded over ranks, ncclAllReduce of Hschur|bschur per damped trial" if shard else "independent windows (one per rank, same seeded content), no data-path collective",
                       "l2": "flushed between timed steps (256 MiB memset)", "lm_iterations_per_step": iters / max(1, args.steps) / (1 if shard else world),
                       "n_pose_dim": last.round1.n_pose_dim, "n_point_edges": g.c.n_pe,
                       "lm_control": "on the device: one CUDA graph with conditional WHILE nodes per optimize()"},
            "kf_windows_per_sec": (1 if shard else world) * args.steps / (ms * 1e-3),
            "e2e": {"value": e2e_iters / e2e_s, "unit": UNIT, "h2d_bytes_per_step": g.nbytes(), "d2h_bytes_per_step": state_bytes,
                    "ms_per_step": 1e3 * e2e_s / n_e2e, "host_inputs": "page-locked (cudaHostRegister), %d of %d bytes; copied to the device from where they lie" % (pinned_inputs, g.nbytes())},
            "e2e_shim": e2e_shim,
            "gpu_launches": launches,
            "host_syncs_per_step": host_syncs / max(1, args.steps),
            # the time-dominant phase of a step is the dense solve of the reduced pose system: FP64 tensor
            # pipe (DMMA) for the tile products, latency-bound critical path through the diagonal tiles.  Peak: FP64 DMMA/DFMA rate
            # measured with tools/ubench/fp64.cu (64 FMA/clk/SM x 148 SMs x 1.965 GHz = 37.2 TFLOP/s); MEASURED_PEAKS.json has no FP64 figure.
            "roofline": {"kernel": "dense Hschur solve: k_chol_dataflow (persistent tiled Cholesky: TMA-staged 64x64 tiles, DMMA, version-counter dataflow) + k_backsolve_chain", "bound": "tensor",
                         "achieved": sol_flops / (sol_ms * 1e-3) / 1e12, "peak": FP64_TFLOPS, "unit": "TFLOP/s",
                         "frac": sol_flops / (sol_ms * 1e-3) / 1e12 / FP64_TFLOPS, "traffic": traffic.get("k_chol_dataflow"), "ms": sol_ms, "algo_flops": sol_flops,
                         "n_p": sol_n, "peak_source": "measured FP64 DMMA rate, tools/ubench/fp64.cu (of measured; bf16 peak in MEASURED_PEAKS.json does not apply to an FP64 solve)"},
            # the kernel the north_star names: Jacobian / assembly over the point edges, HBM-bound
            "roofline_assembly": {"kernel": "k_point_linearize (point-edge Jacobian/assembly)", "bound": "hbm", "achieved": achieved, "peak": peak,
                                  "unit": "GB/s", "frac": achieved / peak, "traffic": traffic.get("k_point_linearize"), "ms": asm_ms,
                                  "algo_bytes": asm_bytes,
                                  "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"},
            "phases_ms_per_call": prof,
            "batched": batched,
            "config3": config3,
            "strong": strong,
            "clocks": sampler.summary(),
        }
        if not args.no_cpu_baseline and not shard and ci != 4 and world == 1:
            v, it, dt = oracle_lm_rate(ppo, g, params=None if ci else params)
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                                   "sample": f"one full local BA call ({it} LM iterations, {dt:.1f} s) of the same window on 1 host thread "
                                             f"(the reference runs g2o single-threaded); host has {os.cpu_count()} cores"}
            try:  # also the oracle's OpenMP variant on all host cores -- NOT the reference's configuration
                nthr = os.cpu_count() or 1
                os.environ.pop("OMP_NUM_THREADS", None)
                v2, it2, dt2 = oracle_lm_rate(ppo, g, threads=nthr, params=None if ci else params)
                out["cpu_baseline_mt"] = {"value": v2, "unit": UNIT, "cores": nthr, "kind": "port, OpenMP variant (not the reference configuration)",
                                          "sample": f"one full local BA call ({it2} LM iterations, {dt2:.1f} s) of the same window"}
            except Exception as exc:  # never lose the bench line over the extra baseline
                out["cpu_baseline_mt"] = {"error": str(exc)}
        print(json.dumps(out), file=out_stream, flush=True)
    if dist is not None:
        dist.barrier()
        if comm is not None:
            ppo.nccl_destroy(comm)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
