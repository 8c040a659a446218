#!/usr/bin/env python
"""bench.py — LM iterations/sec (and KF-windows/sec) of the local bundle adjustment on synthetic windows.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config 2] [--impl ours|reference]

A "step" is one complete local BA call (optimize(5) -> outlier pass -> optimize(10), the reference
schedule of Optimizer.cc:2727-2837) on one synthetic window per GPU.  N=1 workload: BASELINE.json
configs[2] (200 KF / 80k points / 200 planes / 50 cuboids), the window the north_star target
(>= 50 LM it/s on 1 GPU) is quoted on.  N>1: every rank solves its own window of the same size
(independent key-frame windows, no data-path collective) -> weak scaling.
`value`  : device-resident — graph already in HBM, timed with CUDA events on the engine's stream.
`e2e`    : through the C-ABI with host buffers: ppo_ba_set_graph (H2D) + ppo_ba_local_ba + ppo_ba_get_state (D2H).
`--impl reference`: the CPU oracle (single thread, the reference's own configuration: g2o OpenMP off).
"""
import argparse
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


def oracle_lm_rate(ppo, g, full_call, threads=1):
    """Times the CPU oracle (tests-only code, used here as the reported CPU baseline).  threads = 1 is the reference's
    configuration (single-threaded g2o); > 1 is the oracle's OpenMP variant, reported separately and labelled as such."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    o = oracle_lib.Oracle()
    if threads > 1:
        o.set_threads(threads)
    o.set_graph(g)
    t0 = time.perf_counter()
    if full_call:
        r = o.local_ba()
        iters = r.round1.iterations + r.round2.iterations
    else:
        iters = o.optimize(o.params.iters_round1).iterations
    dt = time.perf_counter() - t0
    return iters / dt, iters, dt


def batched_throughput(ppo, ci, params, device, n_win=8, n_thr=8):
    """Extra: W independent windows of the same size on ONE GPU, T host threads / handles / streams (configs[3] style):
    the latency-bound phases of different windows overlap on the device."""
    graphs = [ppo.synth.make_graph(ppo.synth.config(ci, window=100 + w)) for w in range(n_win)]
    engines = [ppo.LocalBA(params, device=device) for _ in range(n_win)]
    for e, g in zip(engines, graphs):
        e.set_graph(g)
    iters = [0] * n_win

    def run_all():
        nxt, lock = [0], threading.Lock()

        def worker():
            while True:
                with lock:
                    w = nxt[0]
                    nxt[0] += 1
                if w >= n_win:
                    return
                engines[w].reset()
                r = engines[w].local_ba()
                iters[w] = r.round1.iterations + r.round2.iterations

        ts = [threading.Thread(target=worker) for _ in range(n_thr)]
        t0 = time.perf_counter()
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        return time.perf_counter() - t0

    run_all()
    dt = min(run_all() for _ in range(3))
    for e in engines:
        e.close()
    return {"windows": n_win, "host_threads": n_thr, "lm_it_per_s": sum(iters) / dt, "kf_windows_per_sec": n_win / dt,
            "note": "wall clock around the whole batch, graphs resident"}


def run_reference(args, rank, world, out_stream):
    """`--impl reference`: the reference's CPU path. It cannot be compiled here (Eigen3 missing),
    so the oracle port is timed, single-threaded like the reference (G2O_OPENMP off). Rank 0 only."""
    if rank != 0:
        return
    from ppo_pkg import ppo
    g = ppo.synth.make_graph(ppo.synth.config(args.config))
    for _ in range(min(args.warmup, 1)):
        oracle_lm_rate(ppo, g, False)
    tot_it, tot_t = 0, 0.0
    for _ in range(max(1, args.steps)):
        _, it, dt = oracle_lm_rate(ppo, g, False)
        tot_it += it
        tot_t += dt
    v = tot_it / tot_t
    sample = f"round 1 only (optimize({5}) = {tot_it // max(1, args.steps)} LM iterations) of the same window per step"
    mt = None
    try:  # for information: the oracle's OpenMP variant on all host cores (NOT the reference's configuration), one sample
        nthr = os.cpu_count() or 1
        v2, it2, dt2 = oracle_lm_rate(ppo, g, False, threads=nthr)
        mt = {"value": v2, "unit": UNIT, "cores": nthr, "kind": "port, OpenMP variant (not the reference configuration)", "sample": sample}
    except Exception as exc:
        mt = {"error": str(exc)}
    print(file=out_stream, flush=True, *[json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot_t / max(1, args.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": workload_name(args.config), "sample": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample}, "cpu_baseline_mt": mt,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0})])


def _protect_stdout():
    """Libraries (NCCL version banner, ...) may print to fd 1; the contract is ONE JSON line on stdout.
    fd 1 is pointed at stderr for the duration of the run and the JSON goes to the saved descriptor."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(saved, "w")


def main():
    args = parse()
    out_stream = _protect_stdout()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world, out_stream)
        return
    import numpy as np
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
    if shard:
        uid = [ppo.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        comm = ppo.nccl_init(uid[0], rank, world, local_rank)
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
    l0 = eng.launch_count()
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
    sampler.stop_flag = True
    # end-to-end through the C-ABI with host buffers
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

    batched = None
    if rank == 0 and world == 1 and not args.no_batched:
        batched = batched_throughput(ppo, ci, params, local_rank)
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
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        traffic = {}
        try:  # dram__bytes_read+write per launch from the committed ncu --set full capture
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
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
                       "l2": "flushed between timed steps (256 MiB memset)", "lm_iterations_per_step": iters / max(1, args.steps) / world,
                       "n_pose_dim": last.round1.n_pose_dim, "n_point_edges": g.c.n_pe},
            "kf_windows_per_sec": (1 if shard else world) * args.steps / (ms * 1e-3),
            "e2e": {"value": e2e_iters / e2e_s, "unit": UNIT, "h2d_bytes_per_step": g.nbytes(), "d2h_bytes_per_step": state_bytes,
                    "ms_per_step": 1e3 * e2e_s / n_e2e},
            "gpu_launches": launches,
            # the time-dominant phase of a step is the dense solve of the reduced pose system: FP64 tensor
            # pipe (DMMA) for the trailing updates, latency-bound panel chain.  Peak: FP64 DMMA/DFMA rate measured with
            # tools/ubench/fp64.cu (64 FMA/clk/SM x 148 SMs x 1.965 GHz = 37.2 TFLOP/s); MEASURED_PEAKS.json has no FP64 figure.
            "roofline": {"kernel": "dense Hschur solve: k_panel_gemm + k_syrk_update (DMMA trailing update, CTA 0 factorises the next diagonal block) + k_backsolve_step", "bound": "tensor",
                         "achieved": sol_flops / (sol_ms * 1e-3) / 1e12, "peak": FP64_TFLOPS, "unit": "TFLOP/s",
                         "frac": sol_flops / (sol_ms * 1e-3) / 1e12 / FP64_TFLOPS, "traffic": None, "ms": sol_ms, "algo_flops": sol_flops,
                         "n_p": sol_n, "peak_source": "measured FP64 DMMA rate, tools/ubench/fp64.cu (of measured; bf16 peak in MEASURED_PEAKS.json does not apply to an FP64 solve)"},
            # the kernel the north_star names: Jacobian / assembly over the point edges, HBM-bound
            "roofline_assembly": {"kernel": "k_point_linearize (point-edge Jacobian/assembly)", "bound": "hbm", "achieved": achieved, "peak": peak,
                                  "unit": "GB/s", "frac": achieved / peak, "traffic": traffic.get("k_point_linearize"), "ms": asm_ms,
                                  "algo_bytes": asm_bytes,
                                  "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"},
            "phases_ms_per_call": prof,
            "batched": batched,
            "clocks": sampler.summary(),
        }
        if not args.no_cpu_baseline and not shard and ci != 4:
            g0 = g if world == 1 else ppo.synth.make_graph(ppo.synth.config(ci))
            v, it, dt = oracle_lm_rate(ppo, g0, True)
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                                   "sample": f"one full local BA call ({it} LM iterations, {dt:.1f} s) of the same window on 1 host thread "
                                             f"(the reference runs g2o single-threaded); host has {os.cpu_count()} cores"}
            try:  # also the oracle's OpenMP variant on all host cores -- NOT the reference's configuration
                nthr = os.cpu_count() or 1
                v2, it2, dt2 = oracle_lm_rate(ppo, g0, True, threads=nthr)
                out["cpu_baseline_mt"] = {"value": v2, "unit": UNIT, "cores": nthr, "kind": "port, OpenMP variant (not the reference configuration)",
                                          "sample": f"one full local BA call ({it2} LM iterations, {dt2:.1f} s) of the same window"}
            except Exception as exc:  # never lose the bench line over the extra baseline
                out["cpu_baseline_mt"] = {"error": str(exc)}
        print(json.dumps(out), file=out_stream, flush=True)
    if dist is not None:
        dist.barrier()
        if comm is not None:
            eng.close()
            ppo.nccl_destroy(comm)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
