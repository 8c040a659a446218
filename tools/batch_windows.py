"""Throughput of a batch of independent windows on ONE GPU (BASELINE configs[3] style): T host threads, each driving its own
handle / CUDA stream, so the latency-bound phases of different windows overlap on the device.
  python tools/batch_windows.py <config> <n_windows> <threads> [rounds]"""
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ppo_pkg import ppo  # noqa: E402

ci, n_win, n_thr = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
rounds = int(sys.argv[4]) if len(sys.argv) > 4 else 3
graphs = [ppo.synth.make_graph(ppo.synth.config(ci, window=w)) for w in range(n_win)]
engines = [ppo.LocalBA() for _ in range(n_win)]
for e, g in zip(engines, graphs):
    e.set_graph(g)


def run_all(n_threads):
    iters = [0] * n_win
    nxt = [0]
    lock = threading.Lock()

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

    ts = [threading.Thread(target=worker) for _ in range(n_threads)]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return time.perf_counter() - t0, sum(iters)


run_all(n_thr)
for thr in sorted({1, 2, 4, n_thr}):
    best = min(run_all(thr) for _ in range(rounds))
    print(f"config {ci}: {n_win} windows, {thr:2d} threads: {best[0] * 1e3:8.1f} ms  -> {n_win / best[0]:7.1f} windows/s, {best[1] / best[0]:8.1f} LM it/s")
