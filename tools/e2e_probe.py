"""Host-side phase times of one end-to-end C-ABI step (set_graph + local_ba + get_state); PPO_BA_TIMING=1 adds the
set_graph breakdown on stderr.  Usage: python tools/e2e_probe.py [config] [reps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ppo_pkg import ppo  # noqa: E402

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
g = ppo.synth.make_graph(ppo.synth.config(cfg))
eng = ppo.LocalBA(device=0)
for r in range(reps):
    t0 = time.perf_counter()
    eng.set_graph(g)
    t1 = time.perf_counter()
    eng.local_ba()
    t2 = time.perf_counter()
    eng.get_state()
    t3 = time.perf_counter()
    print(f"rep {r}: set_graph {1e3 * (t1 - t0):.2f} ms  local_ba {1e3 * (t2 - t1):.2f} ms  get_state {1e3 * (t3 - t2):.2f} ms", flush=True)
