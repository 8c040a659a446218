"""e2e_shim alone (bench.shim_e2e): Optimizer::LocalBACameraPlaneCuboids on a config-sized mock map, cold and warm mirror, for the given
host thread counts of the flattening loops.  PPO_BA_TIMING=1 adds the phase times of every call on stderr.
Usage: python tools/shim_probe.py [config] [threads ...]"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from ppo_pkg import ppo  # noqa: E402

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
threads = [int(a) for a in sys.argv[2:]] or [0]
g = ppo.synth.make_graph(ppo.synth.config(cfg))
L = C.CDLL(os.path.join(ppo.abi.PKG, "lib", "libppo_shim_mock.so"))
L.ppo_shim_set_threads.argtypes = [C.c_int]
for t in threads:
    if t > 0:
        L.ppo_shim_set_threads(t)
    r = bench.shim_e2e(ppo, g)
    print(json.dumps({"threads": L.ppo_shim_get_threads(), "warm_ms": r["ms_per_call"], "warm_host_ms": r["host_ms_per_call"],
                      "cold_ms": r["first_call"]["ms_per_call"], "cold_host_ms": r["first_call"]["host_ms_per_call"], "lm_it_per_s": r["value"]}), flush=True)
