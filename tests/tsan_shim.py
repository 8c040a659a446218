"""ThreadSanitizer pass over the threaded host paths of the shim (stages A, B, G on the worker pool), oracle-backed, no GPU:
  B=oracle/_build; H=point-plane-object-slam_b200/csrc/host
  g++ -O1 -g -std=c++17 -pthread -fsanitize=thread -shared -fPIC -DPPO_SHIM_ON_ORACLE -I include -o /tmp/libppo_shim_tsan.so \
      $H/ppo_optimizer_shim.cpp $H/ppo_mock_world.cpp -L $B -lppo_oracle -Wl,-rpath,$PWD/$B
  TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0" LD_PRELOAD=$(gcc -print-file-name=libtsan.so) python tests/tsan_shim.py
Expected: no "WARNING: ThreadSanitizer" line (round 2: none over threaded flattening with 4 and 8 threads, a full call with
write-back, and two warm calls on a persistent map)."""
import os, sys, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, ROOT)  # (test infrastructure: links the oracle-backed test build of the shim)
import shim_lib
from ppo_pkg import ppo
A = ppo.abi
L = C.CDLL('/tmp/libppo_shim_tsan.so')
shim_lib._declare(L)
L.ppo_shim_last_graph.restype = C.POINTER(A.Graph)
L.ppo_shim_last_rc.restype = C.c_int
g = ppo.synth.make_graph(ppo.synth.config(1, n_kf=24, n_fixed=5, n_pt=12000, n_pl=6, n_cu=3))
for n in (4, 8):
    L.ppo_shim_set_threads(n)
    st, counts, flat = shim_lib.run(g, mixed=True, stop=True, backend=L)
    print("threads", n, "flatten ok", flat.c.n_pt, flat.c.n_pe, flush=True)
g2 = ppo.synth.make_graph(ppo.synth.config(1, n_kf=10, n_fixed=2, n_pt=5000, n_pl=4, n_cu=2))
L.ppo_shim_set_threads(4)
st, counts, flat = shim_lib.run(g2, backend=L)
print("full call ok", counts, flush=True)
# warm path on a persistent world
W = shim_lib.World(g2, backend=L)
for i in range(2):
    W.run()
print("world ok", flush=True)
