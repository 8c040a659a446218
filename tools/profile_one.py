"""One local BA call on a BASELINE config (no warm-up): the command profiled with ncu (launch list / --set full)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ppo_pkg import ppo  # noqa: E402

ci = int(sys.argv[1]) if len(sys.argv) > 1 else 2
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 1
g = ppo.synth.make_graph(ppo.synth.config(ci))
e = ppo.LocalBA()
e.set_graph(g)
for _ in range(rounds):
    e.reset()
    r = e.local_ba()
print("iters", r.round1.iterations + r.round2.iterations, "launches", e.launch_count(), "device ms", r.round1.ms_total + r.round2.ms_total)
