"""Device time of the point-edge Jacobian/assembly kernel on a BASELINE config (CUDA events inside ppo_ba_time_assembly)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ppo_pkg import ppo  # noqa: E402

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
g = ppo.synth.make_graph(ppo.synth.config(cfg))
eng = ppo.LocalBA(device=0)
eng.set_graph(g)
eng.optimize(1)
print("k_point_linearize ms:", [round(eng.time_assembly(20)[0], 5) for _ in range(3)], "PPO_LIN_MINB =", os.environ.get("PPO_LIN_MINB"))
