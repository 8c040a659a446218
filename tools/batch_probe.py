"""ppo_ba_local_ba_batch on ONE GPU: n windows of a config, wall clock of the whole batch (the `batched` object of bench.py).
  python tools/batch_probe.py <config> <n_windows>"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ppo_pkg import ppo  # noqa: E402

ci, n = int(sys.argv[1]), int(sys.argv[2])
print(os.environ.get("PPO_BATCH_SM_OVERSUB", "1.0"), bench.batched_throughput(ppo, ci, ppo.default_params(), 0, n_win=n))
