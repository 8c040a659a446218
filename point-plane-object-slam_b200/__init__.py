"""B200-native local bundle adjustment for point / plane / cuboid SLAM graphs.

Drop-in engine for ONE hot path of benchun123/point-plane-object-SLAM:
Optimizer::LocalBACameraPlaneCuboids and Optimizer::LocalBundleAdjustment.
See DESIGN.md and include/ppo_ba.h.
"""
from . import _abi as abi  # noqa: F401
from . import sharding  # noqa: F401
from . import synth  # noqa: F401
from .engine import EngineError, Handle, LocalBA, default_params, load_library, local_ba_batch, nccl_destroy, nccl_init, nccl_unique_id  # noqa: F401
