"""Import shim: the package directory is named `point-plane-object-slam_b200` (not a Python
identifier), so it is loaded here under the module name `ppo_slam_b200`."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "point-plane-object-slam_b200")


def load():
    if "ppo_slam_b200" in sys.modules:
        return sys.modules["ppo_slam_b200"]
    spec = importlib.util.spec_from_file_location(
        "ppo_slam_b200", os.path.join(_DIR, "__init__.py"), submodule_search_locations=[_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ppo_slam_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


ppo = load()
