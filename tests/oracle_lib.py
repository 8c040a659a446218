"""tests-only binding of the CPU oracle (oracle/ppo_oracle.cpp). Never imported by the product."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ppo_pkg import ppo  # noqa: E402

A = ppo.abi
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(A.ROOT, "oracle", "_build", "libppo_oracle.so")
        if not os.path.exists(path):
            import importlib
            importlib.import_module("ppo_slam_b200._build").build_oracle()
        L = C.CDLL(path)
        ppo.engine._bind(L, "ppo_oracle_")
        L.ppo_oracle_create.argtypes = [C.POINTER(A.Params), C.POINTER(C.c_void_p)]
        L.ppo_oracle_debug_linearize.argtypes = [C.c_void_p, C.POINTER(C.c_int32 * 2), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ppo_oracle_set_threads.argtypes = [C.c_void_p, C.c_int]
        L.ppo_oracle_debug_solve.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]
        _LIB = L
    return _LIB


def default_params():
    return ppo.engine.default_params(lib(), "ppo_oracle_")


class Oracle(ppo.Handle):
    def __init__(self, params=None):
        L = lib()
        p = params or default_params()
        h = C.c_void_p()
        assert L.ppo_oracle_create(C.byref(p), C.byref(h)) == 0
        super().__init__(L, "ppo_oracle_", h)
        self.params = p

    def set_threads(self, n):
        """Host threads of the oracle (1 = the reference's single-threaded g2o configuration, the default)."""
        assert self.lib.ppo_oracle_set_threads(self.h, int(n)) == 0
