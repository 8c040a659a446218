"""ctypes binding of the Optimizer shim + mock-map harness (lib/libppo_shim_mock.so)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ppo_pkg import ppo  # noqa: E402

A = ppo.abi
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(A.PKG, "lib", "libppo_shim_mock.so")
        L = C.CDLL(path)
        L.ppo_mock_run.argtypes = [C.POINTER(A.Graph), C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(A.State), C.POINTER(C.c_int32 * 4)]
        L.ppo_mock_run_global.argtypes = [C.POINTER(A.Graph), C.c_int, C.c_ulong, C.c_int, C.c_void_p, C.POINTER(A.State), C.POINTER(C.c_int32 * 4)]
        L.ppo_shim_last_graph.restype = C.POINTER(A.Graph)
        L.ppo_shim_last_result.restype = C.POINTER(A.Result)
        L.ppo_shim_last_rc.restype = C.c_int
        _LIB = L
    return _LIB


def run(g, mixed=True, fix_camera=False, fix_point=False, stop=False):
    """LocalMapping-style call on a mock map built from the flat graph g. Returns (state, counts, flattened graph)."""
    L = lib()
    st = A.StateArrays(g.c)
    counts = (C.c_int32 * 4)()
    flag = np.array([1 if stop else 0], np.uint8)
    rc = L.ppo_mock_run(C.byref(g.c), int(mixed), int(fix_camera), int(fix_point), flag.ctypes.data, C.byref(st.c), C.byref(counts))
    assert rc == 0
    flat = A.GraphArrays.from_c(L.ppo_shim_last_graph().contents)
    return st, list(counts), flat


def run_global(g, n_iterations=10, n_loop_kf=0, robust=True, stop=False):
    """LoopClosing / Tracking-style Optimizer::GlobalBundleAdjustemnt call on a mock map made of g's key-frames and points."""
    L = lib()
    st = A.StateArrays(g.c)
    counts = (C.c_int32 * 4)()
    flag = np.array([1 if stop else 0], np.uint8)
    rc = L.ppo_mock_run_global(C.byref(g.c), int(n_iterations), int(n_loop_kf), int(robust), flag.ctypes.data, C.byref(st.c), C.byref(counts))
    assert rc == 0
    flat = A.GraphArrays.from_c(L.ppo_shim_last_graph().contents)
    return st, list(counts), flat, L.ppo_shim_last_rc()
