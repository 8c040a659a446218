"""ctypes binding of the Optimizer shim + mock-map harness (lib/libppo_shim_mock.so)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ppo_pkg import ppo  # noqa: E402

A = ppo.abi
_LIB = None


_ORACLE_LIB = None


def _declare(L):
    L.ppo_mock_set_options.argtypes = [C.c_int, C.c_int]
    L.ppo_mock_set_options.restype = None
    L.ppo_mock_run.argtypes = [C.POINTER(A.Graph), C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(A.State), C.POINTER(C.c_int32 * 4)]
    L.ppo_mock_run_global.argtypes = [C.POINTER(A.Graph), C.c_int, C.c_ulong, C.c_int, C.c_void_p, C.POINTER(A.State), C.POINTER(C.c_int32 * 4)]
    L.ppo_shim_last_result.restype = C.POINTER(A.Result)
    L.ppo_mock_run_pose.argtypes = [C.POINTER(A.Graph), C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32 * 3)]
    L.ppo_mock_world_create.argtypes = [C.POINTER(A.Graph)]
    L.ppo_mock_world_create.restype = C.c_void_p
    L.ppo_mock_world_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(A.State), C.POINTER(C.c_int32 * 4)]
    L.ppo_mock_world_erase_observation.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.ppo_mock_world_perturb_point.argtypes = [C.c_void_p, C.c_int, C.c_float]
    L.ppo_mock_world_destroy.argtypes = [C.c_void_p]
    L.ppo_mock_world_destroy.restype = None
    L.ppo_shim_mirror_clear.restype = None
    L.ppo_shim_mirror_stats.argtypes = [C.POINTER(C.c_longlong * 3)]
    L.ppo_shim_mirror_stats.restype = None
    L.ppo_mock_last_call_ms.restype = C.c_double
    L.ppo_shim_set_threads.argtypes = [C.c_int]
    L.ppo_shim_set_threads.restype = None


def oracle_backed_lib():
    """TEST BUILD: the shim source compiled with -DPPO_SHIM_ON_ORACLE and linked to the CPU oracle, so that the shim's host
    logic runs end to end without a GPU.  Lives under oracle/_build (test infrastructure), never in the product lib/."""
    global _ORACLE_LIB
    if _ORACLE_LIB is None:
        import subprocess
        import oracle_lib
        oracle_lib.lib()
        bdir = os.path.join(A.ROOT, "oracle", "_build")
        out = os.path.join(bdir, "libppo_shim_mock_oracle.so")
        host = os.path.join(A.PKG, "csrc", "host")
        srcs = [os.path.join(host, "ppo_optimizer_shim.cpp"), os.path.join(host, "ppo_mock_world.cpp")]
        deps = srcs + [os.path.join(host, f) for f in ("ppo_mock_slam.h", "ppo_convert.h")] + [os.path.join(bdir, "libppo_oracle.so")]
        if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-shared", "-fPIC", "-DPPO_SHIM_ON_ORACLE", "-o", out] + srcs +
                                  ["-L", bdir, "-lppo_oracle", "-Wl,-rpath,$ORIGIN"])
        L = C.CDLL(out)
        _declare(L)
        L.ppo_shim_last_graph.restype = C.POINTER(A.Graph)
        L.ppo_shim_last_rc.restype = C.c_int
        _ORACLE_LIB = L
    return _ORACLE_LIB


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(A.PKG, "lib", "libppo_shim_mock.so")
        L = C.CDLL(path)
        _declare(L)
        L.ppo_mock_run.argtypes = [C.POINTER(A.Graph), C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(A.State), C.POINTER(C.c_int32 * 4)]
        L.ppo_mock_run_global.argtypes = [C.POINTER(A.Graph), C.c_int, C.c_ulong, C.c_int, C.c_void_p, C.POINTER(A.State), C.POINTER(C.c_int32 * 4)]
        L.ppo_shim_last_graph.restype = C.POINTER(A.Graph)
        L.ppo_shim_last_result.restype = C.POINTER(A.Result)
        L.ppo_shim_last_rc.restype = C.c_int
        _LIB = L
    return _LIB


def run(g, mixed=True, fix_camera=False, fix_point=False, stop=False, backend=None):
    """LocalMapping-style call on a mock map built from the flat graph g. Returns (state, counts, flattened graph)."""
    L = backend or lib()
    st = A.StateArrays(g.c)
    counts = (C.c_int32 * 4)()
    flag = np.array([1 if stop else 0], np.uint8)
    rc = L.ppo_mock_run(C.byref(g.c), int(mixed), int(fix_camera), int(fix_point), flag.ctypes.data, C.byref(st.c), C.byref(counts))
    assert rc == 0
    flat = A.GraphArrays.from_c(L.ppo_shim_last_graph().contents)
    return st, list(counts), flat


def run_global(g, n_iterations=10, n_loop_kf=0, robust=True, stop=False, backend=None):
    """LoopClosing / Tracking-style Optimizer::GlobalBundleAdjustemnt call on a mock map made of g's key-frames and points."""
    L = backend or lib()
    st = A.StateArrays(g.c)
    counts = (C.c_int32 * 4)()
    flag = np.array([1 if stop else 0], np.uint8)
    rc = L.ppo_mock_run_global(C.byref(g.c), int(n_iterations), int(n_loop_kf), int(robust), flag.ctypes.data, C.byref(st.c), C.byref(counts))
    assert rc == 0
    flat = A.GraphArrays.from_c(L.ppo_shim_last_graph().contents)
    return st, list(counts), flat, L.ppo_shim_last_rc()


def run_pose(g, kf, backend=None):
    """Tracking-style Optimizer::PoseOptimization on a mock Frame made of key-frame slot kf of g.
    Returns (pose7, outlier flags per associated feature, [return value, SetPose calls, associated features], flat graph, rc)."""
    L = backend or lib()
    pose = np.zeros(7)
    rp = g["pt_rowptr"]
    n = int((g["pe_kf"] == kf).sum())
    out = np.zeros(max(n, 1), np.uint8)
    counts = (C.c_int32 * 3)()
    rc = L.ppo_mock_run_pose(C.byref(g.c), int(kf), pose.ctypes.data, out.ctypes.data, C.byref(counts))
    assert rc == 0
    flat = A.GraphArrays.from_c(L.ppo_shim_last_graph().contents)
    return pose, out[:n], list(counts), flat, L.ppo_shim_last_rc()


class World:
    """A mock map that outlives one call: consecutive Optimizer::LocalBACameraPlaneCuboids calls on the same map (ppo_mock_world_*)."""

    def __init__(self, g, backend=None):
        self.L = backend or lib()
        self.g = g
        self.h = self.L.ppo_mock_world_create(C.byref(g.c))

    def run(self, mixed=True, fix_camera=False, fix_point=False, stop=False):
        st = A.StateArrays(self.g.c)
        counts = (C.c_int32 * 4)()
        flag = np.array([1 if stop else 0], np.uint8)
        rc = self.L.ppo_mock_world_run(self.h, int(mixed), int(fix_camera), int(fix_point), flag.ctypes.data, C.byref(st.c), C.byref(counts))
        assert rc == 0
        flat = A.GraphArrays.from_c(self.L.ppo_shim_last_graph().contents)
        return st, list(counts), flat

    def erase_observation(self, point, kf):
        return self.L.ppo_mock_world_erase_observation(self.h, int(point), int(kf))

    def perturb_point(self, point, dx):
        self.L.ppo_mock_world_perturb_point(self.h, int(point), float(dx))

    def mirror_stats(self):
        out = (C.c_longlong * 3)()
        self.L.ppo_shim_mirror_stats(C.byref(out))
        return {"reused": out[0], "rebuilt": out[1], "pool": out[2]}

    def mirror_clear(self):
        self.L.ppo_shim_mirror_clear()

    def last_call_ms(self):
        return float(self.L.ppo_mock_last_call_ms())

    def close(self):
        if self.h:
            self.L.ppo_mock_world_destroy(self.h)
            self.h = None
