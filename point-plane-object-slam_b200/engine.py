"""Host-side Python mirror of the C-ABI (include/ppo_ba.h).

`LocalBA` wraps a ppo_ba handle with the call sequence of the reference's
Optimizer::LocalBACameraPlaneCuboids / LocalBundleAdjustment (src/Optimizer.cc:1994-2967, :461-786):
set_graph -> optimize(5) -> outlier pass -> optimize(10) -> get_state.  PyTorch is not needed for
this path; all compute happens inside lib/libppo_ba.so (CUDA, sm_100a).  There is NO CPU fallback:
if the library is missing or no GPU is visible the constructor raises.
"""
import ctypes as C
import os

import numpy as np

from . import _abi as A

_LIB = None


class EngineError(RuntimeError):
    pass


def _bind(lib, prefix):
    """Declares argtypes for the ppo_ba-shaped API with the given symbol prefix."""
    H = C.c_void_p
    f = lambda n: getattr(lib, prefix + n)
    f("set_graph").argtypes = [H, C.POINTER(A.Graph)]
    f("optimize").argtypes = [H, C.c_int, C.c_void_p, C.POINTER(A.Stats)]
    f("edge_chi2").argtypes = [H, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    f("set_edge_flags").argtypes = [H, C.c_int, C.c_void_p]
    f("get_edge_flags").argtypes = [H, C.c_int, C.c_void_p]
    f("edge_count").argtypes = [H, C.c_int]
    f("recompute_edge_errors").argtypes = [H, C.c_int]
    f("outlier_pass").argtypes = [H, C.POINTER(C.c_int32 * 3)]
    f("local_ba").argtypes = [H, C.c_void_p, C.POINTER(A.Result)]
    f("get_state").argtypes = [H, C.POINTER(A.State)]
    f("reset").argtypes = [H]
    f("destroy").argtypes = [H]
    f("destroy").restype = None
    f("default_params").argtypes = [C.POINTER(A.Params)]
    f("default_params").restype = None
    return lib


def load_library():
    global _LIB
    if _LIB is None:
        path = os.path.join(A.PKG, "lib", "libppo_ba.so")
        if not os.path.exists(path):
            raise EngineError(f"{path} missing: run __graft_entry__.build() (nvcc, sm_100a). No CPU fallback exists.")
        lib = C.CDLL(path)
        _bind(lib, "ppo_ba_")
        lib.ppo_ba_create.argtypes = [C.POINTER(A.Params), C.c_int, C.POINTER(C.c_void_p)]
        lib.ppo_ba_last_error.argtypes = [C.c_void_p]
        lib.ppo_ba_last_error.restype = C.c_char_p
        lib.ppo_ba_set_profiling.argtypes = [C.c_void_p, C.c_int]
        lib.ppo_ba_launch_count.argtypes = [C.c_void_p]
        lib.ppo_ba_launch_count.restype = C.c_longlong
        lib.ppo_ba_host_sync_count.argtypes = [C.c_void_p]
        lib.ppo_ba_host_sync_count.restype = C.c_longlong
        lib.ppo_ba_set_graph_mode.argtypes = [C.c_void_p, C.c_int]
        lib.ppo_ba_set_params.argtypes = [C.c_void_p, C.POINTER(A.Params)]
        lib.ppo_ba_time_assembly.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        lib.ppo_ba_time_solve.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int)]
        lib.ppo_ba_mark.argtypes = [C.c_void_p, C.c_int]
        lib.ppo_ba_elapsed_ms.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        lib.ppo_ba_flush_l2.argtypes = [C.c_void_p]
        lib.ppo_ba_set_shard.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        lib.ppo_ba_nccl_unique_id.argtypes = [C.c_char_p]
        lib.ppo_ba_nccl_init.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        lib.ppo_ba_nccl_destroy.argtypes = [C.c_void_p]
        lib.ppo_ba_collective_count.argtypes = [C.c_void_p]
        lib.ppo_ba_collective_count.restype = C.c_longlong
        lib.ppo_ba_debug_linearize.argtypes = [C.c_void_p, C.POINTER(C.c_int32 * 2), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ppo_ba_debug_solve.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]
        lib.ppo_ba_host_register.argtypes = [C.c_void_p, C.c_size_t]
        lib.ppo_ba_host_unregister.argtypes = [C.c_void_p]
        lib.ppo_ba_point_edge_outliers.argtypes = [C.c_void_p, C.c_double, C.c_double, C.POINTER(C.POINTER(C.c_int32)), C.POINTER(C.c_int32)]
        lib.ppo_ba_debug_dense_solve.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]
        _LIB = lib
    return _LIB


def default_params(lib=None, prefix="ppo_ba_"):
    lib = lib or load_library()
    p = A.Params()
    getattr(lib, prefix + "default_params")(C.byref(p))
    return p


class Handle:
    """Thin object wrapper over a ppo_ba-shaped C API (the engine, or — in tests — the oracle)."""

    def __init__(self, lib, prefix, handle):
        self.lib, self.prefix, self.h = lib, prefix, handle
        self.graph = None

    def _f(self, name):
        return getattr(self.lib, self.prefix + name)

    def _check(self, rc, what):
        if rc != A.PPO_OK:
            msg = ""
            if self.prefix == "ppo_ba_":
                msg = (self.lib.ppo_ba_last_error(self.h) or b"").decode()
            raise EngineError(f"{self.prefix}{what} failed: rc={rc} {msg}")

    def set_graph(self, g):
        self.graph = g  # keep the numpy arrays alive
        self._check(self._f("set_graph")(self.h, C.byref(g.c)), "set_graph")

    def optimize(self, iters, stop_flag=None):
        st = A.Stats()
        ptr = stop_flag.ctypes.data if stop_flag is not None else None
        rc = self._f("optimize")(self.h, iters, ptr, C.byref(st))
        if rc != A.PPO_E_EMPTY:
            self._check(rc, "optimize")
        return st

    def edge_count(self, kind):
        return self._f("edge_count")(self.h, kind)

    def edge_chi2(self, kind):
        n = self.edge_count(kind)
        chi2 = np.zeros(n)
        dpos = np.zeros(n, np.uint8)
        norm = np.zeros(n)
        self._check(self._f("edge_chi2")(self.h, kind, chi2.ctypes.data, dpos.ctypes.data, norm.ctypes.data), "edge_chi2")
        return chi2, dpos, norm

    def get_edge_flags(self, kind):
        f = np.zeros(self.edge_count(kind), np.uint8)
        self._check(self._f("get_edge_flags")(self.h, kind, f.ctypes.data), "get_edge_flags")
        return f

    def set_edge_flags(self, kind, flags):
        flags = np.ascontiguousarray(flags, np.uint8)
        assert len(flags) == self.edge_count(kind)
        self._check(self._f("set_edge_flags")(self.h, kind, flags.ctypes.data), "set_edge_flags")

    def recompute_edge_errors(self, kind):
        """e->computeError() on the level-1 edges of a kind (Optimizer::PoseOptimization, Optimizer.cc:400-403)."""
        self._check(self._f("recompute_edge_errors")(self.h, kind), "recompute_edge_errors")

    def outlier_pass(self):
        n = (C.c_int32 * 3)()
        self._check(self._f("outlier_pass")(self.h, C.byref(n)), "outlier_pass")
        return list(n)

    def local_ba(self, stop_flag=None):
        res = A.Result()
        ptr = stop_flag.ctypes.data if stop_flag is not None else None
        self._check(self._f("local_ba")(self.h, ptr, C.byref(res)), "local_ba")
        return res

    def get_state(self):
        s = A.StateArrays(self.graph.c)
        self._check(self._f("get_state")(self.h, C.byref(s.c)), "get_state")
        return s

    def reset(self):
        self._check(self._f("reset")(self.h), "reset")

    def debug_linearize(self):
        dims = (C.c_int32 * 2)()
        f = self._f("debug_linearize")
        self._check(f(self.h, C.byref(dims), None, None, None, None), "debug_linearize")
        n_p, n_l = dims[0], dims[1]
        Hpp = np.zeros((n_p, n_p))
        b = np.zeros(n_p + 3 * n_l)
        Hll = np.zeros((n_l, 9))
        chi2 = C.c_double()
        self._check(f(self.h, C.byref(dims), Hpp.ctypes.data, b.ctypes.data, Hll.ctypes.data, C.addressof(chi2)), "debug_linearize")
        return dict(n_p=n_p, n_l=n_l, Hpp=Hpp, b=b, Hll=Hll, chi2=chi2.value)

    def debug_solve(self, lam, n_p, n_l, solve=True):
        S = np.zeros((n_p, n_p))
        bs = np.zeros(n_p)
        x = np.zeros(n_p + 3 * n_l) if solve else None
        ok = C.c_int32()
        self._check(self._f("debug_solve")(self.h, C.c_double(lam), S.ctypes.data, bs.ctypes.data, x.ctypes.data if solve else None, C.byref(ok)), "debug_solve")
        return dict(Hschur=S, bschur=bs, x=x, ok=ok.value)

    def point_edge_outliers(self, chi2_mono, chi2_stereo):
        """Ascending indices of the point edges with chi2 above the threshold of their kind or a non-positive depth."""
        p = C.POINTER(C.c_int32)()
        n = C.c_int32()
        self._check(self._f("point_edge_outliers")(self.h, C.c_double(chi2_mono), C.c_double(chi2_stereo), C.byref(p), C.byref(n)), "point_edge_outliers")
        return np.array(p[:n.value], dtype=np.int32)

    def debug_dense_solve(self, A, b):
        """The handle's linear solver alone on a dense symmetric system (upper triangle of A used): (x, ok)."""
        A = np.ascontiguousarray(A, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.zeros(len(b))
        ok = C.c_int32()
        self._check(self._f("debug_dense_solve")(self.h, len(b), A.ctypes.data, b.ctypes.data, x.ctypes.data, C.byref(ok)), "debug_dense_solve")
        return x, ok.value

    def close(self):
        if self.h:
            self._f("destroy")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class LocalBA(Handle):
    """The B200 engine. Mirrors the reference call: see module docstring."""

    def __init__(self, params=None, device=0):
        lib = load_library()
        p = params or default_params(lib)
        h = C.c_void_p()
        rc = lib.ppo_ba_create(C.byref(p), device, C.byref(h))
        if rc != A.PPO_OK:
            raise EngineError(f"ppo_ba_create failed rc={rc} (PPO_E_NOGPU={A.PPO_E_NOGPU}): the engine needs a CUDA device")
        super().__init__(lib, "ppo_ba_", h)
        self.params = p

    def set_profiling(self, on):
        self.lib.ppo_ba_set_profiling(self.h, int(on))

    def launch_count(self):
        return int(self.lib.ppo_ba_launch_count(self.h))

    def host_sync_count(self):
        return int(self.lib.ppo_ba_host_sync_count(self.h))

    def set_graph_mode(self, enable):
        self._check(self.lib.ppo_ba_set_graph_mode(self.h, int(bool(enable))), "set_graph_mode")

    def time_assembly(self, reps=20):
        ms, by = C.c_double(), C.c_double()
        self._check(self.lib.ppo_ba_time_assembly(self.h, reps, C.byref(ms), C.byref(by)), "time_assembly")
        return ms.value, by.value

    def time_solve(self, reps=10):
        ms, fl, n = C.c_double(), C.c_double(), C.c_int()
        self._check(self.lib.ppo_ba_time_solve(self.h, reps, C.byref(ms), C.byref(fl), C.byref(n)), "time_solve")
        return ms.value, fl.value, n.value

    def mark(self, which):
        self._check(self.lib.ppo_ba_mark(self.h, which), "mark")

    def elapsed_ms(self):
        ms = C.c_double()
        self._check(self.lib.ppo_ba_elapsed_ms(self.h, C.byref(ms)), "elapsed_ms")
        return ms.value

    def flush_l2(self):
        self._check(self.lib.ppo_ba_flush_l2(self.h), "flush_l2")

    def set_shard(self, comm, rank, world):
        self._check(self.lib.ppo_ba_set_shard(self.h, comm, rank, world), "set_shard")

    def collective_count(self):
        return int(self.lib.ppo_ba_collective_count(self.h))

    # reference-named entry points ---------------------------------------------------------------
    def LocalBACameraPlaneCuboids(self, graph, stop_flag=None):
        """Stages B-F of Optimizer::LocalBACameraPlaneCuboids on a flat graph; returns (state, result)."""
        self.set_graph(graph)
        res = self.local_ba(stop_flag)
        return self.get_state(), res

    LocalBundleAdjustment = LocalBACameraPlaneCuboids


def local_ba_batch(engines, stop_flag=None):
    """ppo_ba_local_ba_batch: the windows of several LocalBA handles optimised concurrently (one host thread each inside
    the library).  Returns the list of ppo_ba_result."""
    lib = load_library()
    n = len(engines)
    hs = (C.c_void_p * n)(*[e.h for e in engines])
    res = (A.Result * n)()
    lib.ppo_ba_local_ba_batch.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.POINTER(A.Result)]
    sp = stop_flag.ctypes.data if stop_flag is not None else None
    rc = lib.ppo_ba_local_ba_batch(hs, n, sp, res)
    if rc != A.PPO_OK:
        raise EngineError(f"ppo_ba_local_ba_batch failed: rc={rc}")
    return list(res)


def nccl_unique_id():
    buf = C.create_string_buffer(128)
    rc = load_library().ppo_ba_nccl_unique_id(buf)
    if rc != A.PPO_OK:
        raise EngineError(f"ppo_ba_nccl_unique_id rc={rc} (libnccl.so.2 not loadable?)")
    return buf.raw


def nccl_init(id_bytes, rank, world, device):
    comm = C.c_void_p()
    rc = load_library().ppo_ba_nccl_init(id_bytes, rank, world, device, C.byref(comm))
    if rc != A.PPO_OK:
        raise EngineError(f"ppo_ba_nccl_init rc={rc}")
    return comm


def nccl_destroy(comm):
    load_library().ppo_ba_nccl_destroy(comm)
