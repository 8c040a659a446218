"""Python binding of the synthetic window generator (include/ppo_synth.h)."""
import ctypes as C
import os

from . import _abi as A

_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(A.PKG, "lib", "libppo_synth.so")
        if not os.path.exists(path):
            from . import _build
            _build.build_synth()
        lib = C.CDLL(path)
        lib.ppo_synth_config.argtypes = [C.c_int, C.c_int, C.POINTER(A.SynthCfg)]
        lib.ppo_synth_config.restype = None
        lib.ppo_synth_create.argtypes = [C.POINTER(A.SynthCfg)]
        lib.ppo_synth_create.restype = C.c_void_p
        lib.ppo_synth_graph.argtypes = [C.c_void_p]
        lib.ppo_synth_graph.restype = C.POINTER(A.Graph)
        lib.ppo_synth_truth.argtypes = [C.c_void_p, C.POINTER(A.State)]
        lib.ppo_synth_truth.restype = None
        lib.ppo_synth_destroy.argtypes = [C.c_void_p]
        lib.ppo_synth_destroy.restype = None
        _LIB = lib
    return _LIB


def config(index, window=0, **overrides):
    """BASELINE.json configs[index] as a SynthCfg; keyword overrides replace fields."""
    cfg = A.SynthCfg()
    _lib().ppo_synth_config(index, window, C.byref(cfg))
    for k, v in overrides.items():
        setattr(cfg, k, v)
    return cfg


def make_graph(cfg, with_truth=False):
    """Generates the window and deep-copies it into numpy arrays (GraphArrays)."""
    lib = _lib()
    s = lib.ppo_synth_create(C.byref(cfg))
    try:
        g = A.GraphArrays.from_c(lib.ppo_synth_graph(s).contents)
        if with_truth:
            t = A.StateArrays(g.c)
            lib.ppo_synth_truth(s, C.byref(t.c))
            return g, t
        return g
    finally:
        lib.ppo_synth_destroy(s)
