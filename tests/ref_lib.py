"""tests-only binding of oracle/_ref/libppo_g2o_ref.so: the reference's OWN g2o + vertex/edge sources, compiled unmodified from
/root/reference by oracle/Makefile.ref against the Eigen stand-in (oracle/ref_stub), driven by oracle/ref_driver.cpp.
Available where it has been built (this container; the .so travels to the GPU box).  Never imported by the product."""
import ctypes as C
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ppo_pkg import ppo  # noqa: E402

A = ppo.abi
PATH = os.path.join(A.ROOT, "oracle", "_ref", "libppo_g2o_ref.so")
_LIB = None


def available(build=True):
    if os.path.exists(PATH):
        return True
    if build and os.path.isdir("/root/reference/Thirdparty/g2o"):
        subprocess.run(["make", "-f", "oracle/Makefile.ref", "-j8"], cwd=A.ROOT, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return os.path.exists(PATH)


def lib():
    global _LIB
    if _LIB is None:
        if not available():
            raise RuntimeError("oracle/_ref/libppo_g2o_ref.so not built (needs /root/reference): make -f oracle/Makefile.ref")
        L = C.CDLL(PATH)
        ppo.engine._bind(L, "ppo_ref_")
        L.ppo_ref_create.argtypes = [C.POINTER(A.Params), C.POINTER(C.c_void_p)]
        _LIB = L
    return _LIB


def default_params():
    return ppo.engine.default_params(lib(), "ppo_ref_")


class Ref(ppo.Handle):
    """The reference's g2o solve of a flat window (same wrapper methods as the engine and the oracle)."""

    def __init__(self, params=None):
        L = lib()
        p = params or default_params()
        h = C.c_void_p()
        assert L.ppo_ref_create(C.byref(p), C.byref(h)) == 0
        super().__init__(L, "ppo_ref_", h)
        self.params = p
