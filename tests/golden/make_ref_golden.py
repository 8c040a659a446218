"""Golden vectors produced by the REFERENCE's own code (not by the oracle).

oracle/_ref/libppo_g2o_ref.so is the reference's unmodified g2o + vertex/edge sources compiled from /root/reference
(oracle/Makefile.ref; Eigen replaced by the stand-in in oracle/ref_stub because Eigen is not installed here).  This script
runs it on seeded inputs and stores inputs + outputs:

  ref_functions.npz   per-function vectors: SE3Quat::exp, VertexSE3Expmap::oplusImpl, SE3Quat::map, Plane3D normalize / oplus /
                      ominus / ominus_ver / ominus_par / rigid transform, VertexCuboid::oplusImpl (yaw-only, fixed height),
                      compute3D_BoxCorner, projectOntoImage / projectOntoImageBbox, EdgePointCuboidOnlyObject::computeError,
                      toMinimalVector, RobustKernelHuber::robustify, EdgeSE3ProjectXYZ / EdgeStereoSE3ProjectXYZ computeError +
                      linearizeOplus (analytic Jacobians, float 1/z quirk), the three plane edges and the two camera-cuboid edges
  ref_<case>.npz      whole optimize(5) -> re-levelling -> optimize(10) runs of the three windows of make_golden.py: final
                      estimates, per-edge chi2 and flags, iteration counts, lambda

tests/test_ref_pin.py checks the ORACLE against these files everywhere (they travel with the repository), and against the
library itself where it is built.   Run from the repo root:   python tests/golden/make_ref_golden.py
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from ppo_pkg import ppo  # noqa: E402
import ref_lib  # noqa: E402
import ref_cases  # noqa: E402

A = ppo.abi


def main():
    L = ref_lib.lib()
    fn = ref_cases.function_vectors(L, "ppo_ref_")
    np.savez_compressed(os.path.join(HERE, "ref_functions.npz"), **fn)
    print("ref_functions.npz:", len(fn), "arrays")
    for name in ref_cases.WINDOWS:
        out = ref_cases.run_window(name, lambda p: ref_lib.Ref(p), ref_lib.default_params())
        np.savez_compressed(os.path.join(HERE, "ref_" + name + ".npz"), **out)
        print("ref_%s.npz: iterations %s, chi2 %.6f -> %.6f" % (name, out["iterations"].tolist(), out["chi2"][0], out["chi2"][1]))


if __name__ == "__main__":
    main()
