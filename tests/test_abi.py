"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/ppo_ba.h declares; struct layouts match between C and the ctypes mirror; no compute calls."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header, prefix):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(" + prefix + r"\w+)\s*\(", txt)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as ge
    return ge.build()


def test_library_exports_every_declared_symbol(built):
    lib = C.CDLL(built["cuda"])
    names = _declared("ppo_ba.h", "ppo_ba_")
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ppo_ba.h but not exported"
    syn = C.CDLL(built["synth"])
    for n in _declared("ppo_synth.h", "ppo_synth_"):
        assert hasattr(syn, n), n


def test_struct_layouts_match_c(ppo, tmp_path):
    """sizeof / offsetof of the ctypes mirrors == the C compiler's."""
    A = ppo.abi
    src = tmp_path / "lay.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "ppo_ba.h"\n#include "ppo_synth.h"\nint main(){'
                   'printf("%zu %zu %zu %zu %zu %zu %zu ", sizeof(ppo_ba_params), sizeof(ppo_ba_graph), sizeof(ppo_ba_state), sizeof(ppo_ba_iter), sizeof(ppo_ba_stats), sizeof(ppo_ba_result), sizeof(ppo_synth_cfg));'
                   'printf("%zu %zu %zu %zu %zu\\n", offsetof(ppo_ba_params, lm_max_trials), offsetof(ppo_ba_graph, pt_rowptr), offsetof(ppo_ba_graph, cpe_info), offsetof(ppo_ba_stats, trace), offsetof(ppo_ba_result, n_outlier_point_edges));'
                   'return 0;}')
    exe = tmp_path / "lay"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(A.Params), C.sizeof(A.Graph), C.sizeof(A.State), C.sizeof(A.Iter), C.sizeof(A.Stats), C.sizeof(A.Result), C.sizeof(A.SynthCfg),
            A.Params.lm_max_trials.offset, A.Graph.pt_rowptr.offset, A.Graph.cpe_info.offset, A.Stats.trace.offset, A.Result.n_outlier_point_edges.offset]
    assert got == want


def test_engine_fails_loudly_without_gpu(ppo):
    """No CPU fallback: without a CUDA device ppo_ba_create returns PPO_E_NOGPU and the wrapper raises."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(ppo.EngineError):
        ppo.LocalBA()


def test_default_params_agree_between_engine_and_oracle(ppo, oracle_mod):
    pe, po = ppo.default_params(), oracle_mod.default_params()
    for name, _ in ppo.abi.Params._fields_:
        assert getattr(pe, name) == getattr(po, name), name
    assert pe.huber_mono == float(__import__("numpy").float32(5.991 ** 0.5))  # "const float thHuberMono = sqrt(5.991)"


def test_synth_is_deterministic_and_matches_survey_sizes(ppo):
    import numpy as np
    g1 = ppo.synth.make_graph(ppo.synth.config(0))
    g2 = ppo.synth.make_graph(ppo.synth.config(0))
    for k in g1.a:
        assert np.array_equal(g1[k], g2[k]), k
    assert g1.c.n_kf == 12 and g1.c.n_pt == 2000 and abs(g1.c.n_pe - 12000) < 100
    assert g1["kf_fixed"].sum() == 3  # id 0 + 2 fixed cameras
    assert not np.array_equal(ppo.synth.make_graph(ppo.synth.config(3, window=1))["pt_xyz"], ppo.synth.make_graph(ppo.synth.config(3, window=2))["pt_xyz"])


def test_product_libraries_never_reference_the_oracle(built):
    """The oracle is test infrastructure: no product library may link it or import one of its symbols (the oracle-backed
    shim of tests/shim_lib.py lives under oracle/_build, not under the package's lib/)."""
    for key in ("cuda", "shim", "synth"):
        path = built[key]
        syms = subprocess.check_output(["nm", "-D", path], text=True)
        assert "ppo_oracle" not in syms, path
        needed = subprocess.check_output(["readelf", "-d", path], text=True)
        assert "libppo_oracle" not in needed, path
    assert not any("oracle" in f for f in os.listdir(os.path.dirname(built["cuda"])))


def test_public_headers_compile_as_c_and_cxx():
    """The drop-in boundary is a C-ABI: include/*.h must be valid C99 (plain pointers and sizes, no C++ types) and valid C++."""
    import glob
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for h in sorted(glob.glob(os.path.join(root, "include", "*.h"))):
        for cmd in (["gcc", "-std=c99", "-Wall", "-fsyntax-only", "-x", "c", h], ["g++", "-std=c++17", "-Wall", "-fsyntax-only", "-x", "c++", h]):
            r = subprocess.run(cmd, capture_output=True, text=True)
            assert r.returncode == 0, (cmd, r.stderr)


def test_engine_links_no_vendor_solver_library(built):
    """Every kernel on the path is the repo's own: the engine library links the CUDA runtime only -- no cuBLAS / cuSOLVER / cuSPARSE /
    cuDNN (NCCL is looked up at run time for sharded windows only)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = os.path.join(root, "point-plane-object-slam_b200", "lib", "libppo_ba.so")
    out = subprocess.run(["ldd", lib], capture_output=True, text=True).stdout.lower()
    for name in ("cublas", "cusolver", "cusparse", "cudnn", "cutlass", "nccl"):
        assert name not in out, name
    assert "libcudart" in out
