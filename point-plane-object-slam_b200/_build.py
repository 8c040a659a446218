"""Builds the native libraries in-tree (they travel to the GPU box with the snapshot).

  lib/libppo_ba.so     CUDA engine + C-ABI (include/ppo_ba.h), sm_100a only
  lib/libppo_synth.so  synthetic window generator (include/ppo_synth.h), host C++
  oracle/_build/libppo_oracle.so  CPU restatement (test infrastructure)
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
LIB = os.path.join(PKG, "lib")
CSRC = os.path.join(PKG, "csrc")
ORACLE = os.path.join(ROOT, "oracle")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + "\n")
        raise RuntimeError("build failed: " + cmd[0])
    return r.stdout


def _sources(d, exts):
    out = []
    for base, _, files in os.walk(d):
        for f in files:
            if f.endswith(exts):
                out.append(os.path.join(base, f))
    return sorted(out)


def build_synth(force=False):
    os.makedirs(LIB, exist_ok=True)
    out = os.path.join(LIB, "libppo_synth.so")
    srcs = [os.path.join(CSRC, "host", "ppo_synth.cpp")]
    deps = srcs + [os.path.join(CSRC, "host", "ppo_convert.h"), os.path.join(ROOT, "include", "ppo_synth.h"),
                   os.path.join(ROOT, "include", "ppo_ba.h")]
    if force or _newer(out, deps):
        _run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out] + srcs)
    return out


def build_oracle(force=False):
    bdir = os.path.join(ORACLE, "_build")
    os.makedirs(bdir, exist_ok=True)
    out = os.path.join(bdir, "libppo_oracle.so")
    srcs = [os.path.join(ORACLE, "ppo_oracle.cpp")]
    deps = srcs + [os.path.join(ORACLE, "ppo_oracle_math.h"), os.path.join(ROOT, "include", "ppo_ba.h")]
    if force or _newer(out, deps):
        # -ffp-contract=off: same arithmetic as the reference build (no FMA contraction on x86-64 -O2/-O3 without -march)
        _run(["g++", "-O3", "-std=c++17", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", "-o", out] + srcs)
    return out


def build_cuda(force=False):
    os.makedirs(LIB, exist_ok=True)
    out = os.path.join(LIB, "libppo_ba.so")
    cu = _sources(os.path.join(CSRC, "cuda"), (".cu",))
    deps = cu + _sources(os.path.join(CSRC, "cuda"), (".cuh", ".h")) + [os.path.join(ROOT, "include", "ppo_ba.h")]
    if force or _newer(out, deps):
        if not (os.path.exists(NVCC) or shutil.which("nvcc")):
            raise RuntimeError("nvcc not found; the engine has no CPU fallback")
        nvcc = NVCC if os.path.exists(NVCC) else "nvcc"
        torch_lib = None
        nccl_inc = []
        nccl_link = []
        try:  # NCCL shipped with the torch wheel (nvidia-nccl-cu12): headers + libnccl.so.2
            import nvidia.nccl as _n
            base = os.path.dirname(_n.__file__) if getattr(_n, "__file__", None) else list(_n.__path__)[0]
            inc = os.path.join(base, "include")
            lib = os.path.join(base, "lib")
            if os.path.exists(os.path.join(inc, "nccl.h")):
                nccl_inc = ["-I", inc, "-DPPO_HAVE_NCCL=1"]
                so = [f for f in os.listdir(lib) if f.startswith("libnccl.so")]
                if so:
                    nccl_link = ["-L", lib, "-l:" + so[0], "-Xlinker", "-rpath=" + lib]
        except Exception:
            pass
        _run([nvcc] + NVCC_FLAGS + nccl_inc + ["-shared", "-o", out] + cu + nccl_link + ["-lcudart", "-lpthread"])
    return out


def build_shim(force=False):
    """Optimizer.cc replacement (host shim) + the mock-map harness it is tested with; links the C-ABI library."""
    os.makedirs(LIB, exist_ok=True)
    out = os.path.join(LIB, "libppo_shim_mock.so")
    host = os.path.join(CSRC, "host")
    srcs = [os.path.join(host, "ppo_optimizer_shim.cpp"), os.path.join(host, "ppo_mock_world.cpp")]
    deps = srcs + [os.path.join(host, "ppo_mock_slam.h"), os.path.join(host, "ppo_convert.h"), os.path.join(ROOT, "include", "ppo_ba.h"),
                   os.path.join(LIB, "libppo_ba.so")]
    if force or _newer(out, deps):
        _run(["g++", "-O2", "-std=c++17", "-pthread", "-shared", "-fPIC", "-o", out] + srcs + ["-L", LIB, "-lppo_ba", "-Wl,-rpath,$ORIGIN"])
    return out


def build_ref(force=False):
    """oracle/_ref/libppo_g2o_ref.so: the reference's own g2o + vertex/edge sources compiled where they lie under /root/reference
    (oracle/Makefile.ref, Eigen stand-in oracle/ref_stub).  Only possible where /root/reference exists (this container); on the GPU
    box the prebuilt file is used.  Test infrastructure: building the checker is not using it."""
    out = os.path.join(ORACLE, "_ref", "libppo_g2o_ref.so")
    if not os.path.isdir("/root/reference/Thirdparty/g2o"):
        return out if os.path.exists(out) else None
    cmd = ["make", "-f", os.path.join("oracle", "Makefile.ref"), "-j8"]
    if force:
        cmd.insert(1, "-B")
    r = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout[-4000:] + "\n")
        raise RuntimeError("build failed: oracle/Makefile.ref")
    return out


def build_all(force=False):
    out = {"synth": build_synth(force), "oracle": build_oracle(force), "cuda": build_cuda(force)}
    out["shim"] = build_shim(force)
    out["ref"] = build_ref(force)
    return out


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv))
