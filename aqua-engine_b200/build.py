"""Build the native libraries in-tree (so the .so files travel to the GPU box with gpurun).

  libaqua_cuda.so   nvcc, sm_100a only  — the product (kernels + C ABI, include/aqua_cuda.h)
  libaqua_host.so   g++                 — scene ingest, C++ mirror of the Rust host (include/aqua_host.h)

Floating point: device code is compiled with -fmad=false and host code with
-ffp-contract=off so the single-sourced definitions in csrc/aq_core.h round identically.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-O3,-ffp-contract=off", "-shared",
    "-I", os.path.join(ROOT, "include"), "-I", CSRC,
]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd, **kw):
    print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd, **kw)


def build_cuda(force=False, verbose=False, defines=(), name="libaqua_cuda.so"):
    """`defines`/`name` build an A/B variant (e.g. defines=["AQ_EXP_X=1"], name="libaqua_cuda_x.so");
    select it at run time with AQUA_CUDA_LIB=<name>."""
    out = os.path.join(HERE, name)
    srcs = [os.path.join(CSRC, f) for f in ("aq_cuda.cu", "aq_multi.cu", "aq_bvh_build_gpu.cu", "aq_resolve.cu", "aq_bvh_build.cpp")]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh", ".inl"))]
    deps.append(os.path.join(ROOT, "include", "aqua_cuda.h"))
    if not force and not _newer(out, deps):
        return out
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    flags = list(NVCC_FLAGS)
    if "AQ_TOLERANCE_BUILD=1" in defines:
        # A/B only (what does bit-exactness cost?): contraction on, approximate div/sqrt.
        # Results then agree with the oracle within a tolerance, not bit for bit; never the shipped library.
        flags = [{"-fmad=false": "-fmad=true", "-prec-div=true": "-prec-div=false", "-prec-sqrt=true": "-prec-sqrt=false"}.get(f, f) for f in flags]
    cmd = [nvcc] + flags + [f"-D{d}" for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + srcs + ["-ldl"]
    _run(cmd)
    return out


def build_host(force=False):
    out = os.path.join(HERE, "libaqua_host.so")
    srcs = [os.path.join(HOST, f) for f in ("aq_host.cpp", "aq_jpeg.cpp", "aq_import.cpp")]
    deps = srcs + [os.path.join(ROOT, "include", "aqua_host.h"), os.path.join(ROOT, "include", "aqua_cuda.h")]
    if not force and not _newer(out, deps):
        return out
    _run(["g++", "-O2", "-std=c++17", "-fPIC", "-Wall", "-I", os.path.join(ROOT, "include"),
          "-shared", "-o", out] + srcs)
    return out


def build_all(force=False, verbose=False):
    return build_cuda(force, verbose), build_host(force)


if __name__ == "__main__":
    if "--variant" in sys.argv:  # python build.py --variant NAME DEF1 DEF2 ...
        i = sys.argv.index("--variant")
        build_cuda(force=True, verbose="-v" in sys.argv, defines=[a for a in sys.argv[i + 2:] if a != "-v"],
                   name=f"libaqua_cuda_{sys.argv[i + 1]}.so")
    else:
        build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
