"""In-tree build of the native pieces (sm_100a only).

  libqgtc_b200.so   the C-ABI CUDA library (include/qgtc_b200.h), built with nvcc directly
  QGTC*.so          the PyTorch extension mirroring the reference's `QGTC` module, a thin
                    pybind11 layer over the C ABI (built with the host compiler, links the library)

Both land in qgtc_ppopp22_b200/lib/ so that they travel with the source tree (no JIT cache).
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
INCLUDE = os.path.join(ROOT, "include")

CUDA_SOURCES = ["abi.cu", "pack.cu", "gemm_mma.cu", "gemm_tc05.cu", "gemm_grp.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _nvcc() -> str:
    return os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def lib_path() -> str:
    return os.path.join(LIBDIR, "libqgtc_b200.so")


def build_library(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + [os.path.join(INCLUDE, "qgtc_b200.h")]
    objs, jobs = [], []
    for src in CUDA_SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(LIBDIR, src.replace(".cu", ".o"))
        if force or not _newer(o, [s] + headers):
            jobs.append((src, o, [_nvcc()] + NVCC_FLAGS + ["-I", INCLUDE, "-I", CSRC, "-c", s, "-o", o]))
        objs.append(o)
    # the translation units are independent: compile them side by side (the two tcgen05 engines take ~1.5 min each)
    procs = [(src, o, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)) for src, o, cmd in jobs]
    failed = []
    for src, o, pr in procs:
        out, err = pr.communicate()
        if verbose or pr.returncode:
            sys.stderr.write(out + err)
        if pr.returncode:
            failed.append(src)
        else:
            with open(o + ".ptxas.txt", "w") as f:
                f.write(err)
    if failed:
        raise RuntimeError("nvcc failed for %s" % ", ".join(failed))
    so = lib_path()
    if force or not _newer(so, objs):
        cmd = [_nvcc(), "-shared", "-o", so] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
        subprocess.check_call(cmd)
    return so


def ext_path() -> str:
    suffix = sysconfig.get_config_var("EXT_SUFFIX") or ".so"
    return os.path.join(LIBDIR, "QGTC" + suffix)


def build_extension(force: bool = False, verbose: bool = False) -> str:
    """pybind11/libtorch module named QGTC (drop-in for the reference's extension)."""
    import torch
    from torch.utils import cpp_extension

    src = os.path.join(CSRC, "qgtc_torch.cpp")
    out = ext_path()
    if not force and _newer(out, [src, os.path.join(INCLUDE, "qgtc_b200.h"), lib_path()]):
        return out
    inc = cpp_extension.include_paths() + [INCLUDE, sysconfig.get_paths()["include"], "/usr/local/cuda/include"]
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    cxx = os.environ.get("CXX", "g++")
    cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-w", "-DTORCH_EXTENSION_NAME=QGTC",
           "-DTORCH_API_INCLUDE_EXTENSION_H", "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)]
    for i in inc:
        cmd += ["-isystem", i]
    cmd += [src, "-o", out, "-L", LIBDIR, "-lqgtc_b200", "-Wl,-rpath,$ORIGIN",
            "-L", tlib, "-lc10", "-ltorch_cpu", "-ltorch", "-ltorch_python", "-lc10_cuda",
            "-L", "/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath," + tlib]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("building the QGTC torch extension failed")
    return out


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_library(force, verbose)
    build_extension(force, verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(lib_path())
    print(ext_path())
