#!/usr/bin/env python3
"""Build the UNMODIFIED reference extension (QGTC_PPoPP22/QGTC_module) for sm_100a.

TEST INFRASTRUCTURE ONLY.  The sources are compiled where they lie under
/root/reference (nothing is copied into this repo); the only outputs are the
object files and ``QGTC_ref*.so`` under ``oracle/_ref/`` (git-ignored, but it
travels to the GPU box with gpurun).  The module is named ``QGTC_ref`` so it
cannot shadow this repo's own ``QGTC`` extension.

Used by: tests/test_reference_gpu.py (pins the CPU oracle and our kernels against
the reference's own kernels on a B200) and ``bench.py --impl reference``.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("QGTC_REFERENCE_DIR", "/root/reference/QGTC_module")
OUT = os.path.join(HERE, "_ref")


def build(verbose: bool = False) -> str:
    srcs = [os.path.join(REF, "QGTC_host.cpp"), os.path.join(REF, "QGTC_device.cu")]
    if not all(os.path.exists(s) for s in srcs):
        raise FileNotFoundError("reference sources not found under %s" % REF)
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", "4")
    from torch.utils.cpp_extension import load

    mod = load(
        name="QGTC_ref",
        sources=srcs,
        extra_include_paths=[REF],
        extra_cflags=["-O2", "-w"],
        extra_cuda_cflags=["-O3", "-w", "-gencode", "arch=compute_100a,code=sm_100a"],
        build_directory=OUT,
        verbose=verbose,
        is_python_module=False,   # do not import (importing needs libcuda symbols at dlopen only lazily; keep CPU box safe)
    )
    so = os.path.join(OUT, "QGTC_ref.so")
    return so if os.path.exists(so) else str(mod)


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
