#!/usr/bin/env python3
"""Build the reference extension (QGTC_PPoPP22/QGTC_module) for sm_100a.  TEST INFRASTRUCTURE ONLY.

The sources are compiled where they lie under /root/reference (nothing is copied into this
repo); the only outputs are object files and shared objects under ``oracle/_ref/`` (git-ignored,
but they travel to the GPU box with gpurun).  Two variants:

  QGTC_ref     the UNMODIFIED reference (module name changed by -DTORCH_EXTENSION_NAME only, so
               it cannot shadow this repo's own ``QGTC`` extension).  Timed by
               ``bench.py --impl reference``.
  QGTC_ref_lw  same sources, compiled with ``-include oracle/ref_logical_warpid.h`` so the kernels
               use the logical warp index instead of %warpid (see that header).  This is the
               variant whose outputs are a trustworthy GPU oracle on a B200.

Used by tests/test_reference_gpu.py, tests/golden/make_golden_from_reference.py and bench.py.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("QGTC_REFERENCE_DIR", "/root/reference/QGTC_module")
OUT = os.path.join(HERE, "_ref")

VARIANTS = {
    "QGTC_ref": [],
    "QGTC_ref_lw": ["-include", os.path.join(HERE, "ref_logical_warpid.h")],
}


def so_path(name: str) -> str:
    return os.path.join(OUT, name, name + ".so")


def sources_present() -> bool:
    return all(os.path.exists(os.path.join(REF, f)) for f in ("QGTC_host.cpp", "QGTC_device.cu", "kernel.h"))


def build(name: str = "QGTC_ref", verbose: bool = False) -> str:
    srcs = [os.path.join(REF, "QGTC_host.cpp"), os.path.join(REF, "QGTC_device.cu")]
    if not sources_present():
        raise FileNotFoundError("reference sources not found under %s" % REF)
    bdir = os.path.join(OUT, name)
    os.makedirs(bdir, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", "4")
    from torch.utils.cpp_extension import load

    load(
        name=name,
        sources=srcs,
        extra_include_paths=[REF],
        extra_cflags=["-O2", "-w"],
        extra_cuda_cflags=["-O3", "-w", "-gencode", "arch=compute_100a,code=sm_100a"] + VARIANTS[name],
        build_directory=bdir,
        verbose=verbose,
        is_python_module=False,
    )
    return so_path(name)


def load_variant(name: str = "QGTC_ref"):
    """Import a prebuilt variant as a python module (GPU box: no rebuild, /root/reference is absent)."""
    import importlib.util

    import torch  # noqa: F401  (libtorch must be loaded first)

    path = so_path(name)
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    names = [a for a in sys.argv[1:] if not a.startswith("-")] or list(VARIANTS)
    for n in names:
        print(build(n, verbose="-v" in sys.argv))
