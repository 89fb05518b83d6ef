"""qgtc_ppopp22_b200 -- B200-native (sm_100a) implementation of the QGTC quantized-GNN hot path.

Importing this package makes the drop-in extension importable under the reference's name:

    import qgtc_ppopp22_b200          # registers sys.modules["QGTC"]
    import QGTC                       # val2bit / bit2val / bitMM2Bit / bitMM2Int / ... (QGTC_host.cpp:259-271)

``qgtc_ppopp22_b200.capi`` is the same surface through the C ABI (include/qgtc_b200.h) with ctypes.
There is no CPU fallback: if the native library is missing, importing the ops raises.
"""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.util
import os
import sys

__all__ = ["load_extension", "lib_dir", "capi"]

_PKG = os.path.dirname(os.path.abspath(__file__))


def lib_dir() -> str:
    return os.path.join(_PKG, "lib")


def load_extension():
    """Import the in-tree ``QGTC`` PyTorch extension and register it as ``sys.modules['QGTC']``."""
    mod = sys.modules.get("QGTC")
    if mod is not None and getattr(mod, "__backend__", None) == "qgtc_b200":
        return mod
    import sysconfig

    import torch  # noqa: F401  libtorch must be loaded before the extension

    path = os.path.join(lib_dir(), "QGTC" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))
    if not os.path.exists(path):
        raise ImportError(
            "the QGTC extension is not built (%s missing); run `python -m qgtc_ppopp22_b200.build` "
            "or __graft_entry__.build().  There is no CPU fallback." % path)
    spec = importlib.util.spec_from_file_location("QGTC", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules["QGTC"] = mod
    return mod


class _QGTCFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Makes ``import QGTC`` work after ``import qgtc_ppopp22_b200``: the finder
    answers only for the top-level name ``QGTC`` and loads the in-tree extension lazily, so importing this
    package never needs the built library (or torch) by itself."""

    def find_spec(self, fullname, path=None, target=None):
        if fullname != "QGTC":
            return None
        return importlib.util.spec_from_loader("QGTC", self)

    def create_module(self, spec):
        return load_extension()       # raises ImportError when the extension is not built: no CPU fallback

    def exec_module(self, module):
        pass


if not any(isinstance(f, _QGTCFinder) for f in sys.meta_path):
    sys.meta_path.append(_QGTCFinder())    # appended: a real `QGTC` on sys.path (the reference's own build) still wins


def __getattr__(name):
    if name == "QGTC":
        return load_extension()
    if name == "capi":
        return importlib.import_module(__name__ + ".capi")
    raise AttributeError(name)
