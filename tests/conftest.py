import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # GPU tests are selected with -m gpu; without a device they are skipped rather than failed
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    import bitplane_oracle
    return bitplane_oracle


@pytest.fixture(scope="session")
def coracle():
    import c_oracle
    c_oracle.lib()
    return c_oracle


def rand_features(rng, shape, nbits, wild=True):
    """fp32 matrix exercising the quantizer: in-range integers, halves, and (wild) out-of-range values."""
    hi = 1 << nbits
    x = rng.integers(0, hi, size=shape).astype(np.float32)
    if wild:
        m = rng.random(shape)
        x = np.where(m < 0.05, x + 0.5, x)                                  # ties -> round-half-even
        x = np.where((m >= 0.05) & (m < 0.08), -rng.random(shape) * 3, x)    # negatives -> 1
        x = np.where((m >= 0.08) & (m < 0.11), hi + rng.random(shape) * 5, x)  # above range -> 2^b-1 (or 2^b -> 0 bits)
        x = np.where((m >= 0.11) & (m < 0.12), np.float32(hi), x)            # exactly 2^b -> packs as 0
    return x.astype(np.float32)


def rand_adj(rng, M, K, density=0.5, zero_blocks=False):
    a = (rng.random((M, K)) < density).astype(np.float32)
    if zero_blocks and M >= 16 and K >= 256:
        a[: M // 2, K // 3: 2 * K // 3] = 0       # whole zero tiles for the zero-jump path
        a[M // 4: M // 2, :] = 0
    return a
