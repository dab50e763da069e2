import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


_HAS_GPU = None


def _has_gpu() -> bool:
    """True when the driver sees at least one CUDA device (no torch import, no context creation)."""
    global _HAS_GPU
    if _HAS_GPU is None:
        import ctypes
        try:
            cu = ctypes.CDLL("libcuda.so.1")
            n = ctypes.c_int(0)
            _HAS_GPU = cu.cuInit(0) == 0 and cu.cuDeviceGetCount(ctypes.byref(n)) == 0 and n.value > 0
        except OSError:
            _HAS_GPU = False
    return _HAS_GPU


def pytest_runtest_setup(item):
    # `-m gpu` on a box without a device: skip (the product itself still fails loudly, tests/test_abi.py)
    if item.get_closest_marker("gpu") is not None and not _has_gpu():
        pytest.skip("no CUDA device visible")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def pf13(oracle):
    """The reference-built 4^13 MPHF (oracle/_ref/data/all_13mers.pf, md5 pinned in SURVEY 8(c))."""
    if not os.path.exists(oracle.PF13_PATH):
        pytest.skip("oracle/_ref/data/all_13mers.pf missing (built by __graft_entry__.build() "
                    "in the build container)")
    return oracle.PF13_PATH
