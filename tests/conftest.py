import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


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
