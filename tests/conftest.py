"""pytest configuration: markers, repo-root import path, shared fixtures."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "ref: needs oracle/_ref/libref_join.so (the reference's headers, compiled)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.build(ref=True)
    return pyoracle.Oracle()


@pytest.fixture(scope="session")
def ref(oracle):
    """The reference's own code (oracle/_ref).  Skips when it was never built (it needs the
    reference checkout at build time; the prebuilt .so travels to the GPU box)."""
    from oracle import pyoracle
    if not pyoracle.Ref.available():
        pytest.skip("oracle/_ref/libref_join.so not built (no reference checkout here)")
    return pyoracle.Ref()


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
