"""pytest configuration: the `gpu` marker and shared fixtures.

`-m "not gpu"` runs here (no GPU): oracle vs golden vectors, host logic, library loading / exported symbols.
`-m gpu` runs on a B200: the parity tests proper, every one of them through the C ABI.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from __graft_entry__ import load_package  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (runs on the B200 box only)")


def pytest_collection_modifyitems(config, items):
    """A hung kernel or collective must fail the test, not stall the run: every test gets a generous default timeout."""
    for item in items:
        if item.get_closest_marker("timeout") is None:
            item.add_marker(pytest.mark.timeout(1200))


@pytest.fixture(scope="session")
def bq():
    return load_package()


@pytest.fixture(scope="session")
def ctx(bq):
    c = bq.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def ref():
    """The compiled reference executor (oracle/_ref). Built here by `make -C oracle`; travels to the GPU box."""
    from oracle import ref_engine
    if not ref_engine.available():
        pytest.skip("oracle/_ref/libbosql_ref.so not built")
    return ref_engine
