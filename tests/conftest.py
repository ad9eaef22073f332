import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle_port_lib():
    """Compiles oracle/liboracle.so (plain C restatement) on demand."""
    from oracle import binding
    binding.build(ref=False, port=True)
    return binding.PORT_SO


@pytest.fixture(scope="session")
def product_lib():
    """Builds the CUDA library in-tree when it is missing or stale (nvcc cross-compiles on CPU)."""
    from juicer_b200 import build
    return build.build()
