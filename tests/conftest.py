import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    O.set_mean_mode(0)
    return O


@pytest.fixture(scope="session")
def gpu_ctx():
    from spagxecct_b200 import Context
    ctx = Context(0)  # raises loudly if the CUDA library or the device is missing
    yield ctx
    ctx.close()
