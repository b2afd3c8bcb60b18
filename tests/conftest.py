import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "ref: needs the compiled reference oracle (oracle/_ref)")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The shared library must exist; building is a no-op when it is up to date."""
    from biorseo_b200 import build
    try:
        build.build()
    except Exception as exc:   # no nvcc on this box: fall back to the prebuilt .so
        if not os.path.exists(build.LIB_PATH):
            raise RuntimeError(f"cannot build libbiorseo_b200.so: {exc}")
    yield
