import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # build what is missing (nvcc cross-compiles without a GPU; prebuilt files are kept)
    import subprocess
    from biostrings_b200 import build as b
    if not os.path.exists(b.SO):
        b.build()
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "all"])


def _have_gpu():
    try:
        from biostrings_b200 import _lib
        return _lib.lib().bsgpu_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def backends():
    """CPU oracles available here: always the C port, plus the compiled reference."""
    from oracle import ref
    out = ["port"]
    if ref.available():
        out.append("ref")
    return out
