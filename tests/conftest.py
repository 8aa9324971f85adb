import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    # the C-ABI library and the oracle are built in-tree; build them if a fresh checkout has not run build() yet
    try:
        from ubu_b200 import build as B

        if not os.path.exists(B.LIB):
            B.build()
        if not os.path.exists(os.path.join(ROOT, "host", "libubu_host.so")):
            import __graft_entry__ as G

            G.build()
    except Exception as e:  # the ABI tests will then fail loudly with the loader's message
        sys.stderr.write(f"[conftest] could not build libubu_sw.so: {e}\n")


def _has_gpu() -> bool:
    try:
        from ubu_b200 import _native as N

        return N.lib().ubu_sw_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
