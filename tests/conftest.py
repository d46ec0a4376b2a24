import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "hevc-image-encoder-lite_b200"))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long CPU test, skipped unless HEVCE_SLOW=1")


def pytest_collection_modifyitems(config, items):
    if os.environ.get("HEVCE_SLOW") == "1":
        return
    skip = pytest.mark.skip(reason="set HEVCE_SLOW=1 to run")
    for it in items:
        if "slow" in it.keywords:
            it.add_marker(skip)
