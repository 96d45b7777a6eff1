import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


def pytest_collection_modifyitems(config, items):
    # GPU tests must never silently pass on a box without a GPU: they are deselected by `-m "not gpu"`
    # here, and on the GPU box they fail loudly if CUDA or the extension is missing.
    pass
