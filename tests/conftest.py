import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # a gpu-marked test on a box without a GPU is a hard error of the caller's -m selection, not a skip:
    # the product has no CPU fallback. Only deselect when the user did not ask for gpu tests explicitly.
    pass
