import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def golden(name):
    """Records harvested from the reference's JUnit tests by tests/golden/extract_kats.py."""
    with open(os.path.join(ROOT, "tests", "golden", name + ".json")) as f:
        return json.load(f)["methods"]


@pytest.fixture(scope="session")
def fft_kats():
    return golden("ArrayFftTest")


@pytest.fixture(scope="session")
def conv_kats():
    return golden("ConvolutionCacheTest")
