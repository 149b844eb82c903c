import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.lib()
    # small test images: a few threads beat 64 (OpenMP fork/join per row loop)
    pyoracle.lib().orc_set_num_threads(min(8, os.cpu_count() or 1))
    return pyoracle


@pytest.fixture(scope="session")
def pair256():
    from denseflow_b200 import synth
    return synth.pair(256, 256, 0)
