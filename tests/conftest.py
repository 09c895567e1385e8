import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def kat():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "kat.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def ref_vectors():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_vectors.npz"))


@pytest.fixture(scope="session")
def port():
    from oracle import loader
    loader.build()
    return loader.port()


@pytest.fixture(scope="session")
def checker():
    """the CPU checker of record: the reference library when it is present, else the pinned port"""
    from oracle import loader
    loader.build()
    return loader.ref() if loader.have_ref() else loader.port()
