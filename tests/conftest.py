import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLD = os.path.join(ROOT, "tests", "golden")
WEIGHTS = os.path.join(GOLD, "superpoint_v1.spw")
GOLDEN_CASES = ["g120x160", "g240x320_ragged", "g480x640", "g480x752", "g480x752_cap"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def weights():
    from oracle import weights as OW
    return OW.read_spw(WEIGHTS)


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = dict(np.load(os.path.join(GOLD, name + ".npz")))
        return cache[name]
    return load
