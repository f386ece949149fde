import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_interp():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "interp_cases.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def golden_candles():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "candle_cases.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def golden_kats():
    import json
    with open(os.path.join(GOLDEN, "kats.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_bridge():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "bridge_cases.npz"), allow_pickle=False)
