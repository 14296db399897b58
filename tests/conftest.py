import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def rom():
    import numpy as np
    return np.fromfile(os.path.join(ROOT, "tests", "golden", "hann_rom.i16"), dtype="<i2")


@pytest.fixture(scope="session")
def gui_vectors():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "gui_vectors.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def reference_facts():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "reference_facts.json")) as f:
        return json.load(f)
