import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")
    # build the product library and the checkers once per session (cross-compiles without a GPU)
    import __graft_entry__ as G

    G.build()


@pytest.fixture(scope="session")
def golden():
    import json

    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_frames():
    import numpy as np

    return np.load(os.path.join(ROOT, "tests", "golden", "small_frames.npz"))


@pytest.fixture(scope="session")
def gpu_ctx():
    from tilecoderaytracer_b200 import api

    ctx = api.Context([0])
    yield ctx
    ctx.close()
