import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_outputs.npz"))


@pytest.fixture(scope="session")
def smpl_model():
    import whmr_b200.synthetic as syn
    return syn.make_smpl_model(seed=0, weights="random")


@pytest.fixture(scope="session")
def smpl_model_skeleton():
    import whmr_b200.synthetic as syn
    return syn.make_smpl_model(seed=3, weights="skeleton")
