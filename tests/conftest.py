import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real sm_100 (B200) GPU; run with `-m gpu`")


@pytest.fixture(scope="session")
def golden():
    import torch
    path = os.path.join(ROOT, "tests", "golden", "reference_golden.pt")
    return torch.load(path, weights_only=False)


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    import linkless_link_prediction_b200._native as N
    N.require_gpu()  # raises (does not skip) when the library is missing or the GPU is not sm_100
    return torch.device("cuda:0")
