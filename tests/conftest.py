import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds on CPU")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests need a CUDA device: on a box without one they are skipped instead of failing"""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_expected():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "expected.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_small():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "small_cases.json")) as f:
        return json.load(f)
