"""pytest configuration: the ``gpu`` marker and golden-fixture loaders."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Cases:
    """tests/golden/synthetic_cases.npz as {tag: {field: array}}."""

    def __init__(self):
        z = np.load(os.path.join(GOLDEN, "synthetic_cases.npz"))
        self.cases = {}
        for key in z.files:
            tag, field = key.split("/")
            self.cases.setdefault(tag, {})[field] = z[key]

    def __getitem__(self, tag):
        return self.cases[tag]

    def tags(self):
        return sorted(self.cases)


_cases = None


def load_cases():
    global _cases
    if _cases is None:
        _cases = Cases()
    return _cases


def load_model(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


CASE_TAGS = load_cases().tags()


@pytest.fixture(scope="session")
def cases():
    return load_cases()


@pytest.fixture(scope="session")
def b21():
    return load_model("b21b0214")


@pytest.fixture(scope="session")
def nbmodel():
    return load_model("dfm_notebook")
