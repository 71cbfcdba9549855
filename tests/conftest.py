import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden:
    """Golden vectors produced from the unmodified reference by tests/golden/make_golden.py."""

    def __init__(self, path):
        self._z = np.load(path)
        self.meta = {m["name"]: m for m in json.loads(bytes(self._z["__meta__"]).decode())}

    def names(self):
        return list(self.meta)

    def case(self, name):
        import torch
        pre = name + "/"
        d = {k[len(pre):]: torch.from_numpy(self._z[k]) for k in self._z.files if k.startswith(pre)}
        return d, self.meta[name]


GOLDEN_PATH = os.path.join(ROOT, "tests", "golden", "gat_golden.npz")


@pytest.fixture(scope="session")
def golden():
    return Golden(GOLDEN_PATH)


def golden_case_names():
    z = np.load(GOLDEN_PATH)
    return [m["name"] for m in json.loads(bytes(z["__meta__"]).decode())]
