"""pytest configuration: registers the ``gpu`` marker and shared fixtures."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


class Golden:
    """Read-only view of tests/golden/quanta_golden.npz (outputs of the
    unmodified reference, produced by tests/golden/make_golden.py)."""

    def __init__(self, path):
        self._z = np.load(path)
        self.manifest = json.loads(bytes(self._z["manifest"]).decode())

    def cases(self, row):
        return [c for c in self.manifest if c["row"] == row]

    def x(self, case):
        return self._z[case["x"]]

    def get(self, case, field):
        return self._z[f"{case['name']}/{field}"]


@pytest.fixture(scope="session")
def golden():
    return Golden(os.path.join(ROOT, "tests", "golden", "quanta_golden.npz"))
