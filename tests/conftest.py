import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden", "heist_golden.npz")
GOLDEN_R2 = os.path.join(ROOT, "tests", "golden", "heist_golden_r2.npz")   # make_golden_r2.py: tie angles, big grids, trainer tapes
GOLDEN_R3 = os.path.join(ROOT, "tests", "golden", "heist_golden_r3.npz")   # make_golden_r3.py: guard patrols (strides, no-move steps, resets at every phase)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


class Golden:
    """Fixtures produced by tests/golden/make_golden.py from the unmodified reference."""

    def __init__(self, path=GOLDEN):
        self.z = np.load(path, allow_pickle=False)
        self.meta = json.loads(str(self.z["meta"]))
        self.traces = {t["name"]: t for t in self.meta["traces"]}

    def layout(self, name):
        t = self.traces[name]
        walls = [tuple(w) for w in t["walls"]]
        guards = [{**g, "patrol_path": [tuple(p) for p in g["patrol_path"]]} for g in t["guards"]]
        return walls, t["cameras"], guards, t["budget"]

    def arr(self, name, key):
        return self.z[f"{name}/{key}"]


@pytest.fixture(scope="session")
def golden():
    return Golden()


@pytest.fixture(scope="session")
def golden2():
    return Golden(GOLDEN_R2)


@pytest.fixture(scope="session")
def golden3():
    return Golden(GOLDEN_R3)
