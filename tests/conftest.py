import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, f"{name}.npz")))


def golden_segments(g):
    """yields (s, xyz1, xyz2, lm1, lm2) for every segment pair of a golden case"""
    for s in range(len(g["off1"]) - 1):
        a, b = g["off1"][s], g["off1"][s + 1]
        c, d = g["off2"][s], g["off2"][s + 1]
        e, f = g["lm_off"][s], g["lm_off"][s + 1]
        yield s, g["xyz1"][a:b], g["xyz2"][c:d], g["lm1"][e:f], g["lm2"][e:f]


@pytest.fixture(scope="session")
def pkg():
    return importlib.import_module("4d_plant_analysis_b200")


@pytest.fixture(scope="session")
def syn():
    return importlib.import_module("4d_plant_analysis_b200.synthetic")
