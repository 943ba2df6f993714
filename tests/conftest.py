import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs the reference tree at /root/reference")


@pytest.fixture(scope="session")
def golden_fabrik():
    return dict(np.load(os.path.join(GOLDEN_DIR, "fabrik_reference.npz")))


@pytest.fixture(scope="session")
def golden_fk():
    return dict(np.load(os.path.join(GOLDEN_DIR, "fk_reference.npz")))


@pytest.fixture(scope="session")
def golden_generators():
    return dict(np.load(os.path.join(GOLDEN_DIR, "generators_reference.npz")))


FABRIK_SETS = ["workspace", "interior", "spring50", "spring500", "circle200", "normal05",
               "boundary", "edge"]

# golden vectors held by the reference's own tests (values, cited; not code)
REF_INVERSE_UNIT_POINTS = [[1.0, 2.1, 3.0], [1.567, 2.22, -2.123], [1.02, 3.33, 4.99]]
REF_INVERSE_UNIT_FABRIK = [  # reference tests/inverse_unit.py:23-26
    [1.1263771168937977, 1.95663870779144, -1.581170282866297, -1.2914981807424972],
    [0.9561510602151175, -0.1334947854494175, -1.441291844752837, 0.38467252287989595],
    [1.2735640189772053, 1.4953811089376177, -0.6880936114216039, -1.03376967052818]]
REF_INVERSE_UNIT_OUT_OF_REACH = [[1.0, 2.1, 3.0], [1.567, 2.22, -3.123], [1.02, 3.33, 4.99]]  # :33
REF_FABRIK_UNIT_EFFECTOR = [1.0000000035582093, 2.0000000071394073, 2.999999989135574]  # fabrik_unit.py:28
REF_FABRIK_UNIT_CHAIN = [[0.0, 0.0, 2.0],  # fabrik_unit.py:25-28 (only [3] is asserted upstream)
                         [-0.3524468346566213, -0.7083832214867447, 3.836838163871981],
                         [0.472013795834034, 0.9406164237214534, 4.612121878955813],
                         [1.0000000035582093, 2.0000000071394073, 2.999999989135574]]
REF_FORWARD_UNIT_ANGLES = [  # reference tests/forward_unit.py:22-24
    [1.1489898108341745, 1.6426609377538854, -1.2027772444264693, -1.0663073873609727],
    [-1.5672140776862065, 0.2433182869870163, -1.3760689820099818, 0.0465569704233757],
    [-0.24795388218721454, 0.9644220067435634, -1.5389903144536021, -0.3143083371860276]]
REF_FORWARD_UNIT_POINTS = [[1.34542, 2.99821, 3.67401], [0.01333, -3.72111, -1.09902],
                           [3.95444, -1.00112, 1.00378]]  # forward_unit.py:18-20, decimal=4
REF_POINT_UNIT_DISTANCE = 3.831649  # point_unit.py:22, between (0,0,0) and (-2.22, 3.123, 0.002)
