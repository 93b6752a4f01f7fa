import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def comparable(a, b, margin):
    """ftnunit assert_comparable_real (src/libs/ftnunit.f90:353-364)."""
    return abs(a - b) <= 0.5 * margin * (abs(a) + abs(b))


# Unit-test tolerance of the reference when MATRIX_PRECISION = 4 (src/global_typedefs.F90:55).
TOL = 1.0e-6


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc
    orc.lib()
    return orc
