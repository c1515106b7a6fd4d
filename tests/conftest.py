"""Shared fixtures.  Tests marked ``gpu`` need a CUDA device (run on the B200
box with ``pytest -m gpu``); everything else runs on CPU only."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")


@pytest.fixture(scope="session")
def oracle():
    """The reference's own C++ built against the shim (oracle/_ref)."""
    from oracle import ref
    if not ref.available():
        if not ref.build():
            pytest.skip("oracle/_ref/libtrv_ref.so not built and /root/reference absent")
    return ref


@pytest.fixture(scope="session")
def golden_data_catalogue():
    """3 particles on an equilateral triangle (reference tests/conftest.py:149-170)."""
    return np.loadtxt(GOLDEN / "test_data_catalogue.txt").T   # rows: x y z nz


@pytest.fixture(scope="session")
def golden_rand_catalogue():
    """30 000 uniform points, default_rng(42) (reference tests/conftest.py:173-194);
    verified bit-identical to tests/test_input/ctlgs/test_rand_catalogue.txt by
    tests/golden/make_golden.py."""
    gen = np.random.default_rng(seed=42)
    xyz = gen.uniform(-500., 500., size=(3, 30000))
    nz = np.full(30000, 3. / 1000.**3)
    return np.vstack([xyz, nz])


def load_golden(name):
    return np.loadtxt(GOLDEN / name, unpack=True)
