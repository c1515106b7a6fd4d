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


# ---- shared case builders (oracle tests on CPU, product tests on the GPU) ----

def window_inputs(mod, rand, L):
    """Python-side preparation of compute_3pcf_window (T/threept.py:1969-2010):
    LOS from the original coordinates, centre on the catalogue's own extents, then
    periodise; alpha = 1; particle normalisation with alpha = 1."""
    los_r = mod.compute_los(rand[:3])
    pos_r, _ = mod.centre(rand[:3], rand[:3], L)
    pos_r = mod.periodise(pos_r, L)
    return pos_r, los_r


def twopt_case(mod, stat, kind, degree, data, rand):
    """Inputs of one reference two-point test case (T/twopt.py:322-660, 883-1160,
    1348-1570): alignment, alpha, particle normalisation."""
    L, ng = 1000., 64
    rng = (0.005, 0.105) if stat == "powspec" else (50., 150.)
    kw = dict(boxsize=L, ngrid=ng, assignment="tsc", degree=degree, bin_range=rng, num_bins=4)
    if kind == "gpp":
        pos_d = mod.periodise(data[:3], L)
        norm = mod.norm_particles_2pt(pos_d, data[3], alpha=1.)
        return dict(stat=stat, catalogue_type="sim", pos_d=pos_d, nz_d=data[3],
                    norm_factor=norm, **kw)
    if kind == "win":
        pos_r, los_r = window_inputs(mod, rand, L)
        norm = mod.norm_particles_2pt(pos_r, rand[3], alpha=1.)
        return dict(stat="2pcf-win", catalogue_type="random", pos_r=pos_r, nz_r=rand[3],
                    los_r=los_r, alpha=1., norm_factor=norm, **kw)
    los_d, los_r = mod.compute_los(data[:3]), mod.compute_los(rand[:3])
    pos_d, pos_r = mod.centre(data[:3], rand[:3], L)
    alpha = data.shape[1] / rand.shape[1]
    norm = mod.norm_particles_2pt(pos_r, rand[3], alpha=alpha)
    return dict(stat=stat, catalogue_type="survey", pos_d=pos_d, nz_d=data[3], los_d=los_d,
                pos_r=pos_r, nz_r=rand[3], los_r=los_r, norm_factor=norm, **kw)


def check_twopt_against_golden(out, ext, stat):
    """Comparison rules of the reference's own tests (tests/test_twopt.py:36-52,
    78-92), tightened to the 10 digits the golden files carry."""
    def rel(a, b):
        scale = np.max(np.abs(b))
        return np.max(np.abs(a - b)) / (scale if scale > 0. else 1.)
    if stat == "powspec":
        assert np.allclose(out["kbin"], ext[0])
        assert np.allclose(out["keff"], ext[1], rtol=1.e-9, atol=0.)
        assert np.array_equal(out["nmodes"], ext[2])
        assert rel(out["pk_raw"], ext[3] + 1j * ext[4]) < 2.e-9
        # shot noise of the 3-particle catalogue cancels to rounding: absolute
        # tolerance as in the reference test (atol=1e-6)
        assert np.allclose(out["pk_shot"], ext[5] + 1j * ext[6], atol=1.e-6)
    else:
        assert np.allclose(out["rbin"], ext[0])
        assert np.allclose(out["reff"], ext[1], rtol=1.e-9, atol=0.)
        assert np.array_equal(out["npairs"], ext[2])
        assert rel(out["xi"], ext[3] + 1j * ext[4]) < 2.e-9


TWOPT_CASES = [("powspec", "gpp", "pk{}_gpp.txt"), ("powspec", "lpp", "pk{}_lpp.txt"),
               ("2pcf", "gpp", "xi{}_gpp.txt"), ("2pcf", "lpp", "xi{}_lpp.txt"),
               ("2pcf", "win", "xiw{}.txt")]


