"""CPU tests of the host-side C++ (libtrv_b200.so: parameters, binning, maths,
normalisation, catalogue helpers) against the oracle -- the reference's own C++
-- and against the reference's documented behaviour (reference
tests/test_parameters.py, tests/test_dataobjs.py)."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def core():
    from triumvirate_b200 import core
    return core


# ---- ParameterSet::validate (S/parameters.cpp:466-1270) ----------------------

@pytest.mark.parametrize("stat", ["bispec", "3pcf"])
def test_interlace_cannot_be_enabled_for_three_point(core, stat):
    """reference tests/test_parameters.py:326-333 / S/parameters.cpp:1240-1249."""
    v = core.validate("sim", stat, interlace="true")
    assert v["interlace"] == "false"


@pytest.mark.parametrize("degrees,form,shape", [
    ((0, 0, 0), "full", "triu"), ((2, 2, 0), "full", "triu"), ((2, 0, 2), "full", "full"),
    ((0, 0, 0), "diag", "diag"), ((0, 0, 0), "off-diag", "off-diag"), ((1, 1, 0), "row", "row"),
])
def test_full_form_becomes_triu_for_equal_degrees(core, degrees, form, shape):
    """S/parameters.cpp:845-849."""
    assert core.validate("sim", "bispec", form=form, degrees=degrees)["shape"] == shape


@pytest.mark.parametrize("assignment,order", [("ngp", 1), ("cic", 2), ("tsc", 3), ("pcs", 4)])
def test_assignment_order(core, assignment, order):
    assert core.validate("sim", "bispec", assignment=assignment)["assignment_order"] == order


def test_derived_members_match_the_oracle(core, oracle):
    for stat in ("bispec", "3pcf"):
        for form in ("diag", "full", "row", "off-diag"):
            for deg in ((0, 0, 0), (2, 0, 2), (1, 1, 2)):
                for cat in ("sim", "survey"):
                    a = core.validate(cat, stat, form=form, degrees=deg, interlace="true")
                    b = oracle.validate(cat, stat, form=form, degrees=deg, interlace="true")
                    assert a == b, (stat, form, deg, cat)


@pytest.mark.parametrize("bad", [dict(assignment="xyz"), dict(form="square"),
                                 dict(num_bins=1), dict(idx_bin=99, form="row"),
                                 dict(bin_range=(0.2, 0.1))])
def test_invalid_parameters_raise(core, oracle, bad):
    with pytest.raises((ValueError, RuntimeError)):
        core.validate("sim", "bispec", **bad)
    with pytest.raises(RuntimeError):
        oracle.validate("sim", "bispec", **bad)


# ---- Binning (S/dataobjs.cpp:134-249) ----------------------------------------

@pytest.mark.parametrize("space,scheme,lo,hi,nb", [
    ("fourier", "lin", 0.005, 0.105, 4), ("fourier", "lin", 0.005, 0.205, 20),
    ("fourier", "lin", 0.005, 0.405, 40), ("config", "lin", 50., 150., 4),
    ("fourier", "log", 0.005, 0.5, 12), ("config", "log", 5., 205., 20),
    ("fourier", "linpad", 0.0, 0.3, 12), ("fourier", "logpad", 0.0, 0.3, 12),
    ("config", "linpad", 0., 300., 15), ("config", "logpad", 0., 300., 15),
])
def test_binning_bit_identical_to_oracle(core, oracle, space, scheme, lo, hi, nb):
    a = core.binning(space, scheme, lo, hi, nb)
    b = oracle.binning(space, scheme, lo, hi, nb)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_linear_binning_edges(core):
    """reference tests/test_dataobjs.py: edges vs np.linspace, last edge exact."""
    e, c, w = core.binning("fourier", "lin", 0.005, 0.105, 4)
    assert np.allclose(e, np.linspace(0.005, 0.105, 5), rtol=1e-15, atol=0.)
    assert e[-1] == 0.105 and e[0] == 0.005
    assert np.allclose(w, 0.025)


# ---- maths (S/maths.cpp:167-375) -----------------------------------------------

def test_wigner_3j_and_coupling_match_oracle(core, oracle):
    for j1 in range(0, 5):
        for j2 in range(0, 5):
            for j3 in range(abs(j1 - j2), min(j1 + j2, 5) + 1):
                for m1 in range(-j1, j1 + 1):
                    for m2 in range(-j2, j2 + 1):
                        m3 = -m1 - m2
                        if abs(m3) > j3:
                            continue
                        a = core.w3j(j1, j2, j3, m1, m2, m3)
                        b = oracle.w3j(j1, j2, j3, m1, m2, m3)
                        assert abs(a - b) <= 4e-16 * max(1., abs(b)), (j1, j2, j3, m1, m2, m3)
    for (l1, l2, L) in [(0, 0, 0), (2, 0, 2), (1, 1, 0), (1, 1, 2), (2, 2, 0), (2, 2, 4)]:
        for m1 in range(-l1, l1 + 1):
            for m2 in range(-l2, l2 + 1):
                a = core.coupling(l1, l2, L, m1, m2, -m1 - m2) if abs(m1 + m2) <= L else 0.
                b = oracle.coupling(l1, l2, L, m1, m2, -m1 - m2) if abs(m1 + m2) <= L else 0.
                assert abs(a - b) <= 1e-14 * max(1., abs(b))


def test_reduced_harmonics_match_oracle(core, oracle):
    gen = np.random.default_rng(9)
    v = np.concatenate([gen.normal(size=(500, 3)),
                        [[0., 0., 0.], [0., 0., 1.], [0., 0., -2.], [1., 0., 0.], [0., -1., 0.],
                         [1e-10, 0., 0.], [-3., 0., 4.]]])
    for ell in range(0, 5):
        for m in range(-ell, ell + 1):
            a, b = core.ylm(ell, m, v), oracle.ylm(ell, m, v)
            assert np.max(np.abs(a - b)) < 5.e-15, (ell, m)


def test_bessel_spline_matches_oracle(core, oracle):
    """The spline table (not the exact j_l) is what parity at 1e-8 hinges on."""
    gen = np.random.default_rng(10)
    x = np.concatenate([[0., 0.05, 0.0499999, 999.95, 999.9999, 1000., 1000.0001, 5000.],
                        gen.uniform(0., 1100., 20000)])
    for ell in (0, 1, 2, 3, 4):
        a, b = core.sjl(ell, x), oracle.sjl(ell, x)
        assert np.max(np.abs(a - b)) < 2.e-14, ell


# ---- normalisation (S/threept.cpp:96-136) and catalogue helpers -----------------

def test_particle_normalisation_matches_oracle(core, oracle):
    gen = np.random.default_rng(12)
    n = 5000
    pos = gen.uniform(0., 100., size=(3, n))
    nz = gen.uniform(1e-4, 3e-4, n); ws = gen.uniform(0.5, 1.5, n); wc = gen.uniform(0.5, 1.5, n)
    a = core.norm_particles(pos, nz, ws=ws, wc=wc, alpha=0.1)
    b = oracle.norm_particles(pos, nz, ws=ws, wc=wc, alpha=0.1)
    assert abs(a - b) <= 1e-13 * abs(b)
    with pytest.raises(Exception):
        core.norm_particles(pos, np.zeros(n))       # all-zero nz (S/threept.cpp:120-131)


def test_golden_header_normalisation(core, golden_data_catalogue):
    """bk000_diag_gpp.txt header: 3.703703704e+16 = 1/(3 (3e-9)^2)."""
    from triumvirate_b200 import catalogue as tcat
    data = golden_data_catalogue
    pos = tcat.periodise(data[:3], 1000.)
    assert abs(core.norm_particles(pos, data[3]) / 3.703703704e+16 - 1.) < 1e-9
    # header extents [450, 550] x [456.699, 543.301] x [500, 500]
    assert np.allclose(pos.min(axis=1), [450., 456.699, 500.], atol=1e-3)
    assert np.allclose(pos.max(axis=1), [550., 543.301, 500.], atol=1e-3)


def test_catalogue_helpers_match_oracle_restatement(oracle):
    from triumvirate_b200 import catalogue as tcat
    gen = np.random.default_rng(13)
    d = gen.uniform(-300., 500., size=(3, 1000)); r = gen.uniform(-350., 520., size=(3, 4000))
    assert np.array_equal(tcat.periodise(d, 1000.), oracle.periodise(d, 1000.))
    a, b = tcat.centre(d, r, 1000.), oracle.centre(d, r, 1000.)
    assert np.allclose(a[0], b[0], rtol=0, atol=1e-12) and np.allclose(a[1], b[1], rtol=0, atol=1e-12)
    assert np.array_equal(tcat.compute_los(d), oracle.compute_los(d))
    mid = 0.5 * (a[1].min(axis=1) + a[1].max(axis=1))
    assert np.allclose(mid, 500.)


def test_reduced_harmonic_mesh_tables_equal_the_reference(oracle):
    """SphericalHarmonicCalculator::store_reduced_spherical_harmonic_in_{fourier,config}_space
    (S/maths.cpp:222-302): host tables kept for callers of the reference API."""
    from triumvirate_b200 import core
    box, ng = (300., 210., 450.), (8, 6, 10)
    for space in ("fourier", "config"):
        for ell, m in ((0, 0), (1, -1), (2, 1), (4, -3)):
            a = core.ylm_mesh(space, ell, m, box, ng)
            b = oracle.ylm_mesh(space, ell, m, box, ng)
            assert a.shape == b.shape == ng
            assert np.max(np.abs(a - b)) < 1.e-14, (space, ell, m)
