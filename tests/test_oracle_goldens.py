"""CPU tests: pin the oracle (reference C++ built against the FFTW/GSL shim,
oracle/_ref) on every golden file the reference's own tests hold for the
three-point path (reference tests/test_threept.py:55-75; files in
tests/test_input/stats/), and on known-answer values of the shimmed library
calls (3-j symbols, spherical Bessel functions, reduced harmonics).

Skipped (not failed) when neither oracle/_ref/libtrv_ref.so nor /root/reference
is present.
"""
import numpy as np
import pytest

from conftest import GOLDEN, load_golden, twopt_case, check_twopt_against_golden, TWOPT_CASES, window_inputs

CASES = [((0, 0, 0), "diag", None), ((2, 0, 2), "diag", None), ((0, 0, 0), "row", 0)]


def _prep(ref, kind, data, rand, L):
    if kind == "gpp":
        return dict(catalogue_type="sim", pos_d=ref.periodise(data[:3], L), nz_d=data[3])
    los_d, los_r = ref.compute_los(data[:3]), ref.compute_los(rand[:3])
    pos_d, pos_r = ref.centre(data[:3], rand[:3], L)
    return dict(catalogue_type="survey", pos_d=pos_d, nz_d=data[3], los_d=los_d,
                pos_r=pos_r, nz_r=rand[3], los_r=los_r)


@pytest.mark.parametrize("stat,prefix", [("bispec", "bk"), ("3pcf", "zeta")])
@pytest.mark.parametrize("kind", ["gpp", "lpp"])
@pytest.mark.parametrize("degrees,form,idx_bin", CASES)
def test_oracle_reproduces_reference_goldens(oracle, stat, prefix, kind, degrees, form, idx_bin,
                                             golden_data_catalogue, golden_rand_catalogue):
    L, ng = 1000., 64
    data, rand = golden_data_catalogue, golden_rand_catalogue
    args = _prep(oracle, kind, data, rand, L)
    if kind == "gpp":
        norm = oracle.norm_particles(args["pos_d"], data[3])
    else:
        alpha = data.shape[1] / rand.shape[1]
        norm = oracle.norm_particles(args["pos_r"], rand[3], alpha=alpha)
    rng = (0.005, 0.105) if stat == "bispec" else (50., 150.)
    out = oracle.threept(stat, boxsize=L, ngrid=ng, assignment="tsc", degrees=degrees, form=form,
                         bin_range=rng, num_bins=4, norm_factor=norm, idx_bin=idx_bin or 0, **args)
    tag = "".join(map(str, degrees))
    ftag = form if form != "row" else f"row{idx_bin}"
    ext = load_golden(f"{prefix}{tag}_{ftag}_{kind}.txt")
    names = [k for k in out if k != "elapsed_s"]
    assert np.allclose(out[names[0]], ext[0])
    assert np.allclose(out[names[2]], ext[1], rtol=1.e-9, atol=0.)
    assert np.array_equal(out[names[4]], ext[2])
    assert np.allclose(out[names[1]], ext[3])
    assert np.allclose(out[names[3]], ext[4], rtol=1.e-9, atol=0.)
    assert np.array_equal(out[names[5]], ext[5])
    raw = ext[-4] + 1j * ext[-3]
    shot = ext[-2] + 1j * ext[-1]

    def rel(a, b):
        return np.max(np.abs(a - b) / np.where(np.abs(b) > 0., np.abs(b), 1.))
    # golden files carry 10 significant digits
    assert rel(out[names[6]], raw) < 2.e-9
    assert rel(out[names[7]], shot) < 2.e-9


_window_inputs = window_inputs


@pytest.mark.parametrize("degrees,form,idx_bin", CASES)
def test_oracle_reproduces_window_goldens(oracle, degrees, form, idx_bin, golden_rand_catalogue):
    """zetaw*.txt of the reference (tests/test_threept.py:278-335)."""
    L, ng = 1000., 64
    rand = golden_rand_catalogue
    pos_r, los_r = _window_inputs(oracle, rand, L)
    norm = oracle.norm_particles(pos_r, rand[3], alpha=1.)
    out = oracle.threept_window(pos_r, L, ng, "tsc", degrees, form, (50., 150.), 4, norm, los_r,
                                alpha=1., idx_bin=idx_bin or 0, nz_r=rand[3])
    tag = "".join(map(str, degrees))
    ftag = form if form != "row" else f"row{idx_bin}"
    ext = load_golden(f"zetaw{tag}_{ftag}.txt")
    assert np.allclose(out["r1_bin"], ext[0])
    assert np.allclose(out["r1_eff"], ext[1], rtol=1.e-9, atol=0.)
    assert np.array_equal(out["npairs_1"], ext[2])
    assert np.allclose(out["r2_bin"], ext[3])
    assert np.allclose(out["r2_eff"], ext[4], rtol=1.e-9, atol=0.)
    assert np.array_equal(out["npairs_2"], ext[5])
    raw = ext[-4] + 1j * ext[-3]
    shot = ext[-2] + 1j * ext[-1]

    def rel(a, b):
        return np.max(np.abs(a - b) / np.where(np.abs(b) > 0., np.abs(b), 1.))
    assert rel(out["zeta_raw"], raw) < 2.e-9
    assert rel(out["zeta_shot"], shot) < 2.e-9


# ---- two-point estimators (reference tests/test_twopt.py) -----------------

@pytest.mark.parametrize("degree", [0, 2])
@pytest.mark.parametrize("stat,kind,fname", TWOPT_CASES)
def test_oracle_reproduces_twopt_goldens(oracle, stat, kind, fname, degree,
                                         golden_data_catalogue, golden_rand_catalogue):
    """pk*, xi*, xiw* of the reference (tests/test_twopt.py)."""
    args = twopt_case(oracle, stat, kind, degree, golden_data_catalogue, golden_rand_catalogue)
    out = oracle.twopt(**args)
    check_twopt_against_golden(out, load_golden(fname.format(degree)), stat)


def test_oracle_fixture_file_is_current(oracle):
    """tests/golden/oracle_*.npz (used by the GPU tests where /root/reference is
    absent) still equals what the oracle computes from the committed seed."""
    from conftest import GOLDEN
    fix = np.load(GOLDEN / "oracle_box_L500_n32_seed2024.npz")
    gen = np.random.default_rng(2024)
    L, ng = 500., 32
    pos = gen.uniform(0., L, size=(3, 2000))
    nz = np.full(2000, 2000 / L**3)
    norm = oracle.norm_particles(pos, nz)
    out = oracle.threept("bispec", "sim", pos, L, ng, "pcs", (0, 0, 0), "full", (0.02, 0.30), 6,
                         norm, nz_d=nz)
    for k in ("bk_raw", "bk_shot", "k1_eff", "nmodes_1"):
        ref = fix[f"bk000_pcs_triu/{k}"]
        assert np.allclose(out[k], ref, rtol=1.e-12, atol=0.), k


# ---- known answers for the shimmed GSL calls (oracle/shim/trvshim_gsl.cpp) ---

def test_shim_wigner_3j_known_values(oracle):
    # (j1 j2 j3; 0 0 0) closed forms
    assert abs(oracle.w3j(0, 0, 0, 0, 0, 0) - 1.) < 1e-15
    assert abs(oracle.w3j(1, 1, 0, 0, 0, 0) + 1. / np.sqrt(3.)) < 1e-15
    assert abs(oracle.w3j(2, 2, 0, 0, 0, 0) - 1. / np.sqrt(5.)) < 1e-15
    assert abs(oracle.w3j(2, 0, 2, 0, 0, 0) - 1. / np.sqrt(5.)) < 1e-15
    assert abs(oracle.w3j(1, 1, 2, 0, 0, 0) - np.sqrt(2. / 15.)) < 1e-15
    assert abs(oracle.w3j(2, 2, 2, 0, 0, 0) + np.sqrt(2. / 35.)) < 1e-15
    assert abs(oracle.w3j(1, 1, 2, 1, -1, 0) - 1. / np.sqrt(30.)) < 1e-15
    assert oracle.w3j(1, 1, 1, 0, 0, 0) == 0.          # odd sum
    assert oracle.w3j(2, 0, 2, 1, 0, 0) == 0.          # m-sum rule
    # orthogonality: sum_m (2 0 2; m 0 -m)^2 = 1/(2*2+1) * ... = 1/5 each
    s = sum(oracle.w3j(2, 0, 2, m, 0, -m) ** 2 for m in range(-2, 3))
    assert abs(s - 1.) < 1e-14


def test_shim_spherical_bessel_spline_against_scipy(oracle):
    from scipy.special import spherical_jn
    x = np.concatenate([[0., 1.e-6, 0.049, 0.05, 0.051], np.linspace(0.1, 999.9, 4001),
                        [1000., 1000.5, 2500.]])
    for ell in (0, 1, 2, 4):
        got = oracle.sjl(ell, x)
        want = spherical_jn(ell, x)
        # step-0.05 cubic spline: interior interpolation error ~ h^4 |j''''|/384 < 2e-8;
        # the NATURAL end condition (c_0 = 0 although j_0''(0) = -1/3) costs ~1e-6 in the
        # first few intervals -- a property of gsl_interp_cspline the GPU must reproduce.
        err = np.abs(got - want)
        assert np.max(err[x > 1.]) < 2.e-8, ell
        assert np.max(err) < 3.e-6, ell
        knots = 0.05 * np.arange(0, 2000, 37)
        assert np.max(np.abs(oracle.sjl(ell, knots) - spherical_jn(ell, knots))) < 5.e-15, ell


def test_shim_reduced_harmonics_against_scipy(oracle):
    from scipy.special import sph_harm_y
    gen = np.random.default_rng(5)
    v = gen.normal(size=(200, 3))
    r = np.linalg.norm(v, axis=1)
    theta = np.arccos(v[:, 2] / r)
    phi = np.arctan2(v[:, 1], v[:, 0])
    for ell in (0, 1, 2, 3):
        for m in range(-ell, ell + 1):
            got = oracle.ylm(ell, m, v)
            # reduced harmonic: sqrt(4 pi/(2l+1)) conj(Y_lm) (S/maths.cpp:171-220)
            want = np.sqrt(4. * np.pi / (2 * ell + 1)) * np.conj(sph_harm_y(ell, m, theta, phi))
            assert np.max(np.abs(got - want)) < 1.e-13, (ell, m)
    assert np.all(oracle.ylm(2, 1, np.zeros((1, 3))) == 0.)       # |r| < eps
    assert np.all(oracle.ylm(0, 0, np.zeros((1, 3))) == 1.)


def test_oracle_binning_edges(oracle):
    e, c, w = oracle.binning("fourier", "lin", 0.005, 0.105, 4)
    assert np.allclose(e, np.linspace(0.005, 0.105, 5), rtol=1e-15)
    assert e[-1] == 0.105
    assert np.allclose(c, 0.5 * (e[1:] + e[:-1]))
    e, c, w = oracle.binning("config", "log", 10., 1000., 4)
    assert np.allclose(e, np.logspace(1, 3, 5), rtol=1e-14)


# ---------------------------------------------------------------------------
# Production-shape fixtures (tests/golden/oracle_prod_*.npz)
# ---------------------------------------------------------------------------

def test_pair_unit_assembly_equals_the_full_reference_call(oracle):
    """`oracle.ref.bispec_entries` (setup + loop body per bin pair + binned two-point
    statistics, assembled as S/threept.cpp:1981-2140 does) is what the 512^3 fixtures
    and bench.py's parity_check are made of: it must equal the reference's full
    compute_bispec_in_gpp_box call entry for entry."""
    gen = np.random.default_rng(3)
    L, ng, nb, n = 500., 32, 5, 3000
    pos = gen.uniform(0., L, (3, n))
    full = oracle.threept("bispec", "sim", pos, L, ng, "pcs", (0, 0, 0), "full", (0.02, 0.3), nb, 2.5)
    oracle.bispec_setup(pos, L, ng, "pcs", (0.02, 0.3), nb)
    pairs = [(a, b) for a in range(nb) for b in range(a, nb)]
    ent = oracle.bispec_entries(pairs, nb, n, 2.5)
    oracle.bispec_teardown()
    idx = [oracle.triu_index(a, b, nb) for a, b in pairs]
    assert idx == list(range(len(pairs)))
    for k in ("nmodes_1", "nmodes_2"):
        assert np.array_equal(ent[k], full[k][idx])
    for k in ("k1_eff", "k2_eff", "bk_raw", "bk_shot"):
        assert np.max(np.abs(ent[k] - full[k][idx]) / np.abs(full[k][idx])) < 1.e-14, k


@pytest.mark.parametrize("tag,keys", [
    ("C1", ("bk_raw", "bk_shot", "nmodes_1", "k1_eff")),
    ("C2", ("bk_raw", "bk_shot", "nmodes_1", "k1_eff", "index", "pairs")),
    ("C5proxy", ("bk_raw", "bk_shot", "nmodes_1", "k1_eff", "index", "pairs")),
    ("C4lo", ("zeta_raw", "zeta_shot", "npairs_1", "r1_eff")),
    ("C4hi", ("zeta_raw", "zeta_shot", "npairs_1", "r1_eff")),
    ("C3", ("bk_raw", "bk_shot", "nmodes_1", "k1_eff")),
])
def test_production_fixtures_are_committed_and_sane(tag, keys):
    fix = np.load(GOLDEN / f"oracle_prod_{tag}.npz")
    for k in keys:
        assert k in fix.files, (tag, k)
        assert np.all(np.isfinite(np.asarray(fix[k]).view(float) if np.iscomplexobj(fix[k])
                                  else np.asarray(fix[k], dtype=float))), (tag, k)
    assert "make_golden_production.py" in str(fix["meta"])


def test_c1_fixture_first_four_bins_are_the_reference_golden():
    """C1 with 10 bins on [0.005, 0.105] has bins of width 0.01; the reference's own
    golden (4 bins of width 0.025) covers the same modes: the first shell edge and the
    particle normalisation are common, so nmodes and the shot-noise scale tie the two."""
    fix = np.load(GOLDEN / "oracle_prod_C1.npz")
    ext = np.loadtxt(GOLDEN / "bk000_diag_gpp.txt", unpack=True)
    assert fix["bk_raw"].shape == (10,)
    # 10-bin shells nest in the 4-bin ones only at 0.005, 0.055, 0.105: modes of bins 0..4
    # equal modes of golden bins 0..1
    assert int(fix["nmodes_1"][:5].sum()) == int(ext[2][:2].sum())
    assert int(fix["nmodes_1"].sum()) == int(ext[2].sum())
    assert abs(float(fix["norm_factor"]) - 3.703703704e16) < 1.e7
