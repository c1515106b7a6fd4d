"""Drop-in proof at the reference's own binding layer.

``oracle/build_refcy.py`` compiles the reference's UNMODIFIED Cython modules
(``T/_threept.pyx``, ``_particles.pyx``, ``dataobjs.pyx``, ``parameters.pyx``)
against ``triumvirate_b200/include/trv_compat`` and links them to
``libtrv_b200.so``.  These tests then drive the GPU path exactly as the
reference's Python package does (``T/threept.py:1467-1560``): a ``ParameterSet``
from the reference's test parameter file, a ``Binning``, ``_ParticleCatalogue``
objects and the ``_compute_*`` functions -- and compare with the reference's
golden files.  Skipped when the compiled modules are absent and cannot be
built (no /root/reference).
"""
import copy
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden

# Contents of the reference's tests/test_input/params/test_params.yml.
TEST_PARAMS = {
    "directories": {"catalogues": "", "measurements": ""},
    "files": {"data_catalogue": None, "rand_catalogue": None},
    "catalogue_columns": [],
    "tags": {"output": None},
    "boxsize": {"x": 1000., "y": 1000., "z": 1000.},
    "ngrid": {"x": 64, "y": 64, "z": 64},
    "expand": 1.,
    "cutoff_nyq": None,
    "alignment": "centre",
    "padscale": "box",
    "padfactor": None,
    "assignment": "tsc",
    "interlace": False,
    "catalogue_type": None,
    "statistic_type": None,
    "degrees": {"ell1": None, "ell2": None, "ELL": 0},
    "wa_orders": {"i": None, "j": None},
    "form": "diag",
    "norm_convention": "particle",
    "binning": "lin",
    "range": [0.005, 0.105],
    "num_bins": 4,
    "idx_bin": None,
    "fftw_scheme": "measure",
    "use_fftw_wisdom": False,
    "save_binned_vectors": False,
    "verbose": 20,
    "progbar": False,
}


def _load_binding(name):
    import importlib
    sys.path.insert(0, str(ROOT / "oracle"))
    import build_refcy
    if not build_refcy.build(patched=(name == "trvcy_b200")):
        pytest.skip(f"oracle/_ref/{name} not built and /root/reference absent")
    if str(ROOT / "oracle" / "_ref") not in sys.path:
        sys.path.insert(0, str(ROOT / "oracle" / "_ref"))
    pkg = importlib.import_module(name)
    for sub in ("_particles", "_threept", "_twopt", "dataobjs", "parameters"):
        importlib.import_module(f"{name}.{sub}")
    return pkg


@pytest.fixture(scope="module", params=["trvcy", "trvcy_b200"])
def trvcy(request):
    """The reference's Cython layer built against libtrv_b200.so: unmodified (`trvcy`) and
    with the Python-boundary fixes of bindings/patch_bindings.py applied (`trvcy_b200`)."""
    return _load_binding(request.param)


def _paramset(trvcy, catalogue_type, statistic_type, degrees, form, idx_bin, rng):
    """What T/threept.py:_amalgamate_parameters does to the template."""
    d = copy.deepcopy(TEST_PARAMS)
    d["catalogue_type"], d["statistic_type"] = catalogue_type, statistic_type
    d["degrees"] = dict(zip(("ell1", "ell2", "ELL"), degrees))
    d["form"], d["idx_bin"] = form, idx_bin
    d["range"] = list(rng)
    return trvcy.parameters.ParameterSet(param_dict=d)


def _catalogue(trvcy, pos, nz):
    n = pos.shape[1]
    cols = [np.ascontiguousarray(c, dtype=np.float64) for c in pos]
    return trvcy._particles._ParticleCatalogue(
        cols[0], cols[1], cols[2], np.ascontiguousarray(nz, dtype=np.float64),
        np.ones(n), np.ones(n), verbose=20)


def test_reference_cython_layer_links_against_libtrv_b200(trvcy):
    names = ["_calc_bispec_normalisation_from_mesh", "_calc_bispec_normalisation_from_particles",
             "_compute_3pcf", "_compute_3pcf_in_gpp_box", "_compute_3pcf_window",
             "_compute_bispec", "_compute_bispec_in_gpp_box"]
    for n in names:
        assert callable(getattr(trvcy._threept, n)), n
    for n in ["_calc_powspec_normalisation_from_particles", "_calc_powspec_normalisation_from_mesh",
              "_calc_powspec_normalisation_from_meshes", "_compute_powspec", "_compute_corrfunc",
              "_compute_powspec_in_gpp_box", "_compute_corrfunc_in_gpp_box",
              "_compute_corrfunc_window"]:
        assert callable(getattr(trvcy._twopt, n)), n
    with open("/proc/self/maps") as f:
        assert "triumvirate_b200/libtrv_b200.so" in f.read()


def test_reference_cython_host_objects(trvcy, golden_data_catalogue):
    """ParameterSet.validate(), Binning and the particle normalisation run on
    the host: usable without a GPU."""
    from triumvirate_b200 import catalogue as tcat
    ps = _paramset(trvcy, "sim", "bispec", (0, 0, 0), "diag", None, (0.005, 0.105))
    assert ps["npoint"] == "3pt" and ps["space"] == "fourier"
    b = trvcy.dataobjs.Binning.from_parameter_set(ps)
    assert np.allclose(b.bin_edges, np.linspace(0.005, 0.105, 5), rtol=1e-15)
    assert b.bin_edges[-1] == 0.105
    data = golden_data_catalogue
    cat = _catalogue(trvcy, tcat.periodise(data[:3], 1000.), data[3])
    norm = trvcy._threept._calc_bispec_normalisation_from_particles(cat, alpha=1.)
    assert abs(norm - 1. / (3. * 3.e-9**2)) < 1e-12 * norm      # golden header: 3.703703704e+16
    with pytest.raises(Exception):
        _paramset(trvcy, "sim", "bispec", (0, 0, 0), "nonsense-form", None, (0.005, 0.105))


CASES = [((0, 0, 0), "diag", None), ((2, 0, 2), "diag", None), ((0, 0, 0), "row", 0)]


def _check(out, ext, prefix):
    c = "k" if prefix == "bk" else "r"
    cnt = "nmodes" if prefix == "bk" else "npairs"
    stat = "bk" if prefix == "bk" else "zeta"
    assert np.allclose(out[f"{c}1_bin"], ext[0]) and np.allclose(out[f"{c}2_bin"], ext[3])
    assert np.allclose(out[f"{c}1_eff"], ext[1]) and np.allclose(out[f"{c}2_eff"], ext[4])
    assert np.array_equal(out[f"{cnt}_1"], ext[2]) and np.array_equal(out[f"{cnt}_2"], ext[5])
    raw, shot = ext[-4] + 1j * ext[-3], ext[-2] + 1j * ext[-1]

    def rel(a, b):
        return np.max(np.abs(a - b) / np.where(np.abs(b) > 0., np.abs(b), 1.))
    assert rel(out[f"{stat}_raw"], raw) < 2.e-9
    assert rel(out[f"{stat}_shot"], shot) < 2.e-9


@pytest.mark.gpu
@pytest.mark.parametrize("stat,prefix", [("bispec", "bk"), ("3pcf", "zeta")])
@pytest.mark.parametrize("degrees,form,idx_bin", CASES)
def test_goldens_through_reference_cython_box(trvcy, stat, prefix, degrees, form, idx_bin,
                                              golden_data_catalogue):
    from triumvirate_b200 import catalogue as tcat
    rng = (0.005, 0.105) if stat == "bispec" else (50., 150.)
    ps = _paramset(trvcy, "sim", stat, degrees, form, idx_bin, rng)
    binning = trvcy.dataobjs.Binning.from_parameter_set(ps)
    data = golden_data_catalogue
    cat = _catalogue(trvcy, tcat.periodise(data[:3], 1000.), data[3])
    norm = trvcy._threept._calc_bispec_normalisation_from_particles(cat, alpha=1.)
    fn = getattr(trvcy._threept, f"_compute_{stat}_in_gpp_box")
    out = fn(cat, ps, binning, norm)
    ftag = form if form != "row" else f"row{idx_bin}"
    _check(out, load_golden(f"{prefix}{''.join(map(str, degrees))}_{ftag}_gpp.txt"), prefix)


@pytest.mark.gpu
@pytest.mark.parametrize("stat,prefix", [("bispec", "bk"), ("3pcf", "zeta")])
@pytest.mark.parametrize("degrees,form,idx_bin", CASES[:2])
def test_goldens_through_reference_cython_survey(trvcy, stat, prefix, degrees, form, idx_bin,
                                                 golden_data_catalogue, golden_rand_catalogue):
    from triumvirate_b200 import catalogue as tcat
    rng = (0.005, 0.105) if stat == "bispec" else (50., 150.)
    ps = _paramset(trvcy, "survey", stat, degrees, form, idx_bin, rng)
    binning = trvcy.dataobjs.Binning.from_parameter_set(ps)
    data, rand = golden_data_catalogue, golden_rand_catalogue
    los_d, los_r = tcat.compute_los(data[:3]), tcat.compute_los(rand[:3])
    pos_d, pos_r = tcat.centre(data[:3], rand[:3], 1000.)
    cat_d, cat_r = _catalogue(trvcy, pos_d, data[3]), _catalogue(trvcy, pos_r, rand[3])
    alpha = data.shape[1] / rand.shape[1]
    norm = trvcy._threept._calc_bispec_normalisation_from_particles(cat_r, alpha=alpha)
    fn = getattr(trvcy._threept, f"_compute_{stat}")
    out = fn(cat_d, cat_r, los_d, los_r, ps, binning, norm)
    _check(out, load_golden(f"{prefix}{''.join(map(str, degrees))}_{form}_lpp.txt"), prefix)


@pytest.mark.gpu
def test_window_golden_through_reference_cython(trvcy, golden_rand_catalogue):
    from triumvirate_b200 import catalogue as tcat
    ps = _paramset(trvcy, "random", "3pcf-win", (2, 0, 2), "diag", None, (50., 150.))
    binning = trvcy.dataobjs.Binning.from_parameter_set(ps)
    rand = golden_rand_catalogue
    los_r = tcat.compute_los(rand[:3])
    pos_r, _ = tcat.centre(rand[:3], rand[:3], 1000.)
    cat_r = _catalogue(trvcy, tcat.periodise(pos_r, 1000.), rand[3])
    norm = trvcy._threept._calc_bispec_normalisation_from_particles(cat_r, alpha=1.)
    out = trvcy._threept._compute_3pcf_window(cat_r, los_r, ps, binning, alpha=1.,
                                              norm_factor=norm, wide_angle=False)
    _check(out, load_golden("zetaw202_diag.txt"), "zeta")


# ---- two-point estimators through the reference's own T/_twopt.pyx -----------

def _paramset_2pt(trvcy, catalogue_type, statistic_type, degree, rng):
    d = copy.deepcopy(TEST_PARAMS)
    d["catalogue_type"], d["statistic_type"] = catalogue_type, statistic_type
    d["degrees"] = {"ell1": None, "ell2": None, "ELL": degree}
    d["range"] = list(rng)
    return trvcy.parameters.ParameterSet(param_dict=d)


@pytest.mark.gpu
@pytest.mark.parametrize("degree", [0, 2])
@pytest.mark.parametrize("stat", ["powspec", "2pcf"])
def test_twopt_goldens_through_reference_cython(trvcy, stat, degree, golden_data_catalogue,
                                                golden_rand_catalogue):
    """pk*_gpp / xi*_gpp / pk*_lpp / xi*_lpp through `_compute_powspec*` and
    `_compute_corrfunc*` of the reference's unmodified Cython module."""
    from conftest import check_twopt_against_golden
    from triumvirate_b200 import catalogue as tcat
    rng = (0.005, 0.105) if stat == "powspec" else (50., 150.)
    data, rand = golden_data_catalogue, golden_rand_catalogue
    name = "powspec" if stat == "powspec" else "corrfunc"
    prefix = "pk" if stat == "powspec" else "xi"
    # periodic box
    ps = _paramset_2pt(trvcy, "sim", stat, degree, rng)
    binning = trvcy.dataobjs.Binning.from_parameter_set(ps)
    cat = _catalogue(trvcy, tcat.periodise(data[:3], 1000.), data[3])
    norm = trvcy._twopt._calc_powspec_normalisation_from_particles(cat, alpha=1.)
    out = getattr(trvcy._twopt, f"_compute_{name}_in_gpp_box")(cat, ps, binning, norm)
    check_twopt_against_golden(out, load_golden(f"{prefix}{degree}_gpp.txt"), stat)
    # survey
    ps = _paramset_2pt(trvcy, "survey", stat, degree, rng)
    binning = trvcy.dataobjs.Binning.from_parameter_set(ps)
    los_d, los_r = tcat.compute_los(data[:3]), tcat.compute_los(rand[:3])
    pos_d, pos_r = tcat.centre(data[:3], rand[:3], 1000.)
    cat_d, cat_r = _catalogue(trvcy, pos_d, data[3]), _catalogue(trvcy, pos_r, rand[3])
    alpha = data.shape[1] / rand.shape[1]
    norm = trvcy._twopt._calc_powspec_normalisation_from_particles(cat_r, alpha=alpha)
    out = getattr(trvcy._twopt, f"_compute_{name}")(cat_d, cat_r, los_d, los_r, ps, binning, norm)
    check_twopt_against_golden(out, load_golden(f"{prefix}{degree}_lpp.txt"), stat)


# ---- Python-boundary fixes: bindings/patch_bindings.py -----------

def test_patched_binding_turns_device_errors_into_python_exceptions(monkeypatch,
                                                                    golden_data_catalogue):
    """`except +` on the compute_* externs (T/_threept.pyx:50-95 have none): with the GPU
    path disabled the estimator throws trv::sys::DeviceError -- the patched binding raises
    RuntimeError where the unmodified one would terminate the interpreter."""
    from triumvirate_b200 import catalogue as tcat
    mod = _load_binding("trvcy_b200")
    monkeypatch.setenv("TRV_GPU_MODE", "off")
    ps = _paramset(mod, "sim", "bispec", (0, 0, 0), "diag", None, (0.005, 0.105))
    binning = mod.dataobjs.Binning.from_parameter_set(ps)
    data = golden_data_catalogue
    cat = _catalogue(mod, tcat.periodise(data[:3], 1000.), data[3])
    with pytest.raises(RuntimeError, match="CUDA device"):
        mod._threept._compute_bispec_in_gpp_box(cat, ps, binning, 1.)
    with pytest.raises(RuntimeError, match="CUDA device"):
        mod._twopt._compute_powspec_in_gpp_box(cat, ps, binning, 1.)
    with pytest.raises(RuntimeError, match="CUDA device"):
        mod._threept._calc_bispec_normalisation_from_mesh(cat, ps, 1.)


def test_patched_binding_catalogue_upload_is_one_pass():
    """Raw-array catalogue loader instead of six by-value std::vector conversions
    (T/_particles.pxd:11-14): 2e6 rows in well under a second, same contents."""
    import time
    plain, fixed = _load_binding("trvcy"), _load_binding("trvcy_b200")
    gen = np.random.default_rng(3)
    n = 2 * 10**6
    cols = [gen.uniform(0., 1000., n) for _ in range(3)] + [np.full(n, 1.e-4), np.ones(n),
                                                              gen.uniform(0.5, 1., n)]
    t0 = time.perf_counter()
    a = fixed._particles._ParticleCatalogue(*cols, verbose=20)
    t_fixed = time.perf_counter() - t0
    t0 = time.perf_counter()
    b = plain._particles._ParticleCatalogue(*cols, verbose=20)
    t_plain = time.perf_counter() - t0
    na = fixed._threept._calc_bispec_normalisation_from_particles(a, 0.5)
    nb = plain._threept._calc_bispec_normalisation_from_particles(b, 0.5)
    assert abs(na - nb) <= 1.e-12 * abs(nb)
    assert t_fixed < 0.5 and t_fixed < t_plain, (t_fixed, t_plain)
    with pytest.raises(ValueError):
        fixed._particles._ParticleCatalogue(cols[0][:5], *cols[1:], verbose=20)


@pytest.mark.gpu
def test_patched_binding_los_marshalling_is_a_memcpy(golden_data_catalogue):
    """memcpy of the (N, 3) array instead of the per-particle Python loop
    (T/_threept.pyx:138-152): a survey call with 2e6 randoms spends less time in the binding
    than the unmodified loop alone, and returns the same measurement."""
    import time
    from triumvirate_b200 import catalogue as tcat
    plain, fixed = _load_binding("trvcy"), _load_binding("trvcy_b200")
    gen = np.random.default_rng(17)
    nr = 2 * 10**6
    rand = np.vstack([gen.uniform(-500., 500., (3, nr)), np.full((1, nr), 2.e-3)])
    data = np.vstack([gen.uniform(-500., 500., (3, 20000)), np.full((1, 20000), 2.e-3)])
    los_d, los_r = tcat.compute_los(data[:3]), tcat.compute_los(rand[:3])
    pos_d, pos_r = tcat.centre(data[:3], rand[:3], 1000.)
    res, times = {}, {}
    for name, mod in (("fixed", fixed), ("plain", plain)):
        ps = _paramset(mod, "survey", "bispec", (2, 0, 2), "diag", None, (0.005, 0.105))
        binning = mod.dataobjs.Binning.from_parameter_set(ps)
        cat_d, cat_r = _catalogue(mod, pos_d, data[3]), _catalogue(mod, pos_r, rand[3])
        mod._threept._compute_bispec(cat_d, cat_r, los_d, los_r, ps, binning, 1.)   # warm-up
        t0 = time.perf_counter()
        res[name] = mod._threept._compute_bispec(cat_d, cat_r, los_d, los_r, ps, binning, 1.)
        times[name] = time.perf_counter() - t0
    assert np.allclose(res["fixed"]["bk_raw"], res["plain"]["bk_raw"], rtol=1.e-10, atol=0.)
    assert np.array_equal(res["fixed"]["nmodes_1"], res["plain"]["nmodes_1"])
    assert times["fixed"] < 0.5 * times["plain"], times
