"""The stand-alone program and the file I/O behind it.

``oracle/Makefile`` compiles the reference's UNMODIFIED ``main/triumvirate.cpp`` twice:
``oracle/_ref/bin/triumvirate_ref`` against the reference's own sources (the oracle) and
``oracle/_ref/bin/triumvirate_b200`` against the headers and library of triumvirate_b200
(parameter INI reader, text catalogue reader, measurement writers, alignment helpers).
The two programs are run on the same parameter file and catalogues; their measurement
files must agree (headers byte for byte, data tables to the 10 printed digits), and the
periodic-box run must reproduce the data table of the reference's golden file.
"""
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

BIN = ROOT / "oracle" / "_ref" / "bin"
REF_BIN, B200_BIN = BIN / "triumvirate_ref", BIN / "triumvirate_b200"

TEMPLATE = """# parameter file written by tests/test_program_io.py
catalogue_dir = {cdir}
measurement_dir = {mdir}
data_catalogue_file = {data}
rand_catalogue_file = {rand}
catalogue_columns = {columns}
catalogue_dataset =
output_tag = {tag}
boxsize_x = 1000.
boxsize_y = 1000.
boxsize_z = 1000.
ngrid_x = 64
ngrid_y = 64
ngrid_z = 64
expand = 1.
cutoff_nyq =
alignment = centre
padscale = box
padfactor =
assignment = {assignment}
interlace = {interlace}
catalogue_type = {ctype}
statistic_type = {stat}
ell1 = {ell1}
ell2 = {ell2}
ELL = {ELL}
i_wa = 0
j_wa = 0
form = {form}
norm_convention = {norm}
binning = lin
bin_min = {bmin}
bin_max = {bmax}
num_bins = 4
idx_bin = {idx_bin}
fftw_scheme = measure
use_fftw_wisdom = false
save_binned_vectors = false
verbose = 20
progbar = false
"""


def _need(*paths):
    for p in paths:
        if not p.exists():
            pytest.skip(f"{p} not built (needs /root/reference at build time)")


def _write_catalogues(cdir, data, rand):
    from triumvirate_b200 import catalogue as tcat
    cdir.mkdir(parents=True, exist_ok=True)
    # positions already wrapped into the box: the program's own wrap is then the identity
    # and the periodic-box run sees what the reference's Python front end feeds its golden
    pos = tcat.periodise(data[:3], 1000.)
    with open(cdir / "data_box.txt", "w") as f:
        f.write("# x y z nz\n\n")
        for row in np.vstack([pos, data[3:4]]).T:
            f.write(" ".join(f"{v:.17g}" for v in row) + "\n")
    np.savetxt(cdir / "data_survey.txt", data.T, fmt="%.17g", header="x y z nz")
    # the randoms in two files, to exercise the `::` multi-file syntax
    half = rand.shape[1] // 2
    np.savetxt(cdir / "rand_a.txt", rand[:, :half].T, fmt="%.17g")
    np.savetxt(cdir / "rand_b.txt", rand[:, half:].T, fmt="%.17g")


def _ini(tmp, name, mdir, **kw):
    base = dict(cdir=tmp / "ctlgs", mdir=mdir, data="", rand="", columns="x,y,z,nz", tag="",
                assignment="tsc", interlace="false", ctype="sim", stat="bispec", ell1=0, ell2=0,
                ELL=0, form="diag", norm="particle", bmin=0.005, bmax=0.105, idx_bin=0)
    base.update(kw)
    path = tmp / name
    path.write_text(TEMPLATE.format(**base))
    return path


def _run(binary, ini, env=None, check=True):
    e = dict(os.environ)
    e.update(env or {})
    res = subprocess.run([str(binary), str(ini)], capture_output=True, text=True, env=e, timeout=600)
    if check:
        assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    return res


def _data_rows(path):
    return [ln for ln in Path(path).read_text().splitlines() if not ln.startswith("#")]


def _assert_same_measurement_file(got_path, ref_path, headers=True):
    """Header lines byte for byte; in the data table the coordinate and count columns as
    printed, and every complex statistic to the 10 significant digits the format carries
    (|delta| <= 2e-9 of the row's largest modulus: imaginary parts of real-valued statistics
    and shot-noise columns that cancel exactly are FFT round-off, ~1e-16 of that modulus, and
    differ between any two FFT libraries)."""
    got = Path(got_path).read_text().splitlines()
    ref = Path(ref_path).read_text().splitlines()
    assert len(got) == len(ref)
    for a, b in zip(got, ref):
        if b.startswith("#"):
            if headers:
                assert a == b
            continue
        ca, cb = a.split("\t"), b.split("\t")
        assert len(ca) == len(cb)
        ncoord = 6 if len(cb) == 10 else 3
        assert ca[:ncoord] == cb[:ncoord], (a, b)
        va = np.array(ca[ncoord:], float).view(complex)
        vb = np.array(cb[ncoord:], float).view(complex)
        # relative to the largest statistic of the row: a shot-noise column that cancels to
        # rounding (3-particle catalogue) is noise at the 1e-17 level of the raw power
        assert np.all(np.abs(va - vb) <= 2.e-9 * np.max(np.abs(vb)) + 1.e-300), (a, b)


CASES = {
    # name: (ini keywords, output file name)
    "bk000_diag_gpp": (dict(data="data_box.txt", ctype="sim", stat="bispec"), "bk000_diag"),
    "bk202_diag_lpp": (dict(data="data_survey.txt", rand="rand_a.txt::rand_b.txt", ctype="survey",
                            stat="bispec", ell1=2, ell2=0, ELL=2), "bk202_diag"),
    "zeta000_row_gpp": (dict(data="data_box.txt", ctype="sim", stat="3pcf", form="row", idx_bin=1,
                             bmin=50., bmax=150.), "zeta000_row1"),
    "pk2_lpp_mesh": (dict(data="data_survey.txt", rand="rand_a.txt::rand_b.txt", ctype="survey",
                          stat="powspec", ELL=2, norm="mesh", interlace="true", assignment="cic"),
                     "pk2"),
    "xi0_gpp": (dict(data="data_box.txt", ctype="sim", stat="2pcf", bmin=50., bmax=150.), "xi0"),
    "zetaw000_win": (dict(rand="rand_a.txt::rand_b.txt", ctype="random", stat="3pcf-win",
                          bmin=50., bmax=150.), "zetaw000_diag"),
}


def test_oracle_program_reproduces_the_golden_data_table(tmp_path, golden_data_catalogue,
                                                         golden_rand_catalogue):
    """The reference's own program (oracle build) fed the pre-wrapped catalogue writes the
    data table of tests/test_input/stats/bk000_diag_gpp.txt."""
    _need(REF_BIN)
    _write_catalogues(tmp_path / "ctlgs", golden_data_catalogue, golden_rand_catalogue[:, :600])
    kw, out = CASES["bk000_diag_gpp"]
    ini = _ini(tmp_path, "p.ini", tmp_path / "out_ref", **kw)
    _run(REF_BIN, ini)
    _assert_same_measurement_file(tmp_path / "out_ref" / out, GOLDEN / "bk000_diag_gpp.txt",
                                  headers=False)
    # (the golden's header was written by the reference's Python front end: not compared)


def test_program_help_version_and_missing_parameter_file():
    _need(B200_BIN)
    res = subprocess.run([str(B200_BIN), "--help"], capture_output=True, text=True)
    assert res.returncode == 0 and "parameter-ini-file" in res.stdout
    res = subprocess.run([str(B200_BIN), "--version"], capture_output=True, text=True)
    assert res.returncode == 0 and "GPL-3.0" in res.stdout
    res = subprocess.run([str(B200_BIN)], capture_output=True, text=True)
    assert res.returncode != 0 and "FATAL" in res.stdout


def test_parameter_file_round_trip_equals_the_reference(tmp_path, golden_data_catalogue,
                                                        golden_rand_catalogue):
    """ParameterSet::read_from_file + validate(true) + print_to_file: the `parameters_used`
    file of the two programs agrees line for line (FFTW planner lines aside: the oracle is
    the reference's CPU build).  Needs no GPU: the file is written before any estimator."""
    _need(REF_BIN, B200_BIN)
    _write_catalogues(tmp_path / "ctlgs", golden_data_catalogue, golden_rand_catalogue[:, :600])
    kw = dict(CASES["bk202_diag_lpp"][0], tag="_t")
    _run(REF_BIN, _ini(tmp_path, "a.ini", tmp_path / "out_ref", **kw),
         env={"TRV_OVERRIDE_VERBOSE": "30"})
    _run(B200_BIN, _ini(tmp_path, "b.ini", tmp_path / "out_ref2", **kw),
         env={"TRV_GPU_MODE": "off", "TRV_OVERRIDE_VERBOSE": "30"}, check=False)

    def lines(p):
        return [ln for ln in Path(p).read_text().splitlines()
                if not ln.startswith(("fftw_", "use_fftw", "measurement_dir"))]
    a = lines(tmp_path / "out_ref" / "parameters_used_t")
    b = lines(tmp_path / "out_ref2" / "parameters_used_t")
    assert a == b
    assert "verbose = 30" in a and any(ln.startswith("rand_catalogue_file = ") and "::" in ln for ln in a)


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(CASES))
def test_program_output_files_equal_the_reference(tmp_path, case,
                                                                golden_data_catalogue,
                                                                golden_rand_catalogue):
    """Same parameter file, same catalogue files, the reference's unmodified main program:
    the measurement file written through triumvirate_b200 (GPU) equals the one written
    through the reference's own C++ (CPU) -- header byte for byte, data table to the printed
    precision."""
    _need(REF_BIN, B200_BIN)
    _write_catalogues(tmp_path / "ctlgs", golden_data_catalogue, golden_rand_catalogue)
    kw, out = CASES[case]
    mdir = tmp_path / "out"
    _run(REF_BIN, _ini(tmp_path, "ref.ini", mdir, tag="_ref", **kw))
    _run(B200_BIN, _ini(tmp_path, "b200.ini", mdir, tag="_b200", **kw))
    _assert_same_measurement_file(mdir / (out + "_b200"), mdir / (out + "_ref"))
    if case == "bk000_diag_gpp":
        _assert_same_measurement_file(mdir / (out + "_b200"), GOLDEN / "bk000_diag_gpp.txt",
                                      headers=False)
