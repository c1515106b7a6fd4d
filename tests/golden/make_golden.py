"""Provenance of tests/golden/ (run in the build container, where the
reference tree is mounted at /root/reference).

* ``bk*/zeta*/zetaw*.txt`` and ``test_data_catalogue.txt`` are the reference's
  own golden measurement files and 3-particle catalogue, copied verbatim from
  ``/root/reference/tests/test_input/{stats,ctlgs}/`` (10 significant digits,
  written by ``S/io.cpp:470``).
* The 30 000-point random catalogue is NOT copied (3 MB): it is regenerated
  from ``numpy.random.default_rng(42)`` exactly as the reference's
  ``tests/conftest.py:173-194`` does; this script checks bit-identity with the
  reference's text file.
* ``oracle_*.npz`` are outputs of the reference's C++ itself (oracle/_ref, the
  unmodified sources built against the FFTW/GSL shim) for configurations the
  reference ships no golden for (PCS, triu, off-diag, CIC/NGP, 10 bins, ...).
"""
import shutil
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
REF = Path("/root/reference/tests/test_input")


def copy_reference_goldens():
    for f in sorted((REF / "stats").glob("*.txt")):
        if f.name.startswith(("bk", "zeta")):
            shutil.copy(f, HERE / f.name)
    shutil.copy(REF / "ctlgs" / "test_data_catalogue.txt", HERE / "test_data_catalogue.txt")


def check_rand_catalogue():
    ref = np.loadtxt(REF / "ctlgs" / "test_rand_catalogue.txt").T
    gen = np.random.default_rng(seed=42)
    xyz = gen.uniform(-500., 500., size=(3, 30000))
    assert np.array_equal(xyz, ref[:3]), "default_rng(42) no longer reproduces the catalogue"
    assert np.all(ref[3] == 3. / 1000.**3)


def make_oracle_goldens():
    from oracle import ref
    ref.build()
    gen = np.random.default_rng(2024)
    L, ng = 500., 32
    pos = gen.uniform(0., L, size=(3, 2000))
    nz = np.full(2000, 2000 / L**3)
    norm = ref.norm_particles(pos, nz)
    cases = {}
    for tag, kw in {
        "bk000_pcs_triu": dict(stat="bispec", assignment="pcs", degrees=(0, 0, 0), form="full",
                               bin_range=(0.02, 0.30), num_bins=6),
        "bk202_tsc_offdiag1": dict(stat="bispec", assignment="tsc", degrees=(2, 0, 2),
                                   form="off-diag", idx_bin=1, bin_range=(0.02, 0.30), num_bins=6),
        "bk110_cic_full": dict(stat="bispec", assignment="cic", degrees=(1, 1, 0), form="full",
                               bin_range=(0.02, 0.30), num_bins=5),
        "bk000_ngp_diag10": dict(stat="bispec", assignment="ngp", degrees=(0, 0, 0), form="diag",
                                 bin_range=(0.01, 0.31), num_bins=10),
        "zeta110_tsc_diag": dict(stat="3pcf", assignment="tsc", degrees=(1, 1, 0), form="diag",
                                 bin_range=(20., 220.), num_bins=8),
        "zeta000_pcs_triu": dict(stat="3pcf", assignment="pcs", degrees=(0, 0, 0), form="full",
                                 bin_range=(20., 220.), num_bins=5),
    }.items():
        stat = kw.pop("stat")
        out = ref.threept(stat, "sim", pos, L, ng, kw.pop("assignment"), kw.pop("degrees"),
                          kw.pop("form"), kw.pop("bin_range"), kw.pop("num_bins"), norm,
                          nz_d=nz, **kw)
        out.pop("elapsed_s")
        cases[tag] = out
    flat = {f"{tag}/{k}": v for tag, out in cases.items() for k, v in out.items()}
    np.savez_compressed(HERE / "oracle_box_L500_n32_seed2024.npz", **flat)


if __name__ == "__main__":
    copy_reference_goldens()
    check_rand_catalogue()
    make_oracle_goldens()
    print("golden fixtures refreshed")
