"""TEST INFRASTRUCTURE -- generates tests/golden/oracle_prod_*.npz: outputs of the
reference's own C++ (oracle/_ref, unmodified sources + FFTW/GSL shim) at the
production shapes of BASELINE.json (512^3 meshes, 1e7 / 5e7 particles).

    python tests/golden/make_golden_production.py [C1 C2 C5proxy C4lo C4hi C3]

Needs ~25 GiB of host memory per case and minutes to tens of minutes of CPU time,
so it is run once (in the build container, 8 host cores) and its outputs are
committed; the GPU tests (tests/test_gpu_production.py) regenerate the seeded
inputs from tests/golden/production_cases.py and compare at 1e-8.

The box-bispectrum cases use the reference's loop body per (k1, k2) bin pair
(oracle/ref_driver.cpp: trvref_bispec_setup / _pair / _twopt -- the reference's own
MeshField / FieldStats calls in the order of S/threept.cpp:1543-1672, 1900-1968,
1981-2140), which `oracle.ref.bispec_entries` assembles into data-vector entries
exactly as compute_bispec_in_gpp_box does; tests/test_oracle_goldens.py pins that
assembly against the full reference call on a small mesh.
"""
import json
import os
import platform
import sys
import time
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(HERE))

import production_cases as pc  # noqa: E402


def meta(t0):
    return json.dumps({"host": platform.node(), "cores": os.cpu_count(),
                       "seconds": round(time.time() - t0, 1),
                       "generator": "tests/golden/make_golden_production.py"})


def box_pairs(tag):
    from oracle import ref
    cs = pc.BOX_PAIR_CASES[tag]
    t0 = time.time()
    pos = pc.uniform_box(cs["n"], cs["L"], cs["seed"])
    t_setup = ref.bispec_setup(pos, cs["L"], cs["ngrid"], cs["assignment"], cs["bin_range"],
                               cs["num_bins"])
    ent = ref.bispec_entries(cs["pairs"], cs["num_bins"], cs["n"], 1.)
    pk, sn = ref.bispec_twopt(cs["num_bins"])
    ref.bispec_teardown()
    idx = np.array([ref.triu_index(a, b, cs["num_bins"]) for a, b in cs["pairs"]])
    np.savez_compressed(HERE / f"oracle_prod_{tag}.npz", pairs=np.array(cs["pairs"]), index=idx,
                        setup_s=t_setup, pk=pk, sn=sn, meta=meta(t0), **ent)
    print(tag, "done", meta(t0), flush=True)


def full_call(tag, kw, norm_from=None):
    from oracle import ref
    t0 = time.time()
    kw = dict(kw)
    stat = kw.pop("stat")
    ctype = kw.pop("catalogue_type")
    if norm_from == "data":
        norm = ref.norm_particles(kw["pos_d"], kw["nz_d"])
    elif norm_from == "rand":
        alpha = kw["pos_d"].shape[1] / kw["pos_r"].shape[1]
        norm = ref.norm_particles(kw["pos_r"], kw["nz_r"], wc=kw["wc_r"], alpha=alpha)
    else:
        norm = 1.
    out = ref.threept(stat, ctype, norm_factor=norm, **kw)
    np.savez_compressed(HERE / f"oracle_prod_{tag}.npz", norm_factor=norm, meta=meta(t0), **out)
    print(tag, "done", meta(t0), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["C1", "C2", "C5proxy", "C4lo", "C4hi", "C3"]
    for tag in which:
        if tag in pc.BOX_PAIR_CASES:
            box_pairs(tag)
        elif tag == "C1":
            full_call("C1", pc.c1_inputs(HERE), norm_from="data")
        elif tag in ("C4lo", "C4hi"):
            full_call(tag, pc.c4_inputs(tag[2:]))
        elif tag == "C3":
            full_call("C3", pc.c3_inputs(), norm_from="rand")
        else:
            raise SystemExit(f"unknown case {tag}")
