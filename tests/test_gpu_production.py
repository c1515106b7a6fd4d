"""GPU parity at the PRODUCTION shapes of BASELINE.json (configs C1-C5): the CUDA
path through the C ABI against outputs of the reference's own C++ (oracle/_ref)
committed as ``tests/golden/oracle_prod_*.npz`` (generator:
``tests/golden/make_golden_production.py``; seeded inputs:
``tests/golden/production_cases.py``), plus one live oracle run at 512^3 and the
1-versus-8-share equality of the full 1024^3 data vector.

These exercise what only exists at scale: the sub-grid chosen from the cost table
(n_s = 144 for C2, larger for the 40-bin proxy of C5), the half-spectrum + real
pair-reduction path at 512^3, the radial histogram with q up to 3 (n/2)^2, the
fine-bin rule over 1e8 modes, the tile sort over 2^18 tiles, the staged upload of
5e7 randoms and 64-bit indexing at 1024^3.

Tolerance: |delta| <= 1e-8 |ref| per complex entry (BASELINE.json), exact integer
columns, 1e-12 relative effective coordinates.
"""
import sys

import numpy as np
import pytest

from conftest import GOLDEN

sys.path.insert(0, str(GOLDEN))
import production_cases as pc  # noqa: E402

pytestmark = pytest.mark.gpu

RTOL = 1.e-8


@pytest.fixture(scope="module")
def core():
    from triumvirate_b200 import core
    if core.gpu_count() < 1:
        pytest.fail("no CUDA device visible: GPU tests cannot run (no CPU fallback)")
    return core


@pytest.fixture(scope="module")
def c2_catalogue():
    return pc.uniform_box(10**7, 1000., 42)


def _fixture(tag):
    path = GOLDEN / f"oracle_prod_{tag}.npz"
    assert path.exists(), f"{path} missing: run tests/golden/make_golden_production.py {tag}"
    return np.load(path)


def _rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b) / np.abs(b)))


def _check_full(out, fix, names):
    """Full result struct against a fixture of the reference's full call."""
    c1b, c2b, c1e, c2e, n1, n2, raw, shot = names
    assert np.array_equal(out[n1], fix[n1]) and np.array_equal(out[n2], fix[n2])
    for k in (c1b, c2b):
        assert np.allclose(out[k], fix[k], rtol=1.e-14, atol=0.)
    for k in (c1e, c2e):
        assert _rel(out[k], fix[k]) <= 1.e-12, f"{k}: {_rel(out[k], fix[k]):.2e}"
    for k in (raw, shot):
        ref = fix[k]
        scale = np.maximum(np.abs(ref), 1.e-8 * np.abs(ref).max())
        err = float(np.max(np.abs(out[k] - ref) / scale))
        assert err <= RTOL, f"{k}: max rel err {err:.3e}"


BK = ("k1_bin", "k2_bin", "k1_eff", "k2_eff", "nmodes_1", "nmodes_2", "bk_raw", "bk_shot")
ZETA = ("r1_bin", "r2_bin", "r1_eff", "r2_eff", "npairs_1", "npairs_2", "zeta_raw", "zeta_shot")


# ---------------------------------------------------------------------------
# C1 exactly as BASELINE.json states it: 64^3, TSC, 10 diagonal bins
# ---------------------------------------------------------------------------

def test_c1_ten_bins_against_oracle_fixture(core):
    fix = _fixture("C1")
    kw = pc.c1_inputs(GOLDEN)
    stat = kw.pop("stat")
    ctype = kw.pop("catalogue_type")
    norm = core.norm_particles(kw["pos_d"], kw["nz_d"])
    assert abs(norm - float(fix["norm_factor"])) <= 1.e-14 * abs(norm)
    out = core.threept(stat, ctype, norm_factor=norm, **kw)
    _check_full(out, fix, BK)


# ---------------------------------------------------------------------------
# C2 and the 512^3 proxy of C5: entries of the triu data vector against the
# reference loop body per bin pair
# ---------------------------------------------------------------------------

@pytest.mark.parametrize("tag", ["C2", "C5proxy"])
@pytest.mark.parametrize("mode", ["throughput", "deterministic", "host-arrays"])
def test_box_pairs_at_512_against_oracle_fixture(core, c2_catalogue, tag, mode):
    fix = _fixture(tag)
    cs = pc.BOX_PAIR_CASES[tag]
    pos = c2_catalogue
    assert cs["n"] == pos.shape[1] and cs["seed"] == 42
    kw = dict(boxsize=cs["L"], ngrid=cs["ngrid"], assignment=cs["assignment"], degrees=(0, 0, 0),
              form="full", bin_range=cs["bin_range"], num_bins=cs["num_bins"], norm_factor=1.)
    if mode == "host-arrays":
        # the streamed upload + first assignment of trv_threept_box_arrays (bench.py's e2e leg)
        x, y, z = (np.ascontiguousarray(pos[i]) for i in range(3))
        out = core.threept_box_arrays("bispec", pos.shape[1], x.ctypes.data, y.ctypes.data,
                                      z.ctypes.data, False, **kw)
    else:
        out = core.threept("bispec", "sim", pos_d=pos, deterministic=(mode == "deterministic"), **kw)
    nb = cs["num_bins"]
    assert len(out["bk_raw"]) == nb * (nb + 1) // 2
    idx = fix["index"]
    assert np.array_equal(out["nmodes_1"][idx], fix["nmodes_1"])
    assert np.array_equal(out["nmodes_2"][idx], fix["nmodes_2"])
    assert _rel(out["k1_eff"][idx], fix["k1_eff"]) <= 1.e-12
    assert _rel(out["k2_eff"][idx], fix["k2_eff"]) <= 1.e-12
    e_raw = _rel(out["bk_raw"][idx], fix["bk_raw"])
    e_shot = _rel(out["bk_shot"][idx], fix["bk_shot"])
    assert e_raw <= RTOL, f"{tag} bk_raw: {e_raw:.3e}"
    assert e_shot <= RTOL, f"{tag} bk_shot: {e_shot:.3e}"


def test_c2_live_oracle_pairs_at_512(core, c2_catalogue, oracle):
    """The oracle run LIVE on the GPU box's host cores at 512^3 (setup + two pair
    units of the reference loop) against the GPU result of the same catalogue."""
    cs = pc.BOX_PAIR_CASES["C2"]
    pos = c2_catalogue
    out = core.threept("bispec", "sim", pos_d=pos, boxsize=cs["L"], ngrid=cs["ngrid"],
                       assignment=cs["assignment"], degrees=(0, 0, 0), form="full",
                       bin_range=cs["bin_range"], num_bins=cs["num_bins"], norm_factor=1.)
    nb = cs["num_bins"]
    pairs = [(2, 15), (12, 12)]
    oracle.bispec_setup(pos, cs["L"], cs["ngrid"], cs["assignment"], cs["bin_range"], nb)
    ent = oracle.bispec_entries(pairs, nb, cs["n"], 1.)
    oracle.bispec_teardown()
    idx = np.array([oracle.triu_index(a, b, nb) for a, b in pairs])
    assert np.array_equal(out["nmodes_1"][idx], ent["nmodes_1"])
    assert np.array_equal(out["nmodes_2"][idx], ent["nmodes_2"])
    assert _rel(out["k1_eff"][idx], ent["k1_eff"]) <= 1.e-12
    assert _rel(out["bk_raw"][idx], ent["bk_raw"]) <= RTOL
    assert _rel(out["bk_shot"][idx], ent["bk_shot"]) <= RTOL


# ---------------------------------------------------------------------------
# C4: zeta_110 at 512^3 (spherical-Bessel-weighted transforms on the full grid)
# ---------------------------------------------------------------------------

@pytest.mark.parametrize("which", ["lo", "hi"])
def test_c4_3pcf_at_512_against_oracle_fixture(core, c2_catalogue, which):
    fix = _fixture(f"C4{which}")
    kw = pc.c4_inputs(which)
    stat = kw.pop("stat")
    ctype = kw.pop("catalogue_type")
    kw["pos_d"] = c2_catalogue
    out = core.threept(stat, ctype, norm_factor=float(fix["norm_factor"]), **kw)
    _check_full(out, fix, ZETA)


# ---------------------------------------------------------------------------
# C3: survey B_202 with 5e7 randoms at 512^3
# ---------------------------------------------------------------------------

def test_c3_survey_at_512_against_oracle_fixture(core):
    fix = _fixture("C3")
    kw = pc.c3_inputs()
    stat = kw.pop("stat")
    ctype = kw.pop("catalogue_type")
    alpha = kw["pos_d"].shape[1] / kw["pos_r"].shape[1]
    norm = core.norm_particles(kw["pos_r"], kw["nz_r"], wc=kw["wc_r"], alpha=alpha)
    # 5e7-term sums: the OpenMP reduction order (thread count) moves the last digits
    assert abs(norm - float(fix["norm_factor"])) <= 1.e-9 * abs(norm)
    out = core.threept(stat, ctype, norm_factor=float(fix["norm_factor"]), **kw)
    _check_full(out, fix, BK)


# ---------------------------------------------------------------------------
# C5 at 1024^3: the eight shares of the 8-GPU partition, each computed alone on this
# GPU, sum to the single-share result (1 vs 8 devices) -- bit for bit in
# deterministic mode
# ---------------------------------------------------------------------------

def test_c5_1024_eight_shares_equal_one(core):
    import torch
    n = 10**8
    free, _ = torch.cuda.mem_get_info()
    if free < 100 * 2**30:
        pytest.skip("needs ~90 GiB of free HBM")
    pos = pc.uniform_box(n, 2000., 42)
    d = torch.from_numpy(pos).to("cuda:0")
    del pos
    torch.cuda.synchronize()
    kw = dict(boxsize=2000., ngrid=1024, assignment="pcs", degrees=(0, 0, 0), form="full",
              bin_range=(0.005, 0.405), num_bins=40, norm_factor=1., deterministic=True)

    def run(rank, world):
        return core.threept_box_arrays("bispec", n, d[0].data_ptr(), d[1].data_ptr(),
                                       d[2].data_ptr(), True, part_rank=rank, part_count=world, **kw)

    one = run(0, 1)
    assert len(one["bk_raw"]) == 820 and np.all(np.isfinite(one["bk_raw"].view(float)))
    raw = np.zeros(820, dtype=complex)
    shot = np.zeros(820, dtype=complex)
    for r in range(8):
        part = run(r, 8)
        raw += part["bk_raw"]
        shot += part["bk_shot"]
        assert np.array_equal(part["nmodes_1"], one["nmodes_1"])
        assert np.array_equal(part["k1_eff"], one["k1_eff"])
    assert raw.tobytes() == one["bk_raw"].tobytes()
    assert shot.tobytes() == one["bk_shot"].tobytes()
    # the outermost shell [0.395, 0.405) of the 2^30-cell mesh holds ~6.5e5 modes
    assert int(one["nmodes_2"][-1]) > 600000
    del d
    torch.cuda.empty_cache()
    core.release_contexts()
