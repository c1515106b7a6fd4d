"""Definitions of the production-shape parity cases (BASELINE.json configs C1-C5,
SURVEY.md section 8d): seeded catalogues and estimator arguments, shared by the
fixture generator (``make_golden_production.py``, runs the oracle) and by the GPU
tests that load the fixtures (``tests/test_gpu_production.py``).

Nothing here touches the oracle or the product; it is inputs only.
"""
import numpy as np


def uniform_box(n, L, seed=42):
    """C2 / C4 / C5 catalogue: ``default_rng(seed).uniform(0, L, (3, n))``."""
    return np.random.default_rng(seed).uniform(0., L, size=(3, n))


def shell_octant(gen, n):
    """C3 catalogues: uniform in the octant shell 500 < r < 1500."""
    r = (gen.uniform(500.**3, 1500.**3, n)) ** (1. / 3.)
    mu = gen.uniform(0., 1., n)
    ph = gen.uniform(0., np.pi / 2, n)
    s = np.sqrt(1 - mu**2)
    return np.array([r * s * np.cos(ph), r * s * np.sin(ph), r * mu])


def periodise(pos, boxsize):
    """T/catalogue.py:668-674."""
    pos = np.array(pos, dtype=np.float64, copy=True)
    for ax in range(3):
        lo, hi = pos[ax].min(), pos[ax].max()
        pos[ax] = (pos[ax] + boxsize / 2. - (hi + lo) / 2.) % boxsize
    return pos


def compute_los(pos):
    """T/catalogue.py:437-458; (N, 3)."""
    norm = np.sqrt(pos[0]**2 + pos[1]**2 + pos[2]**2)
    norm[norm == 0.] = 1.
    return np.ascontiguousarray(np.transpose([pos[0] / norm, pos[1] / norm, pos[2] / norm]))


def centre(pos, pos_ref, boxsize):
    """T/catalogue.py:535-545 with the randoms as reference catalogue."""
    origin = np.array([np.mean([pos_ref[ax].min(), pos_ref[ax].max()]) - boxsize / 2.
                       for ax in range(3)])
    return pos - origin[:, None], pos_ref - origin[:, None]


# -- box bispectrum B_000, pair units of the reference loop -------------------

BOX_PAIR_CASES = {
    # BASELINE config 2: 1e7 uniform particles, 512^3, PCS, 20 lin bins on [0.005, 0.205]
    "C2": dict(n=10**7, L=1000., ngrid=512, assignment="pcs", bin_range=(0.005, 0.205),
               num_bins=20, seed=42,
               pairs=[(0, 0), (0, 19), (19, 19), (3, 11), (9, 10), (0, 1), (10, 10), (5, 17)]),
    # BASELINE config 5 at the largest mesh the oracle fits in host memory:
    # 512^3, 40 bins on [0.005, 0.405], L = 1000 (same wavenumbers per cell as 1024^3 / L = 2000)
    "C5proxy": dict(n=10**7, L=1000., ngrid=512, assignment="pcs", bin_range=(0.005, 0.405),
                    num_bins=40, seed=42,
                    pairs=[(0, 0), (0, 39), (39, 39), (17, 30), (20, 21), (38, 39)]),
}


# -- full estimator calls -----------------------------------------------------

def c1_inputs(golden_dir):
    """BASELINE config 1 exactly: the reference's 3-particle catalogue, periodised,
    L = 1000, 64^3, TSC, B_000 diag, 10 lin bins on [0.005, 0.105]
    (reference tests/conftest.py:105-122)."""
    data = np.loadtxt(golden_dir / "test_data_catalogue.txt").T
    pos = periodise(data[:3], 1000.)
    return dict(stat="bispec", catalogue_type="sim", pos_d=pos, nz_d=data[3], boxsize=1000.,
                ngrid=64, assignment="tsc", degrees=(0, 0, 0), form="diag",
                bin_range=(0.005, 0.105), num_bins=10)


def c4_inputs(which):
    """BASELINE config 4: zeta_110 diag on the C2 catalogue, 512^3, TSC; three of the
    twenty 10-Mpc/h bins of [5, 205] per fixture (the bin edges 5 + 10 i are exact)."""
    rng = {"lo": (5., 35.), "hi": (175., 205.)}[which]
    return dict(stat="3pcf", catalogue_type="sim", pos_d=uniform_box(10**7, 1000.),
                boxsize=1000., ngrid=512, assignment="tsc", degrees=(1, 1, 0), form="diag",
                bin_range=rng, num_bins=3)


def c3_inputs(nd=10**6, nr=5 * 10**7):
    """BASELINE config 3: survey B_202 diag, data + 5e7 randoms in an octant shell,
    FKP-like weights, LOS from the original coordinates, centred on the randoms,
    L = 2000, 512^3, TSC; three of the twenty bins of [0.005, 0.205]
    (0.005 + 0.01 i for i = 8..11: edges 0.085 .. 0.115)."""
    pd_ = shell_octant(np.random.default_rng(42), nd)
    pr_ = shell_octant(np.random.default_rng(43), nr)
    los_d, los_r = compute_los(pd_), compute_los(pr_)
    pd_c, pr_c = centre(pd_, pr_, 2000.)
    nz = 3.e-4
    wc = 1. / (1. + 1.e4 * nz)
    return dict(stat="bispec", catalogue_type="survey", boxsize=2000., ngrid=512,
                assignment="tsc", degrees=(2, 0, 2), form="diag",
                bin_range=(0.085, 0.115), num_bins=3,
                pos_d=np.ascontiguousarray(pd_c), nz_d=np.full(nd, nz), wc_d=np.full(nd, wc),
                los_d=los_d,
                pos_r=np.ascontiguousarray(pr_c), nz_r=np.full(nr, nz), wc_r=np.full(nr, wc),
                los_r=los_r)
