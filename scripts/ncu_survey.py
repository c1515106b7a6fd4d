"""Small survey-type B_202 run (2e5 data + 2e6 randoms, 512^3) for kernel launch lists."""
import sys, time, json
sys.path.insert(0, '.')
import numpy as np
from triumvirate_b200 import core, catalogue as tcat
def shell_octant(gen, n):
    r = (gen.uniform(500.**3, 1500.**3, n)) ** (1. / 3.)
    mu = gen.uniform(0., 1., n); ph = gen.uniform(0., np.pi / 2, n)
    s = np.sqrt(1 - mu**2)
    return np.array([r * s * np.cos(ph), r * s * np.sin(ph), r * mu])
gd, gr = np.random.default_rng(42), np.random.default_rng(43)
nd, nr = 2 * 10**5, 2 * 10**6
pd_, pr_ = shell_octant(gd, nd), shell_octant(gr, nr)
los_d, los_r = tcat.compute_los(pd_), tcat.compute_los(pr_)
pd_c, pr_c = tcat.centre(pd_, pr_, 2000.)
nz = 3.e-4; wc = 1. / (1. + 1.e4 * nz)
kw = dict(boxsize=2000., ngrid=512, assignment="tsc", degrees=(2, 0, 2), form="diag",
          bin_range=(0.005, 0.205), num_bins=20,
          pos_d=pd_c, nz_d=np.full(nd, nz), wc_d=np.full(nd, wc), los_d=los_d,
          pos_r=pr_c, nz_r=np.full(nr, nz), wc_r=np.full(nr, wc), los_r=los_r)
norm = core.norm_particles(pr_c, kw["nz_r"], wc=kw["wc_r"], alpha=nd / nr)
core.profile_enable(len(sys.argv) > 1)
for it in range(2):
    t = time.perf_counter(); out = core.threept("bispec", "survey", norm_factor=norm, **kw); dt = time.perf_counter() - t
    print(f"iter {it}: {dt*1e3:.1f} ms", json.dumps({k: round(v*1e3, 2) for k, v in core.profile_report().items()}))
