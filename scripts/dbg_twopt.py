import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from triumvirate_b200 import core, catalogue as tcat
from oracle import ref as oracle
from test_gpu_parity import _survey_inputs
np.set_printoptions(linewidth=200, precision=10)
L, ng = 1000., 32
pd_, pr_, nzd, nzr, wsd, wsr, wcd, wcr = _survey_inputs(33, 1500, 6000, L)
los_d, los_r = tcat.compute_los(pd_), tcat.compute_los(pr_)
pd_c, pr_c = tcat.centre(pd_, pr_, L)
alpha = wsd.sum() / wsr.sum()
norm = oracle.norm_particles_2pt(pr_c, nzr, ws=wsr, wc=wcr, alpha=alpha)
for degree, assignment, il in ((2, "pcs", True), (2, "pcs", False), (0, "pcs", True)):
    kw = dict(boxsize=L, ngrid=ng, assignment=assignment, degree=degree, bin_range=(0.01, 0.10),
              num_bins=5, interlace=il, pos_d=pd_c, nz_d=nzd, ws_d=wsd, wc_d=wcd, los_d=los_d,
              pos_r=pr_c, nz_r=nzr, ws_r=wsr, wc_r=wcr, los_r=los_r, norm_factor=norm)
    a = core.twopt("powspec", "survey", **kw); b = oracle.twopt("powspec", "survey", **kw)
    print(degree, assignment, il)
    for k in ("pk_raw", "pk_shot"):
        print(" ", k, "out", a[k]); print(" ", k, "ref", b[k])
