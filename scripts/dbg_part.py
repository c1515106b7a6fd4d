import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from triumvirate_b200 import core, catalogue as tcat
from test_gpu_parity import _survey_inputs
L, ng = 1000., 32
pd_, pr_, nzd, nzr, wsd, wsr, wcd, wcr = _survey_inputs(77, 1200, 5000, L)
los_d, los_r = tcat.compute_los(pd_), tcat.compute_los(pr_)
pd_c, pr_c = tcat.centre(pd_, pr_, L)
for stat, degrees, form in (("bispec", (2, 0, 2), "diag"), ("bispec", (1, 1, 0), "full"), ("3pcf", (1, 1, 0), "diag")):
    rng = (0.01, 0.09) if stat == "bispec" else (40., 280.)
    kw = dict(boxsize=L, ngrid=ng, assignment="tsc", degrees=degrees, form=form, bin_range=rng,
              num_bins=4, norm_factor=1., pos_d=pd_c, nz_d=nzd, ws_d=wsd, wc_d=wcd, los_d=los_d,
              pos_r=pr_c, nz_r=nzr, ws_r=wsr, wc_r=wcr, los_r=los_r, deterministic=True)
    full = core.threept(stat, "survey", **kw)
    full2 = core.threept(stat, "survey", **kw)
    raw, shot = ("bk_raw", "bk_shot") if stat == "bispec" else ("zeta_raw", "zeta_shot")
    print(stat, degrees, form, "repeat identical:", np.array_equal(full[raw], full2[raw]), np.array_equal(full[shot], full2[shot]))
    for world in (2, 3):
        parts = [core.threept(stat, "survey", part_rank=r, part_count=world, **kw) for r in range(world)]
        for key in (raw, shot):
            tot = sum(p[key] for p in parts)
            print("  world", world, key, "max rel diff", np.max(np.abs(tot - full[key])) / np.max(np.abs(full[key])))
