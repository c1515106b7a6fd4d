"""One small box bispectrum through the sub-grid path (k_xpass_fused<128>, k_shell_ypass<72>,
k_shell_zpass<72>, k_gram_dmma) for compute-sanitizer runs."""
import sys
sys.path.insert(0, '.')
import numpy as np
from triumvirate_b200 import core
gen = np.random.default_rng(3)
pos = gen.uniform(0., 1000., size=(3, 20000))
kw = dict(boxsize=1000., ngrid=128, assignment="pcs", degrees=(0, 0, 0), form="diag",
          bin_range=(0.01, 0.1), num_bins=5, norm_factor=1., pos_d=pos)
before = core.fused_mesh_call_count()
out = core.threept("bispec", "sim", **kw)
assert core.fused_mesh_call_count() == before + 1
print("bk_raw", out["bk_raw"][:2])
