"""Two C4 calls (1e7 uniform particles, 512^3, TSC, zeta_110 diag, 20 r-bins) for launch lists."""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
from triumvirate_b200 import core
n, L = 10**7, 1000.
pos = np.random.default_rng(42).uniform(0., L, size=(3, n))
d = torch.from_numpy(pos).to('cuda:0'); torch.cuda.synchronize()
kw = dict(boxsize=L, ngrid=512, assignment='tsc', degrees=(1, 1, 0), form='diag',
          bin_range=(5., 205.), num_bins=20, norm_factor=1.)
for it in range(2):
    out = core.threept_box_arrays('3pcf', n, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), True, **kw)
print(out['zeta_raw'][:2])
