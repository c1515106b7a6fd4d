"""C5 (1e8 particles, 1024^3, PCS, 40 bins, 820 pairs) phase timings on ONE GPU,
as the whole job and as rank 0 of an 8-way pair partition (what every rank of the
8-GPU run executes, minus the final all-reduce)."""
import sys, time, json
sys.path.insert(0, '.')
import numpy as np, torch
from triumvirate_b200 import core
n, ng, nb, L = 10**8, 1024, 40, 2000.
pos = np.random.default_rng(42).uniform(0., L, size=(3, n))
d = torch.from_numpy(pos).to('cuda:0'); torch.cuda.synchronize()
kw = dict(boxsize=L, ngrid=ng, assignment='pcs', degrees=(0, 0, 0), form='full',
          bin_range=(0.005, 0.405), num_bins=nb, norm_factor=1.)
core.profile_enable(True)
for rank, count in ((0, 1), (0, 8), (1, 8), (3, 8), (5, 8), (6, 8), (7, 8)):
    for it in range(2):
        t = time.perf_counter()
        out = core.threept_box_arrays('bispec', n, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), True,
                                      part_rank=rank, part_count=count, **kw)
        dt = time.perf_counter() - t
        print(f'part {rank}/{count} iter {it}: {dt*1e3:.1f} ms', json.dumps({k: round(v*1e3, 2) for k, v in core.profile_report().items()}), flush=True)
