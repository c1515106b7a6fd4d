"""One C5 share (rank R of W, argv) for kernel-level timing under ncu."""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
from triumvirate_b200 import core
rank, world = int(sys.argv[1]), int(sys.argv[2])
n, ng, nb, L = 10**8, 1024, 40, 2000.
pos = np.random.default_rng(42).uniform(0., L, size=(3, n))
d = torch.from_numpy(pos).to('cuda:0'); del pos; torch.cuda.synchronize()
kw = dict(boxsize=L, ngrid=ng, assignment='pcs', degrees=(0, 0, 0), form='full',
          bin_range=(0.005, 0.405), num_bins=nb, norm_factor=1.)
import time
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2
ms = []
for it in range(iters):
    t = time.perf_counter()
    out = core.threept_box_arrays('bispec', n, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), True,
                                  part_rank=rank, part_count=world, **kw)
    ms.append(round(1e3 * (time.perf_counter() - t), 1))
print(f"share {rank}/{world} ms per call: {ms}", out['bk_raw'][:2])
