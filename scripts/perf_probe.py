"""Development probe: phase timings of the C2 workload (10M uniform particles,
512^3, PCS, B_000 full/triu, 20 bins) on one GPU."""
import sys, time, json
sys.path.insert(0, '.')
import numpy as np, torch
from triumvirate_b200 import core

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10**7
ng = int(sys.argv[2]) if len(sys.argv) > 2 else 512
nb = int(sys.argv[3]) if len(sys.argv) > 3 else 20
L = float(sys.argv[4]) if len(sys.argv) > 4 else 1000.
kmax = 0.005 + 0.01 * nb
pos = np.random.default_rng(42).uniform(0., L, size=(3, n))
dev = torch.device('cuda:0')
d = torch.from_numpy(pos).to(dev)
torch.cuda.synchronize()
kw = dict(boxsize=L, ngrid=ng, assignment='pcs', degrees=(0, 0, 0), form='full',
          bin_range=(0.005, kmax), num_bins=nb, norm_factor=1.)
core.profile_enable(True)
for it in range(3):
    t = time.perf_counter()
    out = core.threept_box_arrays('bispec', n, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), True, **kw)
    dt = time.perf_counter() - t
    print(f'iter {it}: {dt*1e3:.1f} ms (device-resident, profiled)', json.dumps({k: round(v*1e3, 2) for k, v in core.profile_report().items()}))
core.profile_enable(False)
for it in range(3):
    t = time.perf_counter()
    out = core.threept_box_arrays('bispec', n, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), True, **kw)
    dt = time.perf_counter() - t
    print(f'iter {it}: {dt*1e3:.1f} ms (device-resident)')
ph = torch.from_numpy(pos).pin_memory()
for it in range(3):
    t = time.perf_counter()
    out = core.threept_box_arrays('bispec', n, ph[0].data_ptr(), ph[1].data_ptr(), ph[2].data_ptr(), False, **kw)
    dt = time.perf_counter() - t
    print(f'iter {it}: {dt*1e3:.1f} ms (pinned host arrays)')
print('dim', len(out['bk_raw']), out['bk_raw'][:3], out['bk_shot'][:3], out['nmodes_1'][:3])
