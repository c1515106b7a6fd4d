"""Per-step wall times of the C2 workload, no sampler; cgroup throttle counters before/after."""
import sys, time, os
sys.path.insert(0, '.')
import numpy as np, torch
from triumvirate_b200 import core
def cg():
    out = {}
    for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu.stat", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.stat"):
        try: out[f] = open(f).read().strip().replace("\n", "; ")
        except OSError: pass
    return out
print(cg()); print("affinity", len(os.sched_getaffinity(0)), "loadavg", os.getloadavg())
n, ng, nb, L = 10**7, 512, 20, 1000.
pos = np.random.default_rng(42).uniform(0., L, size=(3, n))
d = torch.from_numpy(pos).to('cuda:0'); torch.cuda.synchronize()
kw = dict(boxsize=L, ngrid=ng, assignment='pcs', degrees=(0, 0, 0), form='full',
          bin_range=(0.005, 0.205), num_bins=nb, norm_factor=1.)
def step():
    return core.threept_box_arrays('bispec', n, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), True, **kw)
for _ in range(4): step()
for rep in range(1):
    ts = []
    t0 = time.perf_counter()
    for _ in range(100):
        t = time.perf_counter(); step(); ts.append((time.perf_counter() - t) * 1e3)
    print(f"rep {rep}: mean {np.mean(ts):.2f} med {np.median(ts):.2f} max {np.max(ts):.2f}; outliers at", [(i, round(v, 1), round((sum(ts[:i]))%100,0)) for i, v in enumerate(ts) if v > 12])
    print(cg().get("/sys/fs/cgroup/cpu.stat"), "loadavg", os.getloadavg())
core.profile_enable(True)
worst = None
import collections
tot = collections.defaultdict(list)
for _ in range(200):
    t = time.perf_counter(); step(); dt = (time.perf_counter() - t) * 1e3
    rep_ = core.profile_report()
    for k, v in rep_.items(): tot[k].append(v * 1e3)
    if worst is None or dt > worst[0]: worst = (dt, {k: round(v * 1e3, 2) for k, v in rep_.items()})
print("worst profiled step", worst)

for k, v in tot.items(): print(f"{k:20s} med {np.median(v):7.3f} mean {np.mean(v):7.3f} max {np.max(v):8.3f} n>5ms {sum(1 for x in v if x > np.median(v) + 5)}")
