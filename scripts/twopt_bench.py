"""Two-point estimators on the C2 catalogue (1e7 uniform particles, 512^3, PCS): wall time
per call (second call, host catalogue through trv_twopt) with and without interlacing, and
the reference's own C++ (oracle/_ref, all host cores) on the plain power spectrum."""
import json, sys, time
sys.path.insert(0, '.')
import numpy as np
from triumvirate_b200 import core

n, L, ng = 10**7, 1000., 512
pos = np.random.default_rng(42).uniform(0., L, size=(3, n))
nz = np.full(n, n / L**3)
norm = L**3 / float(n)**2
res = {}
for stat, rng, nb in (("powspec", (0.005, 0.205), 20), ("2pcf", (5., 205.), 20)):
    for degree in (0, 2):
        for il in (False, True):
            kw = dict(boxsize=L, ngrid=ng, assignment="pcs", degree=degree, bin_range=rng, num_bins=nb,
                      norm_factor=norm, pos_d=pos, nz_d=nz, interlace=il)
            ts = []
            for _ in range(3):
                t = time.perf_counter(); out = core.twopt(stat, "sim", **kw); ts.append(time.perf_counter() - t)
            key = f"{stat}/l{degree}/{'interlaced' if il else 'plain'}"
            res[key] = dict(wall_s=ts, estimator_s=out["elapsed_s"])
            print(key, [round(x, 4) for x in ts], round(out["elapsed_s"], 4), file=sys.stderr, flush=True)
if "--reference" in sys.argv:
    from oracle import ref
    if ref.available():
        kw = dict(boxsize=L, ngrid=ng, assignment="pcs", degree=0, bin_range=(0.005, 0.205), num_bins=20,
                  norm_factor=norm, pos_d=pos, nz_d=nz, interlace=False)
        t = time.perf_counter(); o = ref.twopt("powspec", "sim", **kw); dt = time.perf_counter() - t
        res["reference/powspec/l0/plain"] = dict(wall_s=dt, estimator_s=o["elapsed_s"], cores=ref.num_threads())
        kw["norm_factor"] = norm
        a = core.twopt("powspec", "sim", **kw)
        res["reference/max_rel_diff_pk"] = float(np.max(np.abs(a["pk_raw"] - o["pk_raw"]) / np.abs(o["pk_raw"])))
print(json.dumps(res))
