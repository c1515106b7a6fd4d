"""BASELINE.json configs 1, 3, 4, 5 on one GPU: wall time per estimator call
(second call, plans warm) and basic sanity of the outputs.  Config 2 is
bench.py's workload."""
import sys, time, json
sys.path.insert(0, '.')
import numpy as np
from triumvirate_b200 import core, catalogue as tcat

which = sys.argv[1:] or ["C1", "C4", "C3", "C5"]
res = {}

import os
PROFILE = bool(os.environ.get("TRV_PROFILE"))
if PROFILE:
    core.profile_enable(True)

def timed(fn, reps=2):
    out = None; ts = []
    for _ in range(reps):
        t = time.perf_counter(); out = fn(); ts.append(time.perf_counter() - t)
        if PROFILE:
            print(json.dumps({k: round(v * 1e3, 2) for k, v in core.profile_report().items()}),
                  file=sys.stderr, flush=True)
    return out, ts

def shell_octant(gen, n):
    r = (gen.uniform(500.**3, 1500.**3, n)) ** (1. / 3.)
    mu = gen.uniform(0., 1., n); ph = gen.uniform(0., np.pi / 2, n)
    s = np.sqrt(1 - mu**2)
    return np.array([r * s * np.cos(ph), r * s * np.sin(ph), r * mu])

if "C1" in which:
    data = np.loadtxt("tests/golden/test_data_catalogue.txt").T
    pos = tcat.periodise(data[:3], 1000.)
    norm = core.norm_particles(pos, data[3])
    out, ts = timed(lambda: core.threept("bispec", "sim", pos, 1000., 64, "tsc", (0, 0, 0), "diag",
                                         (0.005, 0.105), 10, norm, nz_d=data[3]), reps=3)
    res["C1"] = dict(times_s=ts, dim=len(out["bk_raw"]), bk0=complex(out["bk_raw"][0]).real)

if "C4" in which:
    pos = np.random.default_rng(42).uniform(0., 1000., size=(3, 10**7))
    out, ts = timed(lambda: core.threept("3pcf", "sim", pos, 1000., 512, "tsc", (1, 1, 0), "diag",
                                         (5., 205.), 20, 1.))
    res["C4"] = dict(times_s=ts, dim=len(out["zeta_raw"]), finite=bool(np.all(np.isfinite(out["zeta_raw"].view(float)))),
                     zeta0=[float(out["zeta_raw"][3].real), float(out["zeta_raw"][3].imag)], npairs=int(out["npairs_1"][3]))

if "C3" in which:
    gd, gr = np.random.default_rng(42), np.random.default_rng(43)
    nd, nr = 10**6, 5 * 10**7
    pd_, pr_ = shell_octant(gd, nd), shell_octant(gr, nr)
    los_d, los_r = tcat.compute_los(pd_), tcat.compute_los(pr_)
    pd_c, pr_c = tcat.centre(pd_, pr_, 2000.)
    nz = 3.e-4; wc = 1. / (1. + 1.e4 * nz)
    kw = dict(boxsize=2000., ngrid=512, assignment="tsc", degrees=(2, 0, 2), form="diag",
              bin_range=(0.005, 0.205), num_bins=20,
              pos_d=pd_c, nz_d=np.full(nd, nz), wc_d=np.full(nd, wc), los_d=los_d,
              pos_r=pr_c, nz_r=np.full(nr, nz), wc_r=np.full(nr, wc), los_r=los_r)
    alpha = nd / nr
    norm = core.norm_particles(pr_c, kw["nz_r"], wc=kw["wc_r"], alpha=alpha)
    out, ts = timed(lambda: core.threept("bispec", "survey", norm_factor=norm, **kw))
    res["C3"] = dict(times_s=ts, dim=len(out["bk_raw"]), finite=bool(np.all(np.isfinite(out["bk_raw"].view(float)))),
                     bk3=[float(out["bk_raw"][3].real), float(out["bk_raw"][3].imag)],
                     elapsed_in_estimator_s=out["elapsed_s"])

if "C5" in which:
    import torch
    n = 10**8
    pos = np.random.default_rng(42).uniform(0., 2000., size=(3, n))
    d = torch.from_numpy(pos).to("cuda:0"); torch.cuda.synchronize()
    kw = dict(boxsize=2000., ngrid=1024, assignment="pcs", degrees=(0, 0, 0), form="full",
              bin_range=(0.005, 0.405), num_bins=40, norm_factor=1.)
    out, ts = timed(lambda: core.threept_box_arrays("bispec", n, d[0].data_ptr(), d[1].data_ptr(),
                                                    d[2].data_ptr(), True, **kw), reps=3)
    res["C5"] = dict(times_s=ts, dim=len(out["bk_raw"]), finite=bool(np.all(np.isfinite(out["bk_raw"].view(float)))),
                     bk0=float(out["bk_raw"][0].real), nmodes0=int(out["nmodes_1"][0]),
                     gib_gpu_max=core.counters()["gib_gpu_max"])
print(json.dumps(res))
