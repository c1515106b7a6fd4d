"""3PCF window function (trv::compute_3pcf_window) of 1e7 uniform randoms in a survey-like
geometry, 512^3, TSC, 20 r-bins: wall time per call (second call)."""
import json, sys, time
sys.path.insert(0, '.')
import numpy as np
from triumvirate_b200 import core, catalogue as tcat
n, L, ng = 10**7, 2000., 512
gen = np.random.default_rng(43)
r = gen.uniform(500.**3, 1500.**3, n) ** (1. / 3.)
mu = gen.uniform(0., 1., n); ph = gen.uniform(0., np.pi / 2, n); s = np.sqrt(1 - mu**2)
pos = np.array([r * s * np.cos(ph), r * s * np.sin(ph), r * mu])
los = tcat.compute_los(pos)
pos_c, _ = tcat.centre(pos, pos, L)
pos_c = tcat.periodise(pos_c, L)
nz = np.full(n, 3.e-4)
res = {}
for degrees, wa in (((0, 0, 0), False), ((2, 0, 2), False), ((2, 0, 2), True)):
    ts = []
    for _ in range(2):
        t = time.perf_counter()
        out = core.threept_window(pos_c, L, ng, "tsc", degrees, "diag", (50., 250.), 20, 1., los, alpha=1.,
                                  nz_r=nz, wide_angle=wa, wa_orders=(1, 0) if wa else (0, 0))
        ts.append(time.perf_counter() - t)
    key = f"zetaw{''.join(map(str, degrees))}{'_wa10' if wa else ''}"
    res[key] = dict(wall_s=ts, finite=bool(np.all(np.isfinite(out['zeta_raw'].view(float)))))
    print(key, [round(x, 3) for x in ts], file=sys.stderr, flush=True)
print(json.dumps(res))
