"""C5 (1e8 particles, 1024^3, PCS, 40 bins, 820 pairs) on ONE GPU: the whole job, and single
shares of 2-, 4- and 8-way runs in slab mode and (TRV_NO_SLAB=1) pair-block mode -- what a
rank of the multi-GPU run executes, minus the final all-reduce -- with phase timings; plus the
check that the slab shares sum to the whole."""
import os, sys, time, json
sys.path.insert(0, '.')
import numpy as np, torch
from triumvirate_b200 import core
n, ng, nb, L = 10**8, 1024, 40, 2000.
pos = np.random.default_rng(42).uniform(0., L, size=(3, n))
d = torch.from_numpy(pos).to('cuda:0'); del pos; torch.cuda.synchronize()
kw = dict(boxsize=L, ngrid=ng, assignment='pcs', degrees=(0, 0, 0), form='full',
          bin_range=(0.005, 0.405), num_bins=nb, norm_factor=1.)
def run(rank, count):
    return core.threept_box_arrays('bispec', n, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), True,
                                   part_rank=rank, part_count=count, **kw)
res = {}
full = run(0, 1); full = run(0, 1)
core.profile_enable(True)
quick = os.environ.get("C5_PROBE_SET") == "quick"
combos = ((0, 1), (0, 2), (0, 8), (7, 8)) if quick else \
    ((0, 1), (0, 2), (1, 2), (0, 4), (3, 4), (0, 8), (3, 8), (7, 8))
for no_slab in (("0",) if quick else ("0", "1")):
    os.environ["TRV_NO_SLAB"] = no_slab
    for rank, count in combos:
        if count == 1 and no_slab == "1":
            continue
        for it in range(2):
            t = time.perf_counter()
            out = run(rank, count)
            dt = time.perf_counter() - t
        key = f"{'blocks' if no_slab == '1' else 'slab'} {rank}/{count}"
        res[key] = {"ms": round(dt * 1e3, 1), "phases_ms": {k: round(v * 1e3, 1) for k, v in core.profile_report().items()}}
        print(key, json.dumps(res[key]), flush=True)
core.profile_enable(False)
os.environ["TRV_NO_SLAB"] = "0"
raw = sum(run(r, 8)["bk_raw"] for r in range(8))
err = float(np.max(np.abs(raw - full["bk_raw"]) / np.abs(full["bk_raw"]).max()))
res["slab_8_shares_vs_full_max_rel"] = err
print("slab 8 shares vs full:", err, flush=True)
print(json.dumps(res))
