"""Single-process multi-GPU mode: one trv_threept_box_arrays call spreads over every usable
GPU (one host thread per device inside libtrv_b200.so; TRV_GPU_MAXNUM caps them).  Times
BASELINE configs 2 and 5 with 1 and all GPUs; prints one JSON object."""
import json, os, sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from triumvirate_b200 import core
import bench

res = {"gpus_visible": core.gpu_count()}
for name in ("C2", "C5"):
    wl = bench.WORKLOADS[name]
    pos = bench.make_catalogue(wl)
    d = torch.from_numpy(pos).to("cuda:0"); del pos
    torch.cuda.synchronize()
    kw = dict(boxsize=wl["L"], ngrid=wl["ngrid"], assignment=wl["assignment"], degrees=wl["degrees"],
              form=wl["form"], bin_range=wl["bin_range"], num_bins=wl["num_bins"], norm_factor=1.)
    out = {}
    for cap in (1, core.gpu_count()):
        os.environ["TRV_GPU_MAXNUM"] = str(cap)
        ts = []
        for it in range(4):
            t = time.perf_counter()
            r = core.threept_box_arrays("bispec", wl["n"], d[0].data_ptr(), d[1].data_ptr(),
                                        d[2].data_ptr(), True, **kw)
            ts.append(time.perf_counter() - t)
        out[f"gpus_{cap}"] = {"ms": [round(1e3 * t, 2) for t in ts], "bk0": float(r["bk_raw"][0].real),
                              "bk_last": float(r["bk_raw"][-1].real)}
    res[name] = out
    del d
    core.release_contexts(); torch.cuda.empty_cache()
print(json.dumps(res))
