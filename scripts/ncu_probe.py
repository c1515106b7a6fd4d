"""Two C2 estimator calls (device-resident catalogue) for ncu captures."""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
from triumvirate_b200 import core
import bench
wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "C2"]
pos = bench.make_catalogue(wl)
d = torch.from_numpy(pos).to('cuda:0')
torch.cuda.synchronize()
kw = dict(boxsize=wl["L"], ngrid=wl["ngrid"], assignment=wl["assignment"], degrees=wl["degrees"],
          form=wl["form"], bin_range=wl["bin_range"], num_bins=wl["num_bins"], norm_factor=1.)
for it in range(2):
    out = core.threept_box_arrays('bispec', wl["n"], d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), True, **kw)
print(out['bk_raw'][:2])
