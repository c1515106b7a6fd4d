"""Under torchrun: phase timers of one estimator call per rank with the NCCL communicator
attached (distributed mesh phase unless TRV_NO_DIST_MESH=1).  argv: workload name (C2 | C5)."""
import json, os, sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import torch.distributed as dist
from triumvirate_b200 import core, dist as tdist
import bench

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl")
tdist.init_comm()
wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "C2"]
pos = bench.make_catalogue(wl)
d = torch.from_numpy(pos).to(f"cuda:{local}"); del pos
torch.cuda.synchronize()
kw = dict(boxsize=wl["L"], ngrid=wl["ngrid"], assignment=wl["assignment"], degrees=wl["degrees"],
          form=wl["form"], bin_range=wl["bin_range"], num_bins=wl["num_bins"], norm_factor=1.,
          part_rank=rank, part_count=world)
def call():
    return core.threept_box_arrays("bispec", wl["n"], d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), True, **kw)
for it in range(3):
    call()
ms = []
for it in range(5):
    dist.barrier(); torch.cuda.synchronize()
    t = time.perf_counter(); out = call(); ms.append(round(1e3 * (time.perf_counter() - t), 2))
core.profile_enable(True)
call(); out = call()
rep = {k: round(v * 1e3, 2) for k, v in core.profile_report().items()}
core.profile_enable(False)
for r in range(world):
    dist.barrier()
    if r == rank:
        print(f"rank {rank}/{world} dmesh_calls {core.dmesh_call_count()} ms {ms} phases {json.dumps(rep)} bk0 {out['bk_raw'][0].real!r} sn0 {out['bk_shot'][0].real!r}", flush=True)
core.comm_finalize()
dist.destroy_process_group()
