#!/bin/bash
# BASELINE config 5 on N GPUs of one box (default 8), as the driver would launch bench.py.
mkdir -p gpurun_out
N=${1:-8}
if [ $N -eq 1 ]; then
  timeout 1200 python bench.py --gpus 1 --workload C5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_n1.log 2> gpurun_out/bench_c5_n1.err
else
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus $N --workload C5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_n$N.log 2> gpurun_out/bench_c5_n$N.err
fi
echo "rc=$?"; tail -3 gpurun_out/bench_c5_n$N.err | cut -c1-300
python - <<PY
import json
for l in open("gpurun_out/bench_c5_n$N.log"):
    if l.startswith("{"):
        d=json.loads(l); print("N=$N ms_per_step", round(d["ms_per_step"],2), "e2e ms", round(1e3*d["e2e"]["value"],2), d["clocks"])
PY
