#!/bin/bash
# 1 -> 8 GPU bench of the default workload, as the driver launches it.
mkdir -p gpurun_out
NG=${1:-8}
for n in 1 2 4 8; do
  [ $n -gt $NG ] && continue
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n$n.log 2> gpurun_out/scale_n$n.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2960$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/scale_n$n.log 2> gpurun_out/scale_n$n.err
  fi
  echo "N=$n rc=$?"
  python - <<PY
import json
for l in open("gpurun_out/scale_n$n.log"):
    if l.startswith("{"):
        d=json.loads(l); print("  ms_per_step", round(d["ms_per_step"],3), "e2e ms", round(1e3*d["e2e"]["value"],3), d["config"]["parallelism"])
PY
done
