#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
for n in 2; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_n$n.log 2> gpurun_out/bench_n$n.err
echo "rc=$?"; python - <<PY
import json
for l in open("gpurun_out/bench_n$n.log"):
    if l.startswith("{"):
        d=json.loads(l); print("N=$n ms_per_step", d["ms_per_step"], "e2e", d["e2e"]["value"])
PY
done
timeout 900 python scripts/c5_probe.py > gpurun_out/c5_probe.txt 2>&1; grep "iter 1" gpurun_out/c5_probe.txt | cut -c1-340
