#!/bin/bash
# Distributed mesh phase: per-rank phase timings under torchrun (N = $1), dist vs replicated.
N=${1:-2}
mkdir -p gpurun_out
for wl in C2 C5; do
  for flag in 0 1; do
    echo "== $wl N=$N TRV_NO_DIST_MESH=$flag" | tee -a gpurun_out/r02q_phases_n$N.txt
    TRV_NO_DIST_MESH=$flag timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$flag scripts/dist_phases.py $wl 2>&1 | grep "^rank" | tee -a gpurun_out/r02q_phases_n$N.txt
  done
done
