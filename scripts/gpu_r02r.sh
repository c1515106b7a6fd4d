#!/bin/bash
# Distributed mesh phase: stage trace (TRV_DIST_TRACE=1) at N = $1.
N=${1:-2}
mkdir -p gpurun_out
for wl in C2 C5; do
  echo "== $wl N=$N" | tee -a gpurun_out/r02r_trace_n$N.txt
  TRV_DIST_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 scripts/dist_phases.py $wl 2>&1 | grep -E "^rank|\[dist\] rank 0" | tail -12 | tee -a gpurun_out/r02r_trace_n$N.txt
done
