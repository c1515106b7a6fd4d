#!/bin/bash
# After the slab G / split binned statistics / grouped p2p gather: tests + phases at N = $1.
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02t_pytest_multi.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02t_pytest_multi.log; tail -12 gpurun_out/r02t_pytest_multi.log
for wl in C2 C5; do
  echo "== $wl N=$N" | tee -a gpurun_out/r02t_trace_n$N.txt
  TRV_DIST_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 scripts/dist_phases.py $wl 2>&1 | grep -E "^rank|\[dist\] rank 0" | tail -$((N + 3)) | tee -a gpurun_out/r02t_trace_n$N.txt
done
