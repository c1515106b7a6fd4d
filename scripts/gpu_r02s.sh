#!/bin/bash
# Distributed mesh phase at N = $1: tests, stage trace, per-rank phases, bench.
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02s_pytest_multi.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02s_pytest_multi.log; tail -4 gpurun_out/r02s_pytest_multi.log
for wl in C2 C5; do
  echo "== $wl N=$N" | tee -a gpurun_out/r02s_trace_n$N.txt
  TRV_DIST_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 scripts/dist_phases.py $wl 2>&1 | grep -E "^rank|\[dist\] rank 0" | tail -$((N + 3)) | tee -a gpurun_out/r02s_trace_n$N.txt
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02s_bench_n$N.json 2> gpurun_out/r02s_bench_n$N.err; tail -c 1700 gpurun_out/r02s_bench_n$N.json | head -c 900; tail -3 gpurun_out/r02s_bench_n$N.err
