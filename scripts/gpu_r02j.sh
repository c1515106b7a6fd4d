#!/bin/bash
# Round 2: the multi-GPU path on all GPUs of the box (run with gpurun --gpus N, N = 2..8):
# NCCL + single-process tests, bench under torchrun at N (C2 line + C5 slab-mode key),
# single-process multi-device timings.
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02j_gpus.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02j_pytest_multi.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02j_pytest_multi.log; tail -8 gpurun_out/r02j_pytest_multi.log
for n in $N 4; do
  [ "$n" -gt "$N" ] && continue
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r02j_bench_n$n.json 2> gpurun_out/r02j_bench_n$n.err; tail -c 1800 gpurun_out/r02j_bench_n$n.json; tail -3 gpurun_out/r02j_bench_n$n.err
  [ "$n" -eq 4 ] && break
done
TRV_GPU_MULTI=1 timeout 600 python scripts/multi_single_process.py > gpurun_out/r02j_single_process.json 2> gpurun_out/r02j_single_process.err; cat gpurun_out/r02j_single_process.json; tail -3 gpurun_out/r02j_single_process.err
