#!/bin/bash
# Distributed mesh phase on 2 GPUs: parity test, then bench at N=2.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02p_pytest_multi.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02p_pytest_multi.log; tail -30 gpurun_out/r02p_pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02p_bench_n2.json 2> gpurun_out/r02p_bench_n2.err; tail -c 1600 gpurun_out/r02p_bench_n2.json; tail -5 gpurun_out/r02p_bench_n2.err
