#!/bin/bash
# Round 2: multi-GPU behind the C API (run with gpurun --gpus 2): NCCL tests, single-process
# multi-device mode, bench at N = 1 and N = 2 with the C5 extra key.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02e_gpus.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02e_pytest_multi.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02e_pytest_multi.log; tail -15 gpurun_out/r02e_pytest_multi.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02e_bench_n1.json 2> gpurun_out/r02e_bench_n1.err; tail -c 1200 gpurun_out/r02e_bench_n1.json; tail -3 gpurun_out/r02e_bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02e_bench_n2.json 2> gpurun_out/r02e_bench_n2.err; tail -c 1500 gpurun_out/r02e_bench_n2.json; tail -5 gpurun_out/r02e_bench_n2.err
# single process, both GPUs, no Python in the split: C5 through the C++ entry point
TRV_GPU_MULTI=1 timeout 600 python scripts/multi_single_process.py > gpurun_out/r02e_single_process.json 2> gpurun_out/r02e_single_process.err; cat gpurun_out/r02e_single_process.json; tail -3 gpurun_out/r02e_single_process.err
