#!/bin/bash
# Round 2, first GPU call: production-shape parity tests, bench with parity_check, measured reference arm.
mkdir -p gpurun_out
nproc > gpurun_out/r02a_host.txt; free -g >> gpurun_out/r02a_host.txt; nvidia-smi -L >> gpurun_out/r02a_host.txt
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r02a_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02a_pytest_gpu.log; tail -25 gpurun_out/r02a_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02a_smoke.log 2>&1; tail -1 gpurun_out/r02a_smoke.log
BENCH_DEBUG=1 timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; tail -c 1500 gpurun_out/r02a_bench.json; tail -2 gpurun_out/r02a_bench.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02a_bench_ref.json 2>&1; tail -c 1800 gpurun_out/r02a_bench_ref.json
