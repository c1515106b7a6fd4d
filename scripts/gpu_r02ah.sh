#!/bin/bash
# Round-2 final one-GPU evidence: all GPU tests, smoke, bench (with cpu_baseline + parity_check), launch list.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02ah_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02ah_pytest_gpu.log; tail -3 gpurun_out/r02ah_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02ah_smoke.log 2>&1; tail -1 gpurun_out/r02ah_smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02ah_bench.json 2> gpurun_out/r02ah_bench.err; tail -c 300 gpurun_out/r02ah_bench.json; tail -2 gpurun_out/r02ah_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02ah_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-c5 > gpurun_out/r02ah_ncu_bench.log 2>&1; wc -l gpurun_out/r02ah_launches.csv
