#!/bin/bash
# Round-end evidence: GPU tests, both bench arms, launch list, assignment sweep, configs.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
BENCH_DEBUG=1 timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; tail -c 400 gpurun_out/bench.log; tail -2 gpurun_out/bench.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -c 600 gpurun_out/bench_ref.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; wc -l gpurun_out/launches.csv
timeout 900 python scripts/assign_sweep.py > gpurun_out/assign_sweep.json 2> gpurun_out/assign_sweep.err; tail -3 gpurun_out/assign_sweep.err
timeout 900 python scripts/run_configs.py C1 C4 C3 C5 > gpurun_out/configs.json 2> gpurun_out/configs.err; cat gpurun_out/configs.json
