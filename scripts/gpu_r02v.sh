#!/bin/bash
# Round-2 evidence on one GPU: all GPU tests, smoke, both bench arms, launch list of the bench.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02v_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02v_pytest_gpu.log; tail -4 gpurun_out/r02v_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02v_smoke.log 2>&1; tail -1 gpurun_out/r02v_smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02v_bench_C2.json 2> gpurun_out/r02v_bench_C2.err; tail -c 700 gpurun_out/r02v_bench_C2.json; tail -2 gpurun_out/r02v_bench_C2.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02v_bench_C2_reference.json 2> gpurun_out/r02v_bench_C2_reference.err; tail -c 500 gpurun_out/r02v_bench_C2_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02v_launches_bench_C2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-c5 > gpurun_out/r02v_ncu_bench.log 2>&1; wc -l gpurun_out/r02v_launches_bench_C2.csv
