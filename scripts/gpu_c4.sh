mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_c4.csv python scripts/ncu_c4.py > gpurun_out/ncu_c4.log 2>&1; tail -2 gpurun_out/ncu_c4.log; wc -l gpurun_out/launches_c4.csv
BENCH_DEBUG=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err
