#!/bin/bash
# compute-sanitizer (memcheck, racecheck) over the new shared-memory FFT kernels.
mkdir -p gpurun_out
timeout 100 compute-sanitizer --tool memcheck --error-exitcode 66 --print-limit 10 python scripts/sanitize_subgrid.py > gpurun_out/r02an_memcheck.log 2>&1; echo "memcheck exit $?" | tee -a gpurun_out/r02an_memcheck.log; tail -3 gpurun_out/r02an_memcheck.log
timeout 100 compute-sanitizer --tool racecheck --error-exitcode 66 --print-limit 10 python scripts/sanitize_subgrid.py > gpurun_out/r02an_racecheck.log 2>&1; echo "racecheck exit $?" | tee -a gpurun_out/r02an_racecheck.log; tail -3 gpurun_out/r02an_racecheck.log
