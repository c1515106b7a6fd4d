#!/bin/bash
# compute-sanitizer over the small-shape GPU tests: memcheck (out-of-bounds / misaligned
# accesses) and racecheck (shared-memory hazards in the staged kernels).
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 66 --print-limit 20 \
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
  -k "tiles or bit_exact or empty_bins or degenerate or anisotropic or reference_goldens or twopt_goldens or window_goldens or streamed or mirror" \
  > gpurun_out/memcheck.log 2>&1
echo "memcheck exit $?"; tail -2 gpurun_out/memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 66 --print-limit 20 \
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tiles or bit_exact or gram or streamed" \
  > gpurun_out/racecheck.log 2>&1
echo "racecheck exit $?"; tail -3 gpurun_out/racecheck.log
