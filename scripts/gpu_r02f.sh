#!/bin/bash
# Round 2: full GPU suite after the coverage work (I/O + program, patched bindings, multi-GPU API).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=10 > gpurun_out/r02f_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02f_pytest_gpu.log; tail -40 gpurun_out/r02f_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02f_smoke.log 2>&1; tail -1 gpurun_out/r02f_smoke.log
