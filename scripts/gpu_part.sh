#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python scripts/c5_probe.py > gpurun_out/c5_probe.txt 2>&1; cat gpurun_out/c5_probe.txt
