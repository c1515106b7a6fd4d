#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_program_io.py -m gpu -x -q -k "slab or program" > gpurun_out/r02g_pytest.log 2>&1; tail -4 gpurun_out/r02g_pytest.log
timeout 900 python scripts/c5_slab_probe.py > gpurun_out/r02g_c5_slab_probe.txt 2> gpurun_out/r02g_c5_slab_probe.err; grep -v "^{" gpurun_out/r02g_c5_slab_probe.txt | cut -c1-420; tail -3 gpurun_out/r02g_c5_slab_probe.err
