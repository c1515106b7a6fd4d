#!/bin/bash
# k_xpass_fused: variants A/B on a random 512^3 mesh + one ncu --set full capture of the default.
mkdir -p gpurun_out
timeout 300 python scripts/xpass_bench.py 512 144 0,1,2,3 2>&1 | tee gpurun_out/r02ac_xpass_variants.txt
bash scripts/ncu_one.sh k_xpass_fused r02ac_k_xpass_fused > /dev/null
python scripts/ncu_summary.py gpurun_out/prof_r02ac_k_xpass_fused.csv > gpurun_out/r02ac_ncu_k_xpass_fused.txt; cat gpurun_out/r02ac_ncu_k_xpass_fused.txt
