#!/bin/bash
# usage: ncu_one.sh <kernel-regex> [tag]  -- one `ncu --set full` capture of a kernel in the
# second C2 estimator call; raw metrics exported as CSV next to the report.
k=$1; tag=${2:-$1}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/prof_$tag python scripts/ncu_probe.py > gpurun_out/ncu_$tag.log 2>&1
tail -2 gpurun_out/ncu_$tag.log
ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/prof_$tag.csv 2>/dev/null
ls -la gpurun_out/prof_$tag.*
