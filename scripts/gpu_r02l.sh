#!/bin/bash
# fresh-process timings of single C5 shares (what one rank of an N-GPU run executes)
mkdir -p gpurun_out
for rw in "0 8" "7 8" "0 4" "3 4" "0 1"; do
  timeout 300 python scripts/c5_share_once.py $rw 5 2>&1 | grep share | tee -a gpurun_out/r02l_shares.txt
done
