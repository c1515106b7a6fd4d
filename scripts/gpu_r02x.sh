#!/bin/bash
# N = 4: bench under torchrun, then the C2 / C5 stage trace with NCCL's default and with more P2P channels.
mkdir -p gpurun_out
bash scripts/gpu_r02w.sh 4 4 2>&1 | grep -v "^rank [23]"
for ch in 16 32; do
  echo "== C5 N=4 NCCL_MIN_P2P_NCHANNELS=$ch" | tee -a gpurun_out/r02x_nccl_channels.txt
  NCCL_MIN_P2P_NCHANNELS=$ch NCCL_MAX_P2P_NCHANNELS=$ch TRV_DIST_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 scripts/dist_phases.py C5 2>&1 | grep -E "^rank 0/|\[dist\] rank 0" | tail -4 | tee -a gpurun_out/r02x_nccl_channels.txt
done
