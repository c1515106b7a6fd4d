#!/bin/bash
# Round-2 multi-GPU evidence at N = $1 GPUs: NCCL / single-process / distributed-mesh tests (N = 2 only),
# bench under torchrun for every N in "$2" (default: N), per-rank phases of C2 and C5 at N.
N=${1:-2}
LIST=${2:-$N}
mkdir -p gpurun_out
if [ "$N" -eq 2 ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02w_pytest_multi.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02w_pytest_multi.log; tail -4 gpurun_out/r02w_pytest_multi.log
fi
for n in $LIST; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/r02w_bench_n$n.json 2> gpurun_out/r02w_bench_n$n.err
  python - <<PY
import json
j = json.loads(open('gpurun_out/r02w_bench_n$n.json').read().strip().splitlines()[-1])
print("N", j["n_gpus"], "value", j["value"], "e2e", j["e2e"]["value"], "c5", j["c5"]["ms_per_step"], j["result"]["sha256_deterministic_step"][:12], j["result"]["timed_vs_deterministic_max_rel"], j["run"]["distributed_mesh_calls"])
PY
done
for wl in C2 C5; do
  echo "== $wl N=$N" | tee -a gpurun_out/r02w_trace_n$N.txt
  TRV_DIST_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 scripts/dist_phases.py $wl 2>&1 | grep -E "^rank [01]/|\[dist\] rank 0" | tail -5 | tee -a gpurun_out/r02w_trace_n$N.txt
done
