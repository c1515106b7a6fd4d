#!/bin/bash
# Distributed mesh phase at N = $1 after the slab G / split binned statistics / cached partition:
# stage trace + per-rank phases (C2, C5, C5 with more shell groups), bench.
N=${1:-8}
mkdir -p gpurun_out
run() {  # label, env assignments..., workload last
  local label=$1; shift
  echo "== $label N=$N" | tee -a gpurun_out/r02u_trace_n$N.txt
  env "${@:1:$#-1}" TRV_DIST_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 scripts/dist_phases.py "${@: -1}" 2>&1 | grep -E "^rank [01]/|\[dist\] rank 0" | tail -5 | tee -a gpurun_out/r02u_trace_n$N.txt
}
run C2 X=0 C2
run C5 X=0 C5
run "C5 TRV_SHELL_GROUPS=10" TRV_SHELL_GROUPS=10 C5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02u_bench_n$N.json 2> gpurun_out/r02u_bench_n$N.err; python - <<PY
import json
j = json.loads(open('gpurun_out/r02u_bench_n$N.json').read().strip().splitlines()[-1])
print("N", j["n_gpus"], "value", j["value"], "e2e", j["e2e"]["value"], "c5", j["c5"]["ms_per_step"], j["result"]["sha256_deterministic_step"][:12], j["result"]["timed_vs_deterministic_max_rel"])
PY
tail -2 gpurun_out/r02u_bench_n$N.err
