#!/bin/bash
# Round 2: pruned per-axis shell transform (one GPU and x-slabs): parity tests, C5 phase
# timings, per-kernel launch list of one C5 call, bench with the C5 key.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_production.py -m gpu -x -q -k "slab or C2 or C5proxy or live or shell or C1" > gpurun_out/r02k_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02k_pytest.log; tail -8 gpurun_out/r02k_pytest.log
C5_PROBE_SET=quick timeout 900 python scripts/c5_slab_probe.py > gpurun_out/r02k_c5_probe.txt 2> gpurun_out/r02k_c5_probe.err; head -c 3000 gpurun_out/r02k_c5_probe.txt; tail -3 gpurun_out/r02k_c5_probe.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02k_c5_launches.csv python scripts/c5_share_once.py 0 1 > gpurun_out/r02k_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r02k_c5_launches.csv')) if len(r) > 5]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
tot = collections.defaultdict(lambda: [0, 0.])
for r in rows[1:]:
    v = float(r[vi].replace(',', '')); u = r[ui]
    v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1., 's': 1e3}.get(u, 1e-6)
    k = r[ki][:70]; tot[k][0] += 1; tot[k][1] += v
for k, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{t:10.2f} ms {c:6d}  {k}")
PY
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02k_bench_n1.json 2> gpurun_out/r02k_bench_n1.err; tail -c 1500 gpurun_out/r02k_bench_n1.json; tail -3 gpurun_out/r02k_bench_n1.err
