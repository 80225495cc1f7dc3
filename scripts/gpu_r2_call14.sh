#!/bin/bash
# round 2, call 14: partition spline build -- parity tests, wall time per build against the other two modes, launch list
mkdir -p gpurun_out
T=gpurun_out/r2c14
timeout 900 python -m pytest tests/test_partition_gpu.py tests/test_rowsplit_gpu.py tests/test_parity_spline_gpu.py -m gpu -q --maxfail=20 -p no:cacheprovider > ${T}_pytest.log 2>&1
echo "pytest rc=$?"; tail -15 ${T}_pytest.log
python scripts/bench_spline_build.py --levels 0 --blocks 0 --bc Natural,Periodic,Individual > ${T}_build.jsonl 2> ${T}_build.err || tail -c 600 ${T}_build.err
python - <<'PY'
import json
for ln in open('gpurun_out/r2c14_build.jsonl'):
    d = json.loads(ln); print('%-10s %-10s %-10s %4d  %.4f ms  %.0f GB/s' % (d['shape'], d['boundary'], d['mode'], d['levels'], d['ms'], d['algorithmic_GBps']))
PY
python scripts/bench_spline_build.py c2 long --levels 0 --blocks 0 --bc Natural > ${T}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${T}_launches.csv python scripts/bench_spline_build.py c2 long --levels 0 --blocks 0 --bc Natural > ${T}_ncu.log 2>&1
python - <<'PY'
import csv
try:
    rows = [r for r in csv.reader(open('gpurun_out/r2c14_launches.csv')) if len(r) > 5]
    h = rows[0]; ki, vi = h.index('Kernel Name'), h.index('Metric Value')
    seq = [(r[ki][:60], float(r[vi].replace(',', '')) / 1000) for r in rows[1:]]
    idx = [k for k, (n, _) in enumerate(seq) if 'part_factor' in n and (k == 0 or 'part_factor' not in seq[k - 1][0])]
    for tag, i in (('c2', idx[0]), ('long', idx[-1])):
        print('---', tag)
        for n, v in seq[i:i + 10]: print('%-62s %.1f us' % (n, v))
except Exception as e:
    print('launch list FAILED', e)
PY
