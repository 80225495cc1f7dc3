#!/bin/bash
# round 2, call 22: partition build with the right-hand sides formed by the (shared-memory staged) block solve: spline parity tests, wall times, launch lists
# C2, long and many-column shapes
mkdir -p gpurun_out
T=gpurun_out/r2c22
timeout 900 python -m pytest tests/test_partition_gpu.py tests/test_rowsplit_gpu.py tests/test_parity_spline_gpu.py tests/test_reference_cubic_spline.py tests/test_fuzz_gpu.py tests/test_fullsize_gpu.py -m gpu -q --maxfail=20 -p no:cacheprovider > ${T}_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 ${T}_pytest.log
python scripts/bench_spline_build.py --levels 0 --blocks 0 --bc Natural,Periodic,Individual > ${T}_build.jsonl 2> ${T}_build.err || tail -c 600 ${T}_build.err
python - <<'PY'
import json
for ln in open('gpurun_out/r2c22_build.jsonl'):
    d = json.loads(ln); print('%-10s %-10s %-10s %4d  %.4f ms  %.0f GB/s' % (d['shape'], d['boundary'], d['mode'], d['levels'], d['ms'], d['algorithmic_GBps']))
PY
python scripts/bench_spline_build.py c2 long c5b-shard wide --levels 0 --blocks 0 --bc Natural > ${T}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${T}_launches.csv python scripts/bench_spline_build.py c2 long c5b-shard wide --levels 0 --blocks 0 --bc Natural > ${T}_ncu.log 2>&1
python - <<'PY'
import csv
try:
    rows = [r for r in csv.reader(open('gpurun_out/r2c22_launches.csv')) if len(r) > 5]
    h = rows[0]; ki, vi = h.index('Kernel Name'), h.index('Metric Value')
    seq = [(r[ki][:64], float(r[vi].replace(',', '')) / 1000) for r in rows[1:]]
    # last partition build of every shape: find the last 'part_ab' launches separated by shape changes (time jumps)
    ends = [k for k, (n, _) in enumerate(seq) if 'part_ab' in n]
    shown, last = 0, None
    for e in ends:
        t = seq[e][1]
        if last is None or abs(t - last) / max(t, last) > 0.3:
            b = e
            while b > 0 and 'part_ab' not in seq[b - 1][0] and e - b < 12: b -= 1
            print('---')
            for n, v in seq[b:e + 1]: print('%-66s %.1f us' % (n, v))
        last = t
    seqs = [k for k, (n, _) in enumerate(seq) if 'spline_sweep' in n]
    print('--- sequential builds (factor, rhs, sweep, ab):')
    lastt = None
    for k in seqs:
        t = seq[k][1]
        if lastt is None or abs(t - lastt) / max(t, lastt) > 0.3:
            for n, v in seq[max(k - 2, 0):k + 2]: print('%-66s %.1f us' % (n, v))
            print()
        lastt = t
except Exception as e:
    print('launch list FAILED', e)
PY
