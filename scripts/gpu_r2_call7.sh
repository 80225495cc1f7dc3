#!/bin/bash
# round 2, call 7: reduce kernel with per-row reciprocals (A/B tile height), scatter pass with packed records
mkdir -p gpurun_out
T=gpurun_out/r2c7
timeout 1500 python -m pytest tests -m gpu -q --maxfail=60 -p no:cacheprovider > ${T}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 ${T}_pytest.log
for sh in 2 3; do
  echo "--- NDI_RS_TILE_SHIFT=$sh"
  NDI_RS_TILE_SHIFT=$sh python scripts/bench_spline_build.py c2 long --levels 0,4,5,6 --bc Natural,Periodic 2>&1 | grep rowsplit | python -c "
import sys, json
for ln in sys.stdin:
    d = json.loads(ln); print(d['shape'], d['boundary'], 'L=%d' % d['levels'], '%.4f ms' % d['ms'])"
done
python scripts/bench_spline_build.py c2 long --levels 0 --bc Natural > ${T}_build_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${T}_build_launches.csv python scripts/bench_spline_build.py c2 long --levels 0 --bc Natural > ${T}_build_ncu.log 2>&1
python - <<'PY'
import csv
try:
    rows = [r for r in csv.reader(open('gpurun_out/r2c7_build_launches.csv')) if len(r) > 5]
    h = rows[0]; ki, vi = h.index('Kernel Name'), h.index('Metric Value')
    seq = [(r[ki][:60], float(r[vi].replace(',', '')) / 1000) for r in rows[1:]]
    idx = [k for k, (n, _) in enumerate(seq) if 'rowsplit_matrix' in n]
    for tag, i in (('c2', idx[0]), ('long', idx[-1])):
        print('---', tag)
        for n, v in seq[i:i + 12]: print('%-62s %.1f us' % (n, v))
except Exception as e:
    print('launch list FAILED', e)
PY
for wl in c5a c4x; do
  timeout 600 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu --e2e-steps 1 > ${T}_$wl.json 2> ${T}_$wl.err || tail -c 300 ${T}_$wl.err
  python -c "
import json
d = json.load(open('${T}_$wl.json')); print('$wl ms=%.4f frac=%.3f median=%.4f check=%s' % (d['ms_per_step'], d['roofline']['frac'], d['per_step']['median_ms'], d['check']['bit_exact']))"
done
python bench.py --workload c5a --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > ${T}_cap_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bin_|interp2d" -s 9 -c 3 --csv --log-file ${T}_c5a_launches.csv python bench.py --workload c5a --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > ${T}_c5a_ncu.log 2>&1
grep -E "bin_|interp2d" ${T}_c5a_launches.csv | awk -F'","' '{print $5, $NF}' | head
