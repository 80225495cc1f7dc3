#!/bin/bash
# round 2, call 3: after in-order tile scheduling for binned batches, the reduce kernel's load-ahead, chains beside
# the reduction, pool scratch for big builds: tests, the full bench line, build timings + launch list, ncu captures.
mkdir -p gpurun_out
T=gpurun_out/r2c3
timeout 1500 python -m pytest tests -m gpu -q --maxfail=60 -p no:cacheprovider > ${T}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 ${T}_pytest.log
( time timeout 1200 python bench.py --steps 20 --warmup 3 > ${T}_bench_full.json 2> ${T}_bench_full.err ) 2> ${T}_bench_full.time; echo "bench rc=$?"; tail -4 ${T}_bench_full.time; tail -c 600 ${T}_bench_full.err
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r2c3_bench_full.json'))
    print('HEAD', d['config']['workload'][:30], 'ms=%.4f frac=%.3f median=%.4f e2e=%.3g check=%s' % (d['ms_per_step'], d['roofline']['frac'], d['per_step']['median_ms'], d['e2e']['value'], (d.get('check') or {}).get('bit_exact')))
    sb = d.get('spline_build') or {}
    print('BUILD seq=%.4f rowsplit=%.4f levels=%s' % (sb['sequential']['ms'], sb['rowsplit']['ms'], sb['rowsplit']['levels']))
    print('E2E', {k: (round(v, 2) if isinstance(v, float) else v) for k, v in d['e2e'].items() if 'GBps' in k or 'frac' in k})
    for k, v in d['workloads'].items():
        if 'error' in v: print(k, 'ERROR', v['error'], v.get('trace')); continue
        print(k, {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ('ms_per_step', 'value', 'build_ms', 'allgather_ms', 'rowsplit_levels')},
              'frac=%s' % (v.get('roofline') or {}).get('frac'), 'check=%s' % (v.get('check') or {}).get('bit_exact'), 'e2e=%s' % (v.get('e2e') or {}).get('value'), 'cpu=%s' % (v.get('cpu_baseline') or {}).get('value'))
except Exception as e:
    print('bench_full FAILED', e)
PY
python scripts/bench_spline_build.py c2 long --levels 0,4,5,6 --bc Natural,Periodic > ${T}_spline_build.jsonl 2> ${T}_spline_build.err
python scripts/bench_spline_build.py c5b-shard wide c2 long --levels 0 --bc NotAKnot,Individual >> ${T}_spline_build.jsonl 2>> ${T}_spline_build.err
cat ${T}_spline_build.jsonl; tail -3 ${T}_spline_build.err
python scripts/bench_spline_build.py c2 long --levels 0 --bc Natural > ${T}_build_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${T}_build_launches.csv python scripts/bench_spline_build.py c2 long --levels 0 --bc Natural > ${T}_build_ncu.log 2>&1
python - <<'PY'
import csv
try:
    rows = [r for r in csv.reader(open('gpurun_out/r2c3_build_launches.csv')) if len(r) > 5]
    h = rows[0]; ki, vi = h.index('Kernel Name'), h.index('Metric Value')
    seq = [(r[ki][:60], float(r[vi].replace(',', '')) / 1000) for r in rows[1:]]
    idx = [k for k, (n, _) in enumerate(seq) if 'rowsplit_matrix' in n]
    for tag, i in (('c2', idx[0]), ('long', idx[-1])):
        print('---', tag)
        for n, v in seq[i:i + 12]: print('%-62s %.1f us' % (n, v))
except Exception as e:
    print('launch list FAILED', e)
PY
cap() { # name workload kernel-regex skip
  local name=$1 wl=$2 re=$3 skip=$4
  python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > ${T}_cap_plain_$name.log 2>&1 && \
  ncu --set full --import-source on --clock-control none -k regex:$re -s $skip -c 1 -o ${T}_$name python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > ${T}_cap_$name.log 2>&1
  python profiles/summarize_ncu.py ${T}_$name.ncu-rep gpurun_out/r2_ncu_$name.txt "$name: bench.py --workload $wl, kernel $re" > /dev/null 2>&1
  rm -f ${T}_$name.ncu-rep
  head -22 gpurun_out/r2_ncu_$name.txt | tail -17
}
cap c5a_bilinear_binned c5a interp2d_bilinear 4
cap c5a_bin_scatter c5a bin_scatter 4
cap c5a_bin_totals c5a bin_totals 4
cap c4_bilinear c4 interp2d_bilinear 4
cap c2_cubic c2 interp1d_cubic 4
cap c5b_cubic c5b interp1d_cubic 4
cap c3_linear_pair c3 interp1d_linear 4
