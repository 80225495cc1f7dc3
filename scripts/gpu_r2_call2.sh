#!/bin/bash
# round 2, call 2: full GPU test-suite after the packed-arithmetic fix, the complete default bench line (all
# BASELINE workloads) and the reference arm, A/B of the evaluation-kernel changes, build timings, ncu captures.
mkdir -p gpurun_out
T=gpurun_out/r2c2
timeout 1500 python -m pytest tests -m gpu -q --maxfail=60 -p no:cacheprovider > ${T}_pytest.log 2>&1
echo "pytest rc=$?"; tail -12 ${T}_pytest.log
( time timeout 1200 python bench.py --steps 20 --warmup 3 > ${T}_bench_full.json 2> ${T}_bench_full.err ) 2> ${T}_bench_full.time; echo "bench rc=$?"; cat ${T}_bench_full.time | tail -4; tail -c 600 ${T}_bench_full.err
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r2c2_bench_full.json'))
    print('HEAD', d['config']['workload'][:30], 'ms=%.4f frac=%.3f e2e=%.3g check=%s build=%s' % (d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d.get('check'), d.get('spline_build')))
    for k, v in d['workloads'].items():
        if 'error' in v: print(k, 'ERROR', v['error'], v.get('trace')); continue
        print(k, {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ('ms_per_step', 'value', 'build_ms', 'allgather_ms', 'rowsplit_levels')},
              'frac=%s' % (v.get('roofline') or {}).get('frac'), 'check=%s' % v.get('check'), 'e2e=%s' % (v.get('e2e') or {}).get('value'), 'cpu=%s' % (v.get('cpu_baseline') or {}).get('value'))
except Exception as e:
    print('bench_full FAILED', e)
PY
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > ${T}_bench_ref.json 2> ${T}_bench_ref.err ) 2> ${T}_bench_ref.time; tail -3 ${T}_bench_ref.time; head -c 400 ${T}_bench_ref.json; echo
run() {  # tag workload env... (EXTRA: bench args)
  local tag=$1 wl=$2; shift 2
  env "$@" timeout 600 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu --e2e-steps 1 $EXTRA > ${T}_${wl}_$tag.json 2> ${T}_${wl}_$tag.err || tail -c 400 ${T}_${wl}_$tag.err
  python - <<PY
import json
try:
    d = json.load(open('${T}_${wl}_$tag.json'))
    print('$wl $tag ms=%.4f frac=%.3f median=%.4f best=%.4f check=%s' % (d['ms_per_step'], d['roofline']['frac'], d['per_step']['median_ms'], d['per_step']['best_ms'], (d.get('check') or {}).get('bit_exact')))
except Exception as e:
    print('$wl $tag FAILED', e)
PY
}
P=$PWD/ndarray_interp_b200
for wl in c3 c4 c5a c5b; do
  EXTRA="" run new $wl NDI_X=1
  EXTRA="" run nof2 $wl NDI_B200_LIB=$P/libndi_v_nof2.so
  EXTRA="" run nobc $wl NDI_B200_LIB=$P/libndi_v_nobc.so
  EXTRA="" run old $wl NDI_B200_LIB=$P/libndi_v_old.so NDI_PAIR_TABLE=0
done
EXTRA="--search-mode 5" run merge c2 NDI_X=1
EXTRA="--search-mode 5" run merge c5b NDI_X=1
python scripts/bench_spline_build.py c2 long --levels 0,3,4,5,6 --bc Natural,Periodic > ${T}_spline_build.jsonl 2> ${T}_spline_build.err
python scripts/bench_spline_build.py c5b-shard wide c2 long --levels 0 --bc NotAKnot,Individual >> ${T}_spline_build.jsonl 2>> ${T}_spline_build.err
cat ${T}_spline_build.jsonl; tail -3 ${T}_spline_build.err
# per-kernel times of the C2 build in both modes (cold-cache, serialised: shares, not absolutes)
python scripts/bench_spline_build.py c2 long --levels 0 --bc Natural > ${T}_build_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${T}_build_launches.csv python scripts/bench_spline_build.py c2 long --levels 0 --bc Natural > ${T}_build_ncu.log 2>&1
python - <<'PY'
import csv, collections
try:
    rows = [r for r in csv.reader(open('gpurun_out/r2c2_build_launches.csv')) if len(r) > 5]
    h = rows[0]; ki, vi = h.index('Kernel Name'), h.index('Metric Value')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        agg.setdefault(r[ki][:70], []).append(float(r[vi].replace(',', '')))
    for k, v in agg.items(): print('%-72s n=%3d median=%.1f us' % (k, len(v), sorted(v)[len(v)//2] / 1000.0))
except Exception as e:
    print('launch list FAILED', e)
PY
cap() { # name workload kernel-regex skip
  local name=$1 wl=$2 re=$3 skip=$4
  python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > ${T}_cap_plain_$name.log 2>&1 && \
  ncu --set full --import-source on --clock-control none -k regex:$re -s $skip -c 1 -o ${T}_$name python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > ${T}_cap_$name.log 2>&1
  python profiles/summarize_ncu.py ${T}_$name.ncu-rep gpurun_out/r2_ncu_$name.txt "$name: bench.py --workload $wl, kernel $re" > /dev/null 2>&1
  rm -f ${T}_$name.ncu-rep
  head -30 gpurun_out/r2_ncu_$name.txt
}
cap c3_linear_pair c3 interp1d_linear 4
cap c5a_bilinear_binned c5a interp2d_bilinear 4
