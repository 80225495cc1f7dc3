#!/bin/bash
# round 2, call 17: AUTO = partition build from 1024 rows on -- the whole GPU suite, then the default bench line
mkdir -p gpurun_out
T=gpurun_out/r2c17
timeout 2400 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > ${T}_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 ${T}_pytest.log
timeout 1200 python bench.py > ${T}_bench.json 2> ${T}_bench.err
echo "bench rc=$?"; tail -c 400 ${T}_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2c17_bench.json'))
print('C2 value %.4g q/s  ms %.4f  frac %.3f  e2e %.4g' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value']))
print('check', d['check'])
sb = d['spline_build']
print('build:', {k: (round(v['ms'], 4), v['levels']) for k, v in sb.items() if isinstance(v, dict)}, 'auto', sb['auto'], sb['ms'], 'frac', round(sb['frac_of_hbm_peak'], 4))
for k, v in d['workloads'].items():
    print(k, {kk: v.get(kk) for kk in ('ms_per_step', 'build_ms', 'build_info', 'frac_of_hbm_peak')}, 'frac', (v.get('roofline') or {}).get('frac'), 'check', (v.get('check') or {}).get('bit_exact'))
PY
