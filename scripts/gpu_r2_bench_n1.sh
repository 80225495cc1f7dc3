#!/bin/bash
# the default bench line and the reference arm at one GPU (what the driver runs at round end)
mkdir -p gpurun_out/final
T=gpurun_out/final
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 1200 python bench.py > ${T}/bench_n1.json 2> ${T}/bench_n1.err ) 2> ${T}/bench_n1.time; echo "bench rc=$?"; tail -3 ${T}/bench_n1.time; tail -c 300 ${T}/bench_n1.err
( time timeout 600 python bench.py --impl reference > ${T}/bench_reference.json 2> ${T}/bench_reference.err ) 2> ${T}/bench_reference.time; echo "reference arm rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/final/bench_n1.json')); r = json.load(open('gpurun_out/final/bench_reference.json'))
print('same config:', d['config'] == r['config'], d['config'])
print('C2 value %.4g q/s  ms %.4f  frac %.3f  traffic current %s  e2e %.4g  cpu %.4g  ref arm %.4g' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['traffic_capture']['current'], d['e2e']['value'], d['cpu_baseline']['value'], r['value']))
for k, v in d['workloads'].items():
    print(k, 'ms', v.get('ms_per_step') or v.get('build_ms'), 'frac', (v.get('roofline') or {}).get('frac') or v.get('frac_of_hbm_peak'), 'current', ((v.get('roofline') or {}).get('traffic_capture') or {}).get('current'), 'check', (v.get('check') or {}).get('bit_exact'))
PY
