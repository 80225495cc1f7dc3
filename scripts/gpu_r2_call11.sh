#!/bin/bash
# round 2, call 11: streaming host pipeline (queries uploaded chunk by chunk beside the copy-out): parity suite, e2e of every workload
mkdir -p gpurun_out
T=gpurun_out/r2c11
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > ${T}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 ${T}_pytest.log
run() {  # tag workload env...
  local tag=$1 wl=$2; shift 2
  env "$@" timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu --e2e-steps 5 > ${T}_${wl}_$tag.json 2> ${T}_${wl}_$tag.err || tail -c 400 ${T}_${wl}_$tag.err
  python - <<PY
import json
try:
    d = json.load(open('${T}_${wl}_$tag.json'))
    e = d['e2e']
    print('$wl $tag ms=%.4f check=%s e2e=%.4g q/s  e2e_ms=%.3f d2h=%.1f GB/s of link %s e2e_check=%s' % (d['ms_per_step'], (d.get('check') or {}).get('bit_exact'), e['value'], e['ms_per_step'], e.get('d2h_achieved_GBps', 0), e.get('d2h_link_GBps'), (d.get('check') or {}).get('e2e_bit_exact')))
except Exception as e:
    print('$wl $tag FAILED', e)
PY
}
for wl in c4 c4x c3 c5a c5b c2 c1; do run stream $wl NDI_X=1; done
python scripts/bench_pageable.py > ${T}_pageable.log 2>&1; tail -12 ${T}_pageable.log
python scripts/bench_latency.py > ${T}_latency.log 2>&1; tail -12 ${T}_latency.log
