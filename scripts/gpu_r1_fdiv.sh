#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/t_fdiv.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_fdiv.log
timeout 600 python scripts/exhaustive_fdiv.py --out gpurun_out/fdiv_exhaustive.txt; echo "exhaustive rc=$?"
run() { # name env...
  local name=$1; shift
  env "$@" timeout 600 python bench.py --workload $WL --steps 20 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/fd_${WL}_${name}.json 2> gpurun_out/fd_${WL}_${name}.err || tail -c 400 gpurun_out/fd_${WL}_${name}.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/fd_${WL}_${name}.json'))
    print('$WL $name', 'ms=%.4f'%d['ms_per_step'], 'GB/s=%.0f'%d['roofline']['achieved'], 'frac=%.3f'%d['roofline']['frac'], 'launches', d['gpu_launches'], 'e2e_ms=%.1f'%d['e2e']['ms_per_step'])
except Exception as e: print('$WL $name failed', e)
PY
}
for WL in c3; do run slow NDI_NO_FAST_DIV=1; run fast X=1; done
for WL in c4 c5a; do
  run slow_off NDI_NO_FAST_DIV=1 NDI_BIN_MODE=1
  run fast_off NDI_BIN_MODE=1
  run fast_on16 NDI_BIN_MODE=2 NDI_BAND_MB=16
done
