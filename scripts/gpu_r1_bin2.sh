#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "bilinear or fdiv or fast" > gpurun_out/t_bin2.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_bin2.log
run() { # name env...
  local name=$1; shift
  env "$@" timeout 600 python bench.py --workload $WL --steps 20 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/b2_${WL}_${name}.json 2> gpurun_out/b2_${WL}_${name}.err || tail -c 400 gpurun_out/b2_${WL}_${name}.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/b2_${WL}_${name}.json'))
    print('$WL $name', 'ms=%.4f'%d['ms_per_step'], 'GB/s=%.0f'%d['roofline']['achieved'], 'frac=%.3f'%d['roofline']['frac'], 'launches', d['gpu_launches'], 'e2e_ms=%.1f'%d['e2e']['ms_per_step'])
except Exception as e: print('$WL $name failed', e)
PY
}
for WL in c4 c5a; do
  run off NDI_BIN_MODE=1
  for mb in 8 16 32; do run on$mb NDI_BIN_MODE=2 NDI_BAND_MB=$mb; done
done
WL=c5a ENVV="NDI_BIN_MODE=2 NDI_BAND_MB=16" bash scripts/gpu_ncu_quick.sh on2
WL=c4 ENVV="NDI_BIN_MODE=2 NDI_BAND_MB=16" bash scripts/gpu_ncu_quick.sh on2
