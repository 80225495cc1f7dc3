#!/bin/bash
# round 2, call 4: uniform grids staged in shared memory for the guess search (A/B), pair-kernel gather batching (A/B),
# ncu of the row-split reduce kernel
mkdir -p gpurun_out
T=gpurun_out/r2c4
timeout 1500 python -m pytest tests -m gpu -q --maxfail=60 -p no:cacheprovider > ${T}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 ${T}_pytest.log
run() {  # tag workload env...
  local tag=$1 wl=$2; shift 2
  env "$@" timeout 600 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu --e2e-steps 1 $EXTRA > ${T}_${wl}_$tag.json 2> ${T}_${wl}_$tag.err || tail -c 400 ${T}_${wl}_$tag.err
  python - <<PY
import json
try:
    d = json.load(open('${T}_${wl}_$tag.json'))
    print('$wl $tag ms=%.4f frac=%.3f median=%.4f best=%.4f check=%s e2e=%.4g' % (d['ms_per_step'], d['roofline']['frac'], d['per_step']['median_ms'], d['per_step']['best_ms'], (d.get('check') or {}).get('bit_exact'), d['e2e']['value']))
except Exception as e:
    print('$wl $tag FAILED', e)
PY
}
P=$PWD/ndarray_interp_b200
for wl in c5a c4 c4x c1; do
  EXTRA="" run stage $wl NDI_X=1
  EXTRA="" run nostage $wl NDI_STAGE_GUESS=0
done
EXTRA="" run pb15 c3 NDI_X=1
EXTRA="" run pb25 c3 NDI_B200_LIB=$P/libndi_v_pb25.so
EXTRA="" run pb24 c3 NDI_B200_LIB=$P/libndi_v_pb24.so
EXTRA="" run pb43 c3 NDI_B200_LIB=$P/libndi_v_pb43.so
EXTRA="" run pb15 c3d NDI_X=1
EXTRA="" run pb24 c3d NDI_B200_LIB=$P/libndi_v_pb24.so
python scripts/bench_spline_build.py c2 --levels 4 --bc Natural > ${T}_build_plain.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:rowsplit_reduce -s 8 -c 1 -o ${T}_reduce python scripts/bench_spline_build.py c2 --levels 4 --bc Natural > ${T}_cap_reduce.log 2>&1
python profiles/summarize_ncu.py ${T}_reduce.ncu-rep gpurun_out/r2_ncu_c2_rowsplit_reduce.txt "rowsplit_reduce_kernel<double>, 4096 x 1024 f64, 4 levels" > /dev/null 2>&1
rm -f ${T}_reduce.ncu-rep
cat gpurun_out/r2_ncu_c2_rowsplit_reduce.txt
