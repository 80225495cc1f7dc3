#!/bin/bash
# round 2, call 1: the whole GPU test-suite on the new kernels, then A/B of the evaluation-kernel changes
# (packed f32x2 arithmetic, shared-memory record broadcast, pair table, merge search) and the spline build modes.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2c1_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/r2c1_pytest.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/r2c1_pytest.log
run() {  # tag workload [env...] -- bench args
  local tag=$1 wl=$2; shift 2
  env "$@" timeout 600 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu --e2e-steps 1 $EXTRA > gpurun_out/r2c1_${wl}_$tag.json 2> gpurun_out/r2c1_${wl}_$tag.err || tail -c 400 gpurun_out/r2c1_${wl}_$tag.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r2c1_${wl}_$tag.json'))
    print('$wl $tag ms=%.4f frac=%.3f median=%.4f best=%.4f' % (d['ms_per_step'], d['roofline']['frac'], d['per_step']['median_ms'], d['per_step']['best_ms']), d.get('spline_build'))
except Exception as e:
    print('$wl $tag FAILED', e)
PY
}
P=$PWD/ndarray_interp_b200
for wl in c3 c4 c5a c5b c2 c1; do
  EXTRA="" run new $wl NDI_X=1
  EXTRA="" run nof2 $wl NDI_B200_LIB=$P/libndi_v_nof2.so
  EXTRA="" run nobc $wl NDI_B200_LIB=$P/libndi_v_nobc.so
  EXTRA="" run old $wl NDI_B200_LIB=$P/libndi_v_old.so NDI_PAIR_TABLE=0
done
EXTRA="" run nopair c3 NDI_PAIR_TABLE=0
EXTRA="" run nopair c3d NDI_PAIR_TABLE=0
EXTRA="" run new c3d NDI_X=1
EXTRA="--search-mode 5" run merge c2 NDI_X=1
EXTRA="--search-mode 5" run merge c5b NDI_X=1
EXTRA="--search-mode 4" run lut c2 NDI_X=1
EXTRA="" run seqbuild c2 NDI_BUILD_MODE=1
python scripts/bench_spline_build.py c2 long --levels 0,2,3,4,5,6 --bc Natural,Periodic > gpurun_out/r2c1_spline_build.jsonl 2> gpurun_out/r2c1_spline_build.err
python scripts/bench_spline_build.py c5b-shard wide c2 long --levels 0 --bc NotAKnot,Individual >> gpurun_out/r2c1_spline_build.jsonl 2>> gpurun_out/r2c1_spline_build.err
cat gpurun_out/r2c1_spline_build.jsonl
tail -3 gpurun_out/r2c1_spline_build.err
