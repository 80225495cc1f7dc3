#!/bin/bash
# A/B several builds of the library on a list of workloads:  LIBS="a.so b.so" WLS="c4 c5a" bash scripts/gpu_ab_libs.sh
mkdir -p gpurun_out
for WL in $WLS; do for lib in $LIBS; do
  tag=$(basename $lib .so)
  NDI_B200_LIB=$PWD/ndarray_interp_b200/$lib timeout 600 python bench.py --workload $WL --steps 20 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/ab_${WL}_$tag.json 2> gpurun_out/ab_${WL}_$tag.err || tail -c 300 gpurun_out/ab_${WL}_$tag.err
  python -c "
import json
d=json.load(open('gpurun_out/ab_${WL}_$tag.json')); print('$WL $tag ms=%.4f frac=%.3f'%(d['ms_per_step'], d['roofline']['frac']))"
done; done
