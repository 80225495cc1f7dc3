#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/t_spline.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_spline.log
timeout 600 python scripts/bench_spline_build.py 2>&1 | tee gpurun_out/spline_build.jsonl
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; cat gpurun_out/bench_c2.json
for WL in c4 c5a; do timeout 600 python bench.py --workload $WL --steps 20 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/auto_$WL.json 2>gpurun_out/auto_$WL.err; python -c "
import json; d=json.load(open('gpurun_out/auto_$WL.json')); print('$WL auto ms=%.4f frac=%.3f launches=%d'%(d['ms_per_step'], d['roofline']['frac'], d['gpu_launches']))"; done
