#!/bin/bash
# round-1 evidence run: tests, every workload's bench line, the reference arm, launch list and ncu captures
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/final_pytest.log
for WL in c2 c1 c3 c3d c4 c4x c5a c5b; do
  timeout 900 python bench.py --workload $WL --steps 30 --warmup 5 > gpurun_out/final_$WL.json 2> gpurun_out/final_$WL.err || tail -c 300 gpurun_out/final_$WL.err
  python -c "
import json
d=json.load(open('gpurun_out/final_$WL.json')); c=d.get('cpu_baseline') or {}
print('$WL ms=%.4f q/s=%.3g frac=%.3f e2e=%.3g cpu=%.3g (%s thr) launches/step=%s'%(d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value'], c.get('value',0), c.get('cores'), d['config'].get('launches_per_step')))"
done
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/final_reference_c2.json 2>/dev/null; tail -c 400 gpurun_out/final_reference_c2.json; echo
# launch list of the default bench command
python bench.py --steps 2 --warmup 3 > gpurun_out/final_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches_c2.csv python bench.py --steps 2 --warmup 3 > gpurun_out/final_ncu_launches.log 2>&1
# full captures of the dominant kernels
cap() { # name workload kernel-regex skip env...
  local name=$1 wl=$2 re=$3 skip=$4; shift 4
  env "$@" ncu --set full --import-source on --clock-control none -k regex:$re -s $skip -c 1 -o gpurun_out/final_$name python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/final_cap_$name.log 2>&1
  # the reports are 20 MB each and gpurun brings back 64 MB at most: keep the summary, drop the report
  python profiles/summarize_ncu.py gpurun_out/final_$name.ncu-rep gpurun_out/ncu_$name.txt "$name: bench.py --workload $wl, kernel $re" > /dev/null 2>&1
  rm -f gpurun_out/final_$name.ncu-rep
}
cap c2_cubic c2 interp1d_cubic 4 X=1
cap c3_linear c3 interp1d_linear 4 X=1
cap c4_bilinear c4 interp2d_bilinear 4 X=1
cap c5a_bilinear_binned c5a interp2d_bilinear 4 X=1
cap c5a_bin_scatter c5a bin_scatter 4 X=1
cap c5b_cubic c5b interp1d_cubic 4 X=1
timeout 300 python scripts/bench_spline_build.py > gpurun_out/final_spline_build.jsonl 2>&1
ls -la gpurun_out/ncu_*.txt
