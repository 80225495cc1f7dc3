#!/bin/bash
# round 2, final evidence call (one GPU): the whole GPU suite, ncu captures of every workload's kernels and of the
# partition build's kernels (summaries only leave the box), the traffic stamp, the default bench line (AFTER the stamp,
# so that its roofline.traffic is current), spline build wall times and the launch list of the bench's headline steps.
#   gpurun --timeout 3000 -- 'bash scripts/gpu_r2_final.sh <commit>'
COMMIT=${1:-unknown}
mkdir -p gpurun_out/final
T=gpurun_out/final
timeout 2400 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > ${T}/pytest.log 2>&1
echo "pytest rc=$?"; tail -3 ${T}/pytest.log
cap() { # name workload kernel-regex skip
  local name=$1 wl=$2 re=$3 skip=$4
  python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > ${T}/cap_plain_$name.log 2>&1 && \
  ncu --set full --import-source on --clock-control none -k regex:$re -s $skip -c 1 -o ${T}/$name python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > ${T}/cap_$name.log 2>&1
  python profiles/summarize_ncu.py ${T}/$name.ncu-rep ${T}/ncu_$name.txt "$name: bench.py --workload $wl, kernel $re (commit $COMMIT)" > /dev/null 2>&1
  rm -f ${T}/$name.ncu-rep
  grep -E "^kernel|gpu__time_duration|dram__bytes|lsu_wavefronts|issue_active" ${T}/ncu_$name.txt
}
cap c5a_bilinear_binned c5a interp2d_bilinear 4
cap c5a_bin_scatter c5a bin_scatter 4
cap c5a_bin_totals c5a bin_totals 4
cap c4_bilinear c4 interp2d_bilinear 4
cap c2_cubic c2 interp1d_cubic 4
cap c5b_cubic c5b interp1d_cubic 4
cap c3_linear_pair c3 interp1d_linear 4
capb() { # name kernel-regex skip : partition build kernels at the C2 shape
  local name=$1 re=$2 skip=$3
  ncu --set full --import-source on --clock-control none -k regex:$re -s $skip -c 1 -o ${T}/$name python scripts/bench_spline_build.py c2 --levels 0 --blocks 0 --bc Natural > ${T}/cap_$name.log 2>&1
  python profiles/summarize_ncu.py ${T}/$name.ncu-rep ${T}/ncu_$name.txt "$name: scripts/bench_spline_build.py c2 (4096 x 1024 f64, Natural), kernel $re (commit $COMMIT)" > /dev/null 2>&1
  rm -f ${T}/$name.ncu-rep
  grep -E "^kernel|gpu__time_duration|dram__bytes|registers_per_thread|warps_active|issue_active" ${T}/ncu_$name.txt
}
python scripts/bench_spline_build.py c2 --levels 0 --blocks 0 --bc Natural > ${T}/build_plain.log 2>&1
capb c2_part_local part_local 8
capb c2_part_ab part_ab 8
capb c2_spline_rhs spline_rhs 24
capb c2_part_top part_top_kernel 8
python profiles/stamp_traffic.py ${T} $COMMIT > ${T}/stamp.log 2>&1; tail -3 ${T}/stamp.log
cp profiles/roofline_traffic.json ${T}/roofline_traffic.json
( time timeout 1200 python bench.py > ${T}/bench_n1.json 2> ${T}/bench_n1.err ) 2> ${T}/bench_n1.time; echo "bench rc=$?"; tail -3 ${T}/bench_n1.time; tail -c 300 ${T}/bench_n1.err
( time timeout 600 python bench.py --impl reference > ${T}/bench_reference.json 2> ${T}/bench_reference.err ) 2> ${T}/bench_reference.time; echo "reference arm rc=$?"; tail -3 ${T}/bench_reference.time
python - <<'PY'
import json
d = json.load(open('gpurun_out/final/bench_n1.json'))
print('C2 value %.4g q/s  ms %.4f  frac %.3f  traffic current %s  e2e %.4g  cpu %.4g' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['traffic_capture']['current'], d['e2e']['value'], d['cpu_baseline']['value']))
sb = d['spline_build']
print('build:', {k: (round(v['ms'], 4), v['levels']) for k, v in sb.items() if isinstance(v, dict)}, 'auto', sb['auto'], 'frac', round(sb['frac_of_hbm_peak'], 4))
for k, v in d['workloads'].items():
    print(k, 'ms', v.get('ms_per_step') or v.get('build_ms'), 'frac', (v.get('roofline') or {}).get('frac') or v.get('frac_of_hbm_peak'), 'check', (v.get('check') or {}).get('bit_exact'), 'e2e', (v.get('e2e') or {}).get('value'), 'cpu', (v.get('cpu_baseline') or {}).get('value'))
r = json.load(open('gpurun_out/final/bench_reference.json'))
print('reference arm: value %.4g %s cores %s' % (r['value'], r['unit'], r['cpu_baseline']['cores']))
PY
python scripts/bench_spline_build.py --levels 0 --blocks 0 --bc Natural,NotAKnot,Periodic,Individual > ${T}/spline_build.jsonl 2> ${T}/spline_build.err
python bench.py --steps 2 --warmup 1 --workload c2 --no-cpu --e2e-steps 1 > ${T}/launch_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${T}/launches_c2_bench.csv python bench.py --steps 2 --warmup 1 --workload c2 --no-cpu --e2e-steps 1 > ${T}/launch_ncu.log 2>&1
wc -l ${T}/launches_c2_bench.csv
