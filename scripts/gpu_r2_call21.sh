#!/bin/bash
# round 2, call 21: binning passes request their next chunk's queries one chunk ahead -- bilinear parity tests, C5a / C4 / C4x
# timings, ncu captures of the bilinear workloads' kernels with this library, traffic stamp
COMMIT=${1:-unknown}
mkdir -p gpurun_out/c21
T=gpurun_out/c21
timeout 1500 python -m pytest tests/test_parity_gpu.py tests/test_fuzz_gpu.py tests/test_fullsize_gpu.py tests/test_reference_interp2d.py tests/test_multi_device_gpu.py -m gpu -q --maxfail=20 -p no:cacheprovider > ${T}/pytest.log 2>&1
echo "pytest rc=$?"; tail -3 ${T}/pytest.log
for wl in c5a c4 c4x; do
  timeout 600 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu --e2e-steps 1 > ${T}/$wl.json 2> ${T}/$wl.err || tail -c 300 ${T}/$wl.err
  python -c "
import json
d = json.load(open('${T}/$wl.json')); print('$wl ms=%.4f frac=%.3f median=%.4f check=%s' % (d['ms_per_step'], d['roofline']['frac'], d['per_step']['median_ms'], d['check']['bit_exact']))"
done
cap() { # name workload kernel-regex skip
  local name=$1 wl=$2 re=$3 skip=$4
  ncu --set full --import-source on --clock-control none -k regex:$re -s $skip -c 1 -o ${T}/$name python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > ${T}/cap_$name.log 2>&1
  python profiles/summarize_ncu.py ${T}/$name.ncu-rep ${T}/ncu_$name.txt "$name: bench.py --workload $wl, kernel $re (commit $COMMIT)" > /dev/null 2>&1
  rm -f ${T}/$name.ncu-rep
  grep -E "^kernel|gpu__time_duration|dram__bytes|lsu_wavefronts|issue_active|long_scoreboard|barrier" ${T}/ncu_$name.txt
}
cap c5a_bin_scatter c5a bin_scatter 4
cap c5a_bin_totals c5a bin_totals 4
cap c5a_bilinear_binned c5a interp2d_bilinear 4
cap c4_bilinear c4 interp2d_bilinear 4
python profiles/stamp_traffic.py ${T} $COMMIT > ${T}/stamp.log 2>&1; tail -3 ${T}/stamp.log
cp profiles/roofline_traffic.json ${T}/roofline_traffic.json
