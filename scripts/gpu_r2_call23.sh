#!/bin/bash
# round 2, call 23: top solve 32 columns per block on wide tables; rows per item of the rhs / a-b kernels (A/B: 2, 4, 8)
mkdir -p gpurun_out
T=gpurun_out/r2c23
timeout 900 python -m pytest tests/test_partition_gpu.py tests/test_rowsplit_gpu.py tests/test_parity_spline_gpu.py tests/test_fuzz_gpu.py -m gpu -q --maxfail=20 -p no:cacheprovider -k "spline or partition or rowsplit" > ${T}_pytest.log 2>&1
echo "pytest rc=$?"; tail -2 ${T}_pytest.log
P=$PWD/ndarray_interp_b200
for v in default rg8 rg2; do
  if [ $v = default ]; then unset NDI_B200_LIB; else export NDI_B200_LIB=$P/libndi_v_$v.so; fi
  echo "--- $v"
  python scripts/bench_spline_build.py --levels 0 --blocks 0 --bc Natural 2> ${T}_$v.err | python -c "
import sys, json
for ln in sys.stdin:
    d = json.loads(ln)
    print('%-10s %-10s %4d  %.4f ms' % (d['shape'], d['mode'], d['levels'], d['ms']))"
done
unset NDI_B200_LIB
NDI_B200_LIB=$P/libndi_v_rg8.so timeout 600 python -m pytest tests/test_partition_gpu.py tests/test_parity_spline_gpu.py -m gpu -q --maxfail=5 -p no:cacheprovider > ${T}_pytest_rg8.log 2>&1
echo "pytest rg8 rc=$?"; tail -2 ${T}_pytest_rg8.log
