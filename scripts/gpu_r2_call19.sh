#!/bin/bash
# round 2, call 19: unsigned element types -- the parity / fuzz / mirror tests again after the test fixes
mkdir -p gpurun_out
T=gpurun_out/r2c19
timeout 1200 python -m pytest tests/test_parity_gpu.py tests/test_fuzz_gpu.py tests/test_cpp_mirror.py -m gpu -q --maxfail=30 -p no:cacheprovider > ${T}_pytest.log 2>&1
echo "pytest rc=$?"; tail -12 ${T}_pytest.log
