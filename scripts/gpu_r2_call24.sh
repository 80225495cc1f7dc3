#!/bin/bash
# round 2, call 24: NDI_BUILD_AUTO keeps the reference order where a right NotAKnot row meets a grid that makes the
# reference's system nearly singular -- spline tests, fuzz, mirror, smoke, and the randomised sweep under AUTO
mkdir -p gpurun_out
T=gpurun_out/r2c24
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python -m pytest tests/test_partition_gpu.py tests/test_rowsplit_gpu.py tests/test_parity_spline_gpu.py tests/test_fuzz_gpu.py tests/test_cpp_mirror.py tests/test_reference_cubic_spline.py tests/test_fullsize_gpu.py tests/test_threads_gpu.py -m gpu -q --maxfail=20 -p no:cacheprovider > ${T}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 ${T}_pytest.log
timeout 900 python scripts/fuzz_partition.py 1500 21 auto 2>&1 | tail -6
