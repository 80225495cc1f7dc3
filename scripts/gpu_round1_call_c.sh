#!/bin/bash
# probe-ahead variants of the thin-row linear kernel: parity tests with the variant, then A/B
# variants: the probe-ahead code (NDI_PROBE_AHEAD) was measured with this call and dropped without being committed; what it did and what it measured is in profiles/r01/shuffle_packing.md and in the comment above interp1d_linear_kernel
mkdir -p gpurun_out
NDI_B200_LIB=$PWD/ndarray_interp_b200/libndi_v_pa.so timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_fuzz_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q > gpurun_out/c_pytest_pa.log 2>&1; echo "pytest pa rc=$?"; tail -3 gpurun_out/c_pytest_pa.log
LIBS="libndi_b200.so libndi_v_pa.so libndi_v_pa6.so libndi_v_pa16.so" WLS="c3 c3d" bash scripts/gpu_ab_libs.sh
LIBS="libndi_b200.so libndi_v_pa.so" WLS="c3" bash scripts/gpu_ab_libs.sh
