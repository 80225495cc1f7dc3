#!/bin/bash
# tests with the shipped library, parity tests with the NDI_PACK_SHFL=2 build, then the A/B of the packing variants
# variants: python -m ndarray_interp_b200.build --define NDI_PACK_SHFL=0|1|2 --out libndi_v_packN.so (the shipped default is now 1)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/a_pytest.log
NDI_B200_LIB=$PWD/ndarray_interp_b200/libndi_v_pack2.so timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_parity_spline_gpu.py tests/test_fuzz_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q > gpurun_out/a_pytest_pack2.log 2>&1; echo "pytest pack2 rc=$?"; tail -3 gpurun_out/a_pytest_pack2.log
LIBS="libndi_b200.so libndi_v_pack1.so libndi_v_pack2.so" WLS="c3 c3d c4 c5a c5b" bash scripts/gpu_ab_libs.sh
