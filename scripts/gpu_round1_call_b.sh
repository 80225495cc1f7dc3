#!/bin/bash
# tests with the one-point bucket entries, then A/B against the previous table format and other table sizes
# variants: libndi_v_oldlut.so = the library of commit dacadee (previous table format); --define NDI_LUT_PER_POINT=4|16 --out libndi_v_lutN.so
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/b_pytest.log
LIBS="libndi_v_oldlut.so libndi_b200.so libndi_v_lut4.so libndi_v_lut16.so" WLS="c3 c3d c4 c5a c5b c1" bash scripts/gpu_ab_libs.sh
