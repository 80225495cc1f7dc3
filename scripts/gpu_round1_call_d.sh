#!/bin/bash
# store flavour of the output rows: st.global.cs (shipped) / plain / .wt / .cg
# variants: python -m ndarray_interp_b200.build --define NDI_STORE_MODE=1|2|3 --out libndi_v_stN.so
mkdir -p gpurun_out
LIBS="libndi_b200.so libndi_v_st1.so libndi_v_st2.so libndi_v_st3.so" WLS="c3 c4 c5b c2" bash scripts/gpu_ab_libs.sh
