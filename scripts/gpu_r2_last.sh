#!/bin/bash
# last call of round 2: the whole GPU suite on the final commit, smoke, the default bench line and the reference arm
mkdir -p gpurun_out/final
timeout 2400 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > gpurun_out/final/pytest_last.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/final/pytest_last.log
bash scripts/gpu_r2_bench_n1.sh
