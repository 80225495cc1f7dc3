#!/bin/bash
# round 2, call 18: unsigned element types (u32, u64) -- the whole GPU suite
mkdir -p gpurun_out
T=gpurun_out/r2c18
timeout 2400 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > ${T}_pytest.log 2>&1
echo "pytest rc=$?"; tail -25 ${T}_pytest.log
