#!/bin/bash
for st in 10 20 30 60; do python bench.py --workload c5a --steps $st --warmup 5 --no-cpu --e2e-steps 1 > gpurun_out/var_c5a_$st.json 2>/dev/null; python -c "
import json
d=json.load(open('gpurun_out/var_c5a_$st.json')); print('c5a steps=$st ms=%.4f frac=%.3f clocks=%s'%(d['ms_per_step'], d['roofline']['frac'], d['clocks']))"; done
nvidia-smi --query-gpu=power.draw,power.limit,clocks.sm,clocks.mem,temperature.gpu --format=csv
