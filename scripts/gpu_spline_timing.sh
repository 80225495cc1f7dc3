#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "spline or cubic or coeff or periodic" > gpurun_out/t_spline.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_spline.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:spline --csv --log-file gpurun_out/spline_launches.csv python scripts/bench_spline_build.py > gpurun_out/spline_ncu.log 2>&1; python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/spline_launches.csv")))
h=[i for i,r in enumerate(rows) if "Kernel Name" in r][0]
hd=rows[h]; ki,vi,gi,bi=hd.index("Kernel Name"),hd.index("Metric Value"),hd.index("Grid Size"),hd.index("Block Size")
seen={}
for r in rows[h+1:]:
    key=(r[ki][:44],r[gi],r[bi]); seen.setdefault(key,[]).append(float(r[vi].replace(",","")))
for k,v in seen.items(): print(k, len(v), "min %.1f us  max %.1f us"%(min(v)/1e3,max(v)/1e3))
PY
