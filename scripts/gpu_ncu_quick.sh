#!/bin/bash
# per-kernel time + DRAM bytes + L2 hit rate for one bench workload:  WL=c4 ENVV="NDI_BIN_MODE=2" bash scripts/gpu_ncu_quick.sh tag
tag=$1
env $ENVV ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors.sum,sm__warps_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:'bin_|bilinear|linear|cubic|spline' -s 9 -c 6 --csv --log-file gpurun_out/q_${WL}_${tag}.csv \
  python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/q_${WL}_${tag}.log 2>&1
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/q_${WL}_${tag}.csv')))
hdr=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hdr]; ki,mi,vi,ii=h.index('Kernel Name'),h.index('Metric Name'),h.index('Metric Value'),h.index('ID')
d={}
for r in rows[hdr+1:]:
    d.setdefault((r[ii],r[ki][:60]),{})[r[mi]]=r[vi]
for k,v in d.items(): print('$WL $tag',k, {a.split('.')[0].replace('__','_')[-22:]:b for a,b in v.items()})
PY
