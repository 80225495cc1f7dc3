#!/bin/bash
# access-pattern ceiling of the thin-row kernels (scripts/gather_ceiling.cu) + SM<->L2 traffic of C3's kernel beside it
mkdir -p gpurun_out
[ -x scripts/gather_ceiling.bin ] || nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o scripts/gather_ceiling.bin scripts/gather_ceiling.cu
M=gpu__time_duration.sum,lts__t_sectors.sum,lts__t_sectors_srcunit_tex.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum,l1tex__m_xbar2l1tex_read_bytes.sum,l1tex__m_l1tex2xbar_write_bytes.sum,l1tex__data_pipe_lsu_wavefronts.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__cycles_elapsed.max
scripts/gather_ceiling.bin > gpurun_out/gather_ceiling.jsonl 2> gpurun_out/gather_ceiling.err; echo "ceiling rc=$?"; cat gpurun_out/gather_ceiling.jsonl
ncu --metrics $M --clock-control none -c 4 --csv --log-file gpurun_out/gather_ceiling_ncu.csv scripts/gather_ceiling.bin > /dev/null 2>&1
for WL in c3 c4 c5a; do
ncu --metrics $M --clock-control none -k regex:interp -s 4 -c 1 --csv --log-file gpurun_out/l2_$WL.csv python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > /dev/null 2>&1
done
python - <<'PY'
import csv,glob
for f in sorted(glob.glob('gpurun_out/gather_ceiling_ncu.csv')+glob.glob('gpurun_out/l2_*.csv')):
    rows=[r for r in csv.reader(open(f)) if len(r)>10]
    if not rows: continue
    hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
    out={}
    for r in rows[1:]:
        out.setdefault((r[ii],r[ki][:60]),{})[r[mi]]=r[vi]
    for k,v in out.items():
        print(f, k, {a.replace('lts__t_sectors','S').replace('.sum',''):b for a,b in v.items()})
PY
