#!/bin/bash
# round 2, call 8: L1TEX wavefront breakdown (global reads / writes / shared / shuffles) of the thin-row kernels as built
mkdir -p gpurun_out
T=gpurun_out/r2c8
M='regex:^l1tex__data_pipe_lsu_wavefronts(_mem_(lg|shared).*)?\.sum$,smsp__inst_executed.sum,gpu__time_duration.sum,smsp__inst_executed_op_shared_ld.sum,smsp__inst_executed_op_shared_st.sum,smsp__inst_executed_op_global_ld.sum,smsp__inst_executed_op_global_st.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum'
for wl in c5a c4 c3; do
  python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > ${T}_${wl}_plain.log 2>&1 || { echo "$wl plain FAILED"; tail -3 ${T}_${wl}_plain.log; continue; }
  ncu --metrics "$M" --clock-control none -k regex:'bin_|bilinear|linear' -s 6 -c 3 --csv --log-file ${T}_wf_$wl.csv \
    python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > ${T}_wf_$wl.log 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('${T}_wf_$wl.csv')) if len(r) > 5]
h=rows[0]; ki,mi,vi,ii=h.index('Kernel Name'),h.index('Metric Name'),h.index('Metric Value'),h.index('ID')
d={}
for r in rows[1:]:
    d.setdefault((r[ii],r[ki][:50]),{})[r[mi]]=r[vi]
for k,v in d.items():
    print('$wl',k)
    for a,b in sorted(v.items()): print('     %-75s %s' % (a,b))
PY
done
