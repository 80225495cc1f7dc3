#!/bin/bash
# round 2, call 9: band sweeps (ndi_sweep.cu) -- parity suite, C4 / C4x with the sweeps off and at several band sizes,
# and what one warp-wide memory instruction costs on the L1TEX data pipe (scripts/wavefront_probe.cu)
mkdir -p gpurun_out
T=gpurun_out/r2c9
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > ${T}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 ${T}_pytest.log
run() {  # tag workload env...
  local tag=$1 wl=$2; shift 2
  env "$@" timeout 600 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu --e2e-steps 1 > ${T}_${wl}_$tag.json 2> ${T}_${wl}_$tag.err || tail -c 400 ${T}_${wl}_$tag.err
  python - <<PY
import json
try:
    d = json.load(open('${T}_${wl}_$tag.json'))
    print('$wl $tag ms=%.4f frac=%.3f median=%.4f best=%.4f check=%s e2e=%.4g' % (d['ms_per_step'], d['roofline']['frac'], d['per_step']['median_ms'], d['per_step']['best_ms'], (d.get('check') or {}).get('bit_exact'), d['e2e']['value']))
except Exception as e:
    print('$wl $tag FAILED', e)
PY
}
P=$PWD/ndarray_interp_b200
run off c4 NDI_SWEEP_MODE=0
for mb in 16 24 32 48 64; do run mb$mb c4 NDI_SWEEP_MB=$mb; done
run nostage c4 NDI_SWEEP_STAGE_X=0
run min3 c4 NDI_B200_LIB=$P/libndi_v_sw3.so
run min5 c4 NDI_B200_LIB=$P/libndi_v_sw5.so
run off c4x NDI_SWEEP_MODE=0
run mb32 c4x NDI_X=1
# per-kernel counters of the swept C4 launch (the 2^24-query launches are the long ones)
M='regex:^l1tex__data_pipe_lsu_wavefronts(_mem_(lg|shared).*)?\.sum$,smsp__inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed_op_shared_ld.sum,smsp__inst_executed_op_shared_st.sum,smsp__inst_executed_op_global_ld.sum,smsp__inst_executed_op_global_st.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active'
ncu --metrics "$M" --clock-control none -k regex:'sweep' -c 8 --csv --log-file ${T}_wf_c4.csv \
    python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > ${T}_wf_c4.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('${T}_wf_c4.csv')) if len(r) > 5]
h=rows[0]; ki,mi,vi,ii=h.index('Kernel Name'),h.index('Metric Name'),h.index('Metric Value'),h.index('ID')
d={}
for r in rows[1:]:
    d.setdefault((r[ii],r[ki][:50]),{})[r[mi]]=r[vi]
best=max(d.items(), key=lambda kv: float(kv[1]['gpu__time_duration.sum'].replace(',','')))
print('c4 swept', best[0])
for a,b in sorted(best[1].items()): print('     %-75s %s' % (a,b))
PY
# wavefronts per instruction, per access pattern
M2='regex:^l1tex__data_pipe_lsu_wavefronts(_mem_(lg|shared))?\.sum$,smsp__inst_executed_op_shared_ld.sum,smsp__inst_executed_op_shared_st.sum,smsp__inst_executed_op_global_ld.sum,smsp__inst_executed_op_global_st.sum,smsp__inst_executed_op_shfl.sum,gpu__time_duration.sum,lts__t_sectors.sum'
./scripts/wavefront_probe.bin > ${T}_probe_plain.log 2>&1 && \
ncu --metrics "$M2" --clock-control none --csv --log-file ${T}_probe.csv ./scripts/wavefront_probe.bin > ${T}_probe.log 2>&1
python - <<PY
import csv, re
rows=[r for r in csv.reader(open('${T}_probe.csv')) if len(r) > 5]
h=rows[0]; ki,mi,vi,ii=h.index('Kernel Name'),h.index('Metric Name'),h.index('Metric Value'),h.index('ID')
names=[l.split()[0] for l in open('${T}_probe_plain.log') if l.strip()]
d={}
for r in rows[1:]:
    d.setdefault(int(r[ii]),{})[r[mi]]=float(r[vi].replace(',',''))
print('%-22s %10s %10s %10s %10s %8s' % ('pattern','wf/instr','lg','shared','sectors/i','us'))
for i,v in sorted(d.items()):
    n = v.get('smsp__inst_executed_op_global_ld.sum',0)+v.get('smsp__inst_executed_op_global_st.sum',0)+v.get('smsp__inst_executed_op_shared_ld.sum',0)+v.get('smsp__inst_executed_op_shared_st.sum',0)
    base = 148*4*8*4   # the staging stores of the 16 KB block (STS.32 x 16 per thread)
    if n == 0 or 'SHFL' in names[i]: n = 148*4*8*256
    print('%-22s %10.2f %10.2f %10.2f %10.2f %8.1f' % (names[i] if i < len(names) else i, v['l1tex__data_pipe_lsu_wavefronts.sum']/n, v['l1tex__data_pipe_lsu_wavefronts_mem_lg.sum']/n if 'l1tex__data_pipe_lsu_wavefronts_mem_lg.sum' in v else -1, v.get('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',0)/n, v['lts__t_sectors.sum']/n, v['gpu__time_duration.sum']/1000))
PY
