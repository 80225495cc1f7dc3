#!/bin/bash
# round 2, call 10: band sweeps with the lean pair arithmetic and the FIFO ring / y prefetch; band size of the binned
# C5a batch now that tiles are handed out in order; bulk-copy (TMA) gather probe
mkdir -p gpurun_out
T=gpurun_out/r2c10
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider -k "bilinear or fuzz" > ${T}_pytest.log 2>&1
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
for mb in 24 32 48 64; do run mb$mb c4 NDI_SWEEP_MB=$mb; done
run mb32 c4x NDI_X=1
for mb in 8 16 32; do run band$mb c5a NDI_BAND_MB=$mb; done
M='regex:^l1tex__data_pipe_lsu_wavefronts(_mem_(lg|shared).*)?\.sum$,smsp__inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed_op_shared_ld.sum,smsp__inst_executed_op_shared_st.sum,smsp__inst_executed_op_global_ld.sum,smsp__inst_executed_op_global_st.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio'
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
./scripts/tma_gather_probe.bin 2>&1 | tee ${T}_tma_probe.jsonl
