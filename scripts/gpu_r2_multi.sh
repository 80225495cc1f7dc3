#!/bin/bash
# round 2, multi-GPU call:  gpurun --gpus N -- 'bash scripts/gpu_r2_multi.sh N'
# (a) single-process fan-out tests (ndi_interp*_replicate, clone_to_device) -- need >= 2 devices
# (b) the default bench line under torchrun at N ranks (C2 weak-scaled headline; C5 strong-scaled at 2^28 queries;
#     column-sharded (4096, 131072) build + all-gather), every rank's sample checked against the oracle
N=${1:-2}
mkdir -p gpurun_out
T=gpurun_out/r2m${N}
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > ${T}_smi.txt 2>&1
timeout 900 python -m pytest tests/test_multi_device_gpu.py -m gpu -q -p no:cacheprovider > ${T}_pytest_multi.log 2>&1
echo "multi-device pytest rc=$?"; tail -4 ${T}_pytest_multi.log
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 > ${T}_bench.json 2> ${T}_bench.err ) 2> ${T}_bench.time
echo "bench rc=$?"; tail -4 ${T}_bench.time; tail -c 800 ${T}_bench.err
python - <<PY
import json
try:
    d = json.load(open('${T}_bench.json'))
    print('HEAD n_gpus=%d' % d['n_gpus'], d['config']['workload'][:30], 'value=%.4g ms=%.4f frac=%.3f e2e=%.4g check=%s route=%s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d.get('check'), d['run'].get('spline_route')))
    print('E2E', {k: (round(v, 2) if isinstance(v, float) else v) for k, v in d['e2e'].items() if 'GBps' in k or 'frac' in k}, 'numa', d['run']['numa_node'])
    for k, v in d['workloads'].items():
        if 'error' in v: print(k, 'ERROR', v['error'], v.get('trace')); continue
        print(k, {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ('ms_per_step', 'value', 'build_ms', 'allgather_ms', 'allgather_GBps_per_gpu', 'build_info', 'scaling', 'queries_per_gpu')},
              'frac=%s' % (v.get('roofline') or {}).get('frac'), 'check=%s' % (v.get('check') or {}), 'e2e=%s' % (v.get('e2e') or {}).get('value'))
except Exception as e:
    print('bench FAILED', e)
PY
