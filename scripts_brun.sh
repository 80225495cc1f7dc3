for w in c3 c4 c5a c5b c2; do for m in 0 1 2; do
timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu --e2e-steps 1 --search-mode $m > gpurun_out/b3_${w}_$m.json 2> gpurun_out/b3_${w}_$m.err || tail -c 300 gpurun_out/b3_${w}_$m.err
python -c "
import json,sys
d=json.load(open('gpurun_out/b3_${w}_$m.json'))
print('$w mode $m', 'ms=%.4f'%d['ms_per_step'], 'GB/s=%.0f'%d['roofline']['achieved'], 'frac=%.3f'%d['roofline']['frac'])
"; done; done
