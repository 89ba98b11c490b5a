#!/bin/bash
o=gpurun_out
timeout 900 python bench.py > $o/r02h_bench.json 2> $o/r02h_bench.err || tail -20 $o/r02h_bench.err
python -c "
import json;d=json.load(open('$o/r02h_bench.json'))
print('value', round(d['value']), d['ms_per_step'], 'parity', d['parity_checked'], 'e2e', (d['e2e'] or {}).get('value'), 'e2e_obj', (d['e2e_objects'] or {}).get('value'), 'obj', (d['python_objects'] or {}).get('value'))
for m,r in d['modes'].items(): print(m, round(r['value']), r['kernel_ms'], round(r['peaks_frac'],3), r['parity']['ok'], r['detections'])
print(d['cpu_baseline']); print(d['clocks'])
"
for p in 1 2 3; do python bench.py --pipeline $p --no-e2e --no-cpu-baseline --no-objects --no-parity --steps 100 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('pipeline $p', {m: (round(r['value']), round(r['ms_per_step'],4)) for m,r in d['modes'].items()})"; done
for gb in 512 256 128; do python bench.py --global-batch $gb --no-e2e --no-cpu-baseline --no-objects --steps 100 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('gb $gb', d['parity_checked'], {m: (round(r['value']), round(r['ms_per_step'],4), round(r['kernel_ms']['peaks'],4)) for m,r in d['modes'].items()})"; done
