#!/bin/bash
# GPU trip: parity suite, the bench line, knob sweep
o=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=15 --durations=8 > $o/r02a_pytest.log 2>&1; tail -25 $o/r02a_pytest.log
timeout 900 python bench.py --steps 50 --warmup 5 > $o/r02a_bench.json 2> $o/r02a_bench.err || tail -20 $o/r02a_bench.err
python -c "
import json;d=json.load(open('$o/r02a_bench.json'))
print('value', round(d['value']), d['ms_per_step'], 'parity', d['parity_checked'], 'e2e', (d['e2e'] or {}).get('value'), 'e2e_obj', (d['e2e_objects'] or {}).get('value'), 'obj', (d['python_objects'] or {}).get('value'))
for m,r in d['modes'].items(): print(m, round(r['value']), r['kernel_ms'], round(r['peaks_frac'],3), r['parity'], r['detections'])
print(d['cpu_baseline']); print(d['roofline']['schedule'])
"
timeout 1500 python tools/sweep.py base w0 w100 gr32 f32gr ng3c7 ng3c6 ng3c7gr ng3c7grw0 --out $o/r02a_sweep.json 2>&1 | tee $o/r02a_sweep.log | tail -40
