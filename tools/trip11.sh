#!/bin/bash
o=gpurun_out
for occ in 0 4 3 2; do for gb in 128 256; do for pipe in 2 4 6; do
SDNET_PEAKS_OCC=$occ python bench.py --global-batch $gb --pipeline $pipe --no-e2e --no-cpu-baseline --no-objects --no-parity --steps 200 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('occ $occ gb $gb pipe $pipe', {m: (round(r['value']), round(r['ms_per_step'],4), round(r['kernel_ms']['peaks'],4)) for m,r in d['modes'].items()})"
done; done; done
