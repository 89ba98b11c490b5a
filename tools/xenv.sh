#!/bin/bash
# Time the product library under env-knob settings: bash tools/xenv.sh "A=1 B=2" "C=3" ...  ("" = defaults)
extra=${XENV_ARGS:-}
for e in "$@"; do
  env $e python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline $extra 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('[$e]', round(d['value']), {k: round(v,4) for k,v in d['roofline']['kernel_ms'].items()})"
done
