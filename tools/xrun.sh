#!/bin/bash
# Time experiment builds: bash tools/xrun.sh <name> [<name> ...]   ("base" = the product library)
for n in "$@"; do
  lib=structuredetector_b200/csrc/exp/lib_$n.so; [ "$n" = base ] && lib=
  SDNET_DECODE_LIB=$lib python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$n', round(d['value']), {k: round(v,4) for k,v in d['roofline']['kernel_ms'].items()})"
done
