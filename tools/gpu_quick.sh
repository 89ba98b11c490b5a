#!/bin/bash
# quick bench-only cycle: usage bash tools/gpu_quick.sh <tag> [env assignments...]
tag=$1; shift
for gb in 1024 128; do
  env "$@" python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --global-batch $gb > gpurun_out/q_${tag}_$gb.json 2> gpurun_out/q_${tag}_$gb.err
  python - <<PY
import json
d=json.load(open("gpurun_out/q_${tag}_$gb.json")); r=d["roofline"]
print("$tag noise", $gb, "img/s %.0f" % d["value"], {k: round(v,4) for k,v in r["kernel_ms"].items()}, "frac %.3f" % r["frac"], "cand/plane %.0f" % d["detections"]["candidates_per_plane_mean"])
PY
done
