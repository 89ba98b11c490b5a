#!/bin/bash
# One GPU-box cycle: parity tests, short bench at two shard sizes and both modes, one ncu capture.
# usage (under gpurun): bash tools/gpu_cycle.sh <tag> [ncu]
tag=${1:-x}
python -m pytest tests -m gpu -x -q --durations=5 2>&1 | tail -14
for gb in 1024 128; do
  python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --global-batch $gb > gpurun_out/b_${tag}_$gb.json 2> gpurun_out/b_${tag}_$gb.err
  python - <<PY
import json
d=json.load(open("gpurun_out/b_${tag}_$gb.json")); r=d["roofline"]
print("noise", $gb, "img/s %.0f" % d["value"], {k: round(v,4) for k,v in r["kernel_ms"].items()}, "frac %.3f" % r["frac"], d["clocks"])
PY
done
python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --mode blobs > gpurun_out/b_${tag}_blobs.json 2>&1
python - <<PY
import json
d=json.load(open("gpurun_out/b_${tag}_blobs.json")); r=d["roofline"]
print("blobs 1024 img/s %.0f" % d["value"], {k: round(v,4) for k,v in r["kernel_ms"].items()}, "frac %.3f" % r["frac"])
PY
if [ "$2" = "ncu" ]; then
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:sdnet_peaks -s 3 -c 1 -o gpurun_out/prof_peaks_${tag} python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu.log 2>&1
  tail -1 gpurun_out/ncu.log
fi
