#!/bin/bash
# usage (under gpurun --gpus G): bash tools/scale_run.sh <tag> <maxN>
tag=$1; maxn=${2:-8}
for n in 1 2 4 8; do
  [ $n -gt $maxn ] && break
  if [ $n -eq 1 ]; then
    python bench.py --gpus 1 --steps 200 --warmup 20 --no-cpu-baseline --no-e2e > gpurun_out/scale_${tag}_n$n.json 2> gpurun_out/scale_${tag}_n$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 200 --warmup 20 --no-e2e > gpurun_out/scale_${tag}_n$n.json 2> gpurun_out/scale_${tag}_n$n.err
  fi
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/scale_${tag}_n$n.json")); print($n, round(d["value"]), round(d["ms_per_step"],4), {k: round(v,4) for k,v in d["roofline"]["kernel_ms"].items()}, d["clocks"])
except Exception as e:
    print($n, "failed", e); print(open("gpurun_out/scale_${tag}_n$n.err").read()[-800:])
PY
done
