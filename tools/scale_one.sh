#!/bin/bash
# usage (under gpurun --gpus N): bash tools/scale_one.sh <tag> <N>   -> gpurun_out/<tag>_bench_n<N>.json (full global batch 1024)
tag=$1; n=$2
if [ $n -eq 1 ]; then
  python bench.py --gpus 1 --no-cpu-baseline --no-e2e --no-objects > gpurun_out/${tag}_bench_n$n.json 2> gpurun_out/${tag}_bench_n$n.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --no-e2e --no-cpu-baseline --no-objects > gpurun_out/${tag}_bench_n$n.json 2> gpurun_out/${tag}_bench_n$n.err
fi
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${tag}_bench_n$n.json")); print($n, round(d["value"]), round(d["ms_per_step"],4), "parity", d["parity_checked"], "gather", d["gather_bit_exact"], {m: (round(r["value"]), round(r["ms_per_step"],4), r["ms_per_step_by_rank"]) for m,r in d["modes"].items()}, d["clocks"], d["config"]["gather"][:60])
except Exception as e:
    print($n, "failed", e); print(open("gpurun_out/${tag}_bench_n$n.err").read()[-800:])
PY
