#!/bin/bash
# parity tests + short benches over the configurations tracked in profiles/: bash tools/gpu_all.sh <tag>
tag=${1:-x}
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/t_$tag.log 2>&1; tail -4 gpurun_out/t_$tag.log | cut -c1-200
run() { name=$1; shift
  python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline "$@" > gpurun_out/a_${tag}_$name.json 2> gpurun_out/a_${tag}_$name.err
  python -c "
import json;d=json.load(open('gpurun_out/a_${tag}_$name.json'));r=d['roofline']
print('$name', round(d['value']), {k: round(v,4) for k,v in r['kernel_ms'].items()}, 'frac %.3f' % r['frac'], 'cand %.0f' % d['detections']['candidates_per_plane_mean'])"
}
run noise; run blobs --mode blobs; run gb128 --global-batch 128; run f16 --dtype f16; run bf16 --dtype bf16; run f16blobs --dtype f16 --mode blobs
