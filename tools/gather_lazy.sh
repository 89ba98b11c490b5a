#!/bin/bash
# lazy arrival waits of the fused gather: multirank parity, then 128 images per rank lazy vs per-step waits
o=gpurun_out; NP=${NP:-2}
[ $NP -le 4 ] && python -m pytest tests/test_gpu_multirank.py -q -m gpu 2>&1 | tail -3
run() { tag=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus $NP --global-batch $((128*NP)) --steps 400 --warmup 30 --no-e2e --no-objects --no-cpu-baseline $EXTRA > $o/r02w_$tag.json 2> $o/r02w_$tag.err || tail -5 $o/r02w_$tag.err
python -c "
import json;d=json.load(open('$o/r02w_$tag.json'))
print('$tag', round(d['value']), d['gather_bit_exact'], {m: (round(r['ms_per_step'],4), r['ms_per_step_by_rank']) for m,r in d['modes'].items()})"; }
run lazy X=1
EXTRA="--gather-wait step" run step X=1
[ $NP -ge 8 ] && EXTRA="--pipeline 12" run lazy_p12 X=1
true
