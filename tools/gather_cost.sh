#!/bin/bash
# N GPUs (default 2), 128 images per rank: what the gather costs now (thin exact select, self-cleaning workspace)
o=gpurun_out; NP=${NP:-2}
run() { tag=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus $NP --global-batch $((128*NP)) --steps 400 --warmup 30 --no-e2e --no-objects --no-cpu-baseline --no-parity $EXTRA > $o/r02v_$tag.json 2> $o/r02v_$tag.err || tail -5 $o/r02v_$tag.err
python -c "
import json;d=json.load(open('$o/r02v_$tag.json'))
print('$tag', round(d['value']), {m: (round(r['ms_per_step'],4), {k: round(v,4) for k,v in r['kernel_ms'].items()}) for m,r in d['modes'].items()})"; }
run solo SDNET_GATHER_DIAG=local,nobarrier
run stores_only SDNET_GATHER_DIAG=nobarrier
run flags X=1
run barrier SDNET_GATHER_SYNC=barrier
EXTRA="--gather nccl" run nccl X=1
