#!/bin/bash
# N GPUs (default 8), 128 images per rank: where does the gather's cost come from
o=gpurun_out; NP=${NP:-8}
run() { tag=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus $NP --global-batch $((128*NP)) --steps 300 --warmup 30 --no-e2e --no-objects --no-cpu-baseline $EXTRA > $o/r02q_$tag.json 2> $o/r02q_$tag.err || tail -5 $o/r02q_$tag.err
python -c "
import json;d=json.load(open('$o/r02q_$tag.json'))
print('$tag', round(d['value']), round(d['ms_per_step'],4), d['gather_bit_exact'], {m: (round(r['ms_per_step'],4), r['ms_per_step_by_rank'], {k: round(v,4) for k,v in r['kernel_ms'].items()}) for m,r in d['modes'].items()})"; }
run auto X=1
run peer SDNET_GATHER_STORES=peer
EXTRA=--no-parity run local SDNET_GATHER_DIAG=local
EXTRA=--no-parity run nobar SDNET_GATHER_DIAG=nobarrier
EXTRA=--no-parity run solo SDNET_GATHER_DIAG=local,nobarrier
EXTRA="--no-parity --pipeline 8" run auto_p8 X=1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29531 tools/host_rate.py 128 6 2>&1 | grep "img x" | sort | head -20
