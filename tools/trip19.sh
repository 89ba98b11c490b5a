#!/bin/bash
# N GPUs (default 2), 128 images per rank: completion flags vs barrier, multicast vs peer stores
o=gpurun_out; NP=${NP:-2}
python -m pytest tests/test_gpu_multirank.py -q -m gpu 2>&1 | tail -3
run() { tag=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus $NP --global-batch $((128*NP)) --steps 300 --warmup 30 --no-e2e --no-objects --no-cpu-baseline $EXTRA > $o/r02r_$tag.json 2> $o/r02r_$tag.err || tail -5 $o/r02r_$tag.err
python -c "
import json;d=json.load(open('$o/r02r_$tag.json'))
print('$tag', round(d['value']), round(d['ms_per_step'],4), d['gather_bit_exact'], {m: (round(r['ms_per_step'],4), {k: round(v,4) for k,v in r['kernel_ms'].items()}) for m,r in d['modes'].items()})"; }
run flags X=1
run barrier SDNET_GATHER_SYNC=barrier
run flags_peer SDNET_GATHER_STORES=peer
EXTRA=--no-parity run solo SDNET_GATHER_DIAG=local,nobarrier
