#!/bin/bash
# 2 GPUs, 128 images per rank (the 8-GPU shard): where does the multi-GPU step time go
o=gpurun_out
run() { tag=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus 2 --global-batch 256 --steps 300 --warmup 30 --no-e2e --no-objects $EXTRA > $o/r02o_$tag.json 2> $o/r02o_$tag.err || tail -5 $o/r02o_$tag.err
python -c "
import json;d=json.load(open('$o/r02o_$tag.json'))
print('$tag', round(d['value']), round(d['ms_per_step'],4), d['gather_bit_exact'], {m: (round(r['ms_per_step'],4), {k: round(v,4) for k,v in r['kernel_ms'].items()}) for m,r in d['modes'].items()})"; }
run fused X=1
EXTRA="--gather nccl" run nccl X=1
run tf0 SDNET_DECODE_LIB=structuredetector_b200/csrc/exp/lib_tf0.so
run tf2 SDNET_DECODE_LIB=structuredetector_b200/csrc/exp/lib_tf2.so
EXTRA="--pipeline 8" run pipe8 X=1
EXTRA="--pipeline 4" run pipe4 X=1
python bench.py --global-batch 128 --no-e2e --no-cpu-baseline --no-objects --no-parity --steps 300 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('1 GPU x 128', {m: (round(r['value']), round(r['ms_per_step'],4)) for m,r in d['modes'].items()})"
