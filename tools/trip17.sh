#!/bin/bash
# 2 GPUs, 128 images per rank: multicast stores vs peer stores, 1-warp CTAs, gather diagnostics
o=gpurun_out
run() { tag=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NP:-2} --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus ${NP:-2} --global-batch $((128*${NP:-2})) --steps 300 --warmup 30 --no-e2e --no-objects $EXTRA > $o/r02p_$tag.json 2> $o/r02p_$tag.err || tail -5 $o/r02p_$tag.err
python -c "
import json;d=json.load(open('$o/r02p_$tag.json'))
print('$tag', round(d['value']), round(d['ms_per_step'],4), d['gather_bit_exact'], d['config']['gather'][:90], {m: (round(r['ms_per_step'],4), {k: round(v,4) for k,v in r['kernel_ms'].items()}) for m,r in d['modes'].items()})"; }
run auto X=1
run peer SDNET_GATHER_STORES=peer
run w1 SDNET_DECODE_LIB=structuredetector_b200/csrc/exp/lib_w1.so
EXTRA=--no-parity run local SDNET_GATHER_DIAG=local
EXTRA=--no-parity run nobar SDNET_GATHER_DIAG=nobarrier
for lib in "" structuredetector_b200/csrc/exp/lib_w1.so; do
SDNET_DECODE_LIB=$lib python bench.py --global-batch 128 --no-e2e --no-cpu-baseline --no-objects --no-parity --steps 300 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('1 GPU x 128 [$lib]', {m: (round(r['value']), round(r['ms_per_step'],4), round(r['kernel_ms']['peaks'],4)) for m,r in d['modes'].items()})"
SDNET_DECODE_LIB=$lib python bench.py --no-e2e --no-cpu-baseline --no-objects --no-parity --steps 100 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('1 GPU x 1024 [$lib]', {m: (round(r['value']), round(r['ms_per_step'],4), round(r['kernel_ms']['peaks'],4)) for m,r in d['modes'].items()})"
done
python -m pytest tests/test_gpu_multirank.py -q -m gpu 2>&1 | tail -3
