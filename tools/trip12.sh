#!/bin/bash
# 2 GPUs: multi-rank parity tests + the bench at N=2
o=gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q > $o/r02k_pytest_multirank.log 2>&1; tail -5 $o/r02k_pytest_multirank.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 20 > $o/r02k_bench_n2.json 2> $o/r02k_bench_n2.err || tail -20 $o/r02k_bench_n2.err
python -c "
import json;d=json.load(open('$o/r02k_bench_n2.json'))
print('N=2 value', round(d['value']), d['ms_per_step'], 'parity', d['parity_checked'], 'gather_bit_exact', d['gather_bit_exact'], d['config']['gather'][:40])
for m,r in d['modes'].items(): print(m, round(r['value']), r['ms_per_step'], r['kernel_ms'], r['gather_bit_exact'])
"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 200 --warmup 20 --gather nccl --no-e2e > $o/r02k_bench_n2_nccl.json 2> $o/r02k_bench_n2_nccl.err || tail -20 $o/r02k_bench_n2_nccl.err
python -c "
import json;d=json.load(open('$o/r02k_bench_n2_nccl.json'))
print('N=2 nccl value', round(d['value']), d['ms_per_step'], 'gather_bit_exact', d['gather_bit_exact'])"
