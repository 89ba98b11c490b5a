#!/bin/bash
# one 8-GPU box: the scaling line at N = 8 and 4, the 4-rank parity test
o=gpurun_out
nvidia-smi -L | wc -l
for n in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+n)) bench.py --gpus $n --steps 300 --warmup 30 > $o/r02m_bench_n$n.json 2> $o/r02m_bench_n$n.err || tail -20 $o/r02m_bench_n$n.err
python -c "
import json;d=json.load(open('$o/r02m_bench_n$n.json'))
print('N=$n value', round(d['value']), d['ms_per_step'], 'parity', d['parity_checked'], 'gather_bit_exact', d['gather_bit_exact'], 'e2e', (d['e2e'] or {}).get('value'))
for m,r in d['modes'].items(): print(' ', m, round(r['value']), r['ms_per_step'], r['kernel_ms'], r['gather_bit_exact'])
"
done
timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -q > $o/r02m_pytest_multirank.log 2>&1; tail -3 $o/r02m_pytest_multirank.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus 8 --steps 300 --warmup 30 --gather nccl --no-e2e --no-parity > $o/r02m_bench_n8_nccl.json 2> $o/r02m_bench_n8_nccl.err
python -c "
import json;d=json.load(open('$o/r02m_bench_n8_nccl.json'));print('N=8 nccl', round(d['value']), d['ms_per_step'])"
