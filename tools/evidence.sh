#!/bin/bash
# Round evidence on one B200 (run under gpurun): parity suite, the bench lines kept in profiles/, the
# ncu launch list and one full capture per kernel.  usage: bash tools/evidence.sh <round tag, e.g. r01>
tag=${1:-r01}; o=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $o/${tag}_pytest_gpu.log 2>&1; tail -3 $o/${tag}_pytest_gpu.log
python bench.py > $o/${tag}_bench_n1.json 2> $o/${tag}_bench_n1.err || tail -5 $o/${tag}_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > $o/${tag}_bench_ref.json 2> $o/${tag}_bench_ref.err || tail -5 $o/${tag}_bench_ref.err
python bench.py --mode blobs --no-cpu-baseline > $o/${tag}_bench_n1_blobs.json 2>/dev/null
python bench.py --pipeline 1 --no-cpu-baseline --no-e2e > $o/${tag}_bench_n1_serial.json 2>/dev/null
python bench.py --dtype f16 --no-cpu-baseline > $o/${tag}_bench_n1_f16.json 2>/dev/null
python bench.py --dtype bf16 --no-cpu-baseline > $o/${tag}_bench_n1_bf16.json 2>/dev/null
cmd="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --pipeline 1"
$cmd > $o/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_launches_n1.csv $cmd > $o/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sdnet_ -s 9 -c 3 -o $o/${tag}_kernels_n1 $cmd > $o/ncu_full.log 2>&1; tail -1 $o/ncu_full.log
for f in n1 n1_blobs n1_serial n1_f16 n1_bf16; do python -c "
import json;d=json.load(open('$o/${tag}_bench_$f.json'));print('$f', round(d['value']), round(d['ms_per_step'],4), {k: round(v,4) for k,v in d['roofline']['kernel_ms'].items()}, 'frac %.3f' % d['roofline']['frac'], d['clocks']['reasons'], (d.get('e2e') or {}).get('value'))"; done
python -c "
import json;d=json.load(open('$o/${tag}_bench_ref.json'));print('ref', d['value'], d['cpu_baseline'])"
