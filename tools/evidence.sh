#!/bin/bash
# Round evidence on one B200 (run under gpurun): parity suite, the bench lines kept in profiles/, the ncu launch list and
# full captures of every kernel.  usage: bash tools/evidence.sh <round tag, e.g. r02>
tag=${1:-r02}; o=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $o/${tag}_pytest_gpu.log 2>&1; tail -3 $o/${tag}_pytest_gpu.log
python bench.py > $o/${tag}_bench_n1.json 2> $o/${tag}_bench_n1.err || tail -5 $o/${tag}_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > $o/${tag}_bench_ref.json 2> $o/${tag}_bench_ref.err || tail -5 $o/${tag}_bench_ref.err
python bench.py --pipeline 1 --no-cpu-baseline --no-e2e --no-objects > $o/${tag}_bench_n1_serial.json 2>/dev/null
for w in cfg2 cfg3 cfg4; do python bench.py --workload $w --steps 300 > $o/${tag}_bench_$w.json 2>/dev/null; done
for d in f16 bf16; do python bench.py --dtype $d --no-cpu-baseline --no-objects > $o/${tag}_bench_n1_$d.json 2>/dev/null; done
python bench.py --global-batch 128 --no-cpu-baseline --no-e2e --no-objects > $o/${tag}_bench_n1_gb128.json 2>/dev/null
for f in n1 n1_serial cfg2 cfg3 cfg4 n1_f16 n1_bf16 n1_gb128; do python -c "
import json;d=json.load(open('$o/${tag}_bench_$f.json'));print('$f', round(d['value']), round(d['ms_per_step'],4), 'parity', d['parity_checked'], {m: (round(r['value']), {k: round(v,4) for k,v in r['kernel_ms'].items()}, round(r['peaks_frac'],3)) for m,r in d['modes'].items()}, d['clocks']['reasons'], 'e2e', (d.get('e2e') or {}).get('value'), (d.get('cpu_baseline') or {}).get('value'))"; done
python -c "
import json;d=json.load(open('$o/${tag}_bench_ref.json'));print('ref', d['value'], d['cpu_baseline'])"
python tools/time_suppress.py 256 > $o/${tag}_suppress.log 2>&1; cat $o/${tag}_suppress.log
# ---- ncu: the launch list, then one full capture per kernel and input regime (each after the same command ran clean)
base="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-objects --no-parity --pipeline 1"
cmd="$base --mode noise"
$cmd > $o/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_launches_n1.csv $cmd > $o/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sdnet_ -s 9 -c 3 -o $o/${tag}_kernels_n1 $cmd > $o/ncu_full.log 2>&1; tail -1 $o/ncu_full.log
cmd="$base --mode blobs"
$cmd > $o/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:sdnet_ -s 9 -c 3 -o $o/${tag}_kernels_n1_blobs $cmd > $o/ncu_full.log 2>&1; tail -1 $o/ncu_full.log
cmd="$base --mode blobs --global-batch 128"
$cmd > $o/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:sdnet_ -s 9 -c 3 -o $o/${tag}_kernels_n1_gb128 $cmd > $o/ncu_full.log 2>&1; tail -1 $o/ncu_full.log
for d in f16 bf16; do cmd="$base --mode noise --dtype $d"
$cmd > $o/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:sdnet_peaks -s 3 -c 1 -o $o/${tag}_peaks_$d $cmd > $o/ncu_full.log 2>&1; tail -1 $o/ncu_full.log; done
cmd="python tools/time_suppress.py 256"
$cmd > $o/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:sdnet_suppress -s 2 -c 1 -o $o/${tag}_suppress $cmd > $o/ncu_full.log 2>&1; tail -1 $o/ncu_full.log
