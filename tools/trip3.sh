#!/bin/bash
o=gpurun_out
cmd="python bench.py --steps 2 --warmup 3 --mode noise --no-e2e --no-cpu-baseline --no-objects --no-parity --pipeline 1"
$cmd > $o/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sdnet_peaks -s 4 -c 1 -o $o/r02c_peaks $cmd > $o/ncu.log 2>&1; tail -2 $o/ncu.log
