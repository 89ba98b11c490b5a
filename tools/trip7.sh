#!/bin/bash
o=gpurun_out
timeout 600 python -m pytest tests/test_gpu_rawdecoder.py -m gpu -q > $o/r02g_pytest.log 2>&1; tail -5 $o/r02g_pytest.log
python tools/time_suppress.py 256 2>&1 | tail -4
SDNET_SUPPRESS_PATH=w python tools/time_suppress.py 256 2>&1 | tail -4
cmd="python bench.py --steps 2 --warmup 3 --mode blobs --no-e2e --no-cpu-baseline --no-objects --no-parity --pipeline 1"
$cmd > $o/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sdnet_peaks\|sdnet_tail -s 8 -c 2 -o $o/r02g_blobs $cmd > $o/ncu.log 2>&1; tail -2 $o/ncu.log
