#!/bin/bash
o=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=15 -x > $o/r02j_pytest.log 2>&1; tail -4 $o/r02j_pytest.log
timeout 900 python tools/sweep.py base prev wp0 f24 f32 --images 1024,128 --out $o/r02j_sweep.json 2>&1 | tee $o/r02j_sweep.log | tail -30
for args in "128 noise" "1024 noise"; do python tools/trace_warps.py $args 2>&1 | tail -9; done
