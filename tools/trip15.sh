#!/bin/bash
o=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=15 > $o/r02n_pytest.log 2>&1; tail -6 $o/r02n_pytest.log
timeout 900 python tools/sweep.py base s0 s1 --images 1024,128 --out $o/r02n_sweep.json 2>&1 | tee $o/r02n_sweep.log | tail -14
for dt in f16 bf16; do timeout 600 python tools/sweep.py base s0 --images 1024 --dtype $dt --out $o/r02n_$dt.json 2>&1 | tail -4; done
