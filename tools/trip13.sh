#!/bin/bash
o=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=15 > $o/r02l_pytest.log 2>&1; tail -4 $o/r02l_pytest.log
for dt in f16 bf16; do timeout 600 python tools/sweep.py base --images 1024,128 --dtype $dt --out $o/r02l_$dt.json 2>&1 | tail -4; done
