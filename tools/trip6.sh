#!/bin/bash
o=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=15 > $o/r02f_pytest.log 2>&1; tail -5 $o/r02f_pytest.log
timeout 1500 python tools/sweep.py base q0 d0 ng4c5 ng4c5q0 d24 f8 --images 1024,128 --out $o/r02f_sweep.json 2>&1 | tee $o/r02f_sweep.log | tail -40
timeout 600 python tools/sweep.py base --images 1024 --dtype f16 --out $o/r02f_f16.json 2>&1 | tail -3
timeout 600 python tools/sweep.py base --images 1024 --dtype bf16 --out $o/r02f_bf16.json 2>&1 | tail -3
