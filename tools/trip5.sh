#!/bin/bash
o=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=15 > $o/r02e_pytest.log 2>&1; tail -5 $o/r02e_pytest.log
timeout 1500 python tools/sweep.py base b0 b0d0 b1 b4 d12 d24 ng3c6 ng3c6b0d0 --images 1024,128 --out $o/r02e_sweep.json 2>&1 | tee $o/r02e_sweep.log | tail -40
