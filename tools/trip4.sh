#!/bin/bash
o=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=15 > $o/r02d_pytest.log 2>&1; tail -8 $o/r02d_pytest.log
timeout 1500 python tools/sweep.py base d0 d4 d16 ng3c6 ng3c6d0 ng3c6d4 --out $o/r02d_sweep.json 2>&1 | tee $o/r02d_sweep.log | tail -40
