#!/bin/bash
o=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=15 > $o/r02b_pytest.log 2>&1; tail -8 $o/r02b_pytest.log
timeout 1500 python tools/sweep.py base lateh f32 ng3c6 ng3c6f8 ng3c6f32 ng4c6 --out $o/r02b_sweep.json 2>&1 | tee $o/r02b_sweep.log | tail -40
timeout 900 python tools/sweep.py base base@SDNET_CHUNK_GROUPS=16 base@SDNET_CHUNK_GROUPS=24 base@SDNET_CHUNK_GROUPS=32 base@SDNET_CHUNK_GROUPS=43 base@SDNET_CHUNK_GROUPS=64 ng3c6@SDNET_CHUNK_GROUPS=32 --images 128,256 --out $o/r02b_chunks.json 2>&1 | tee $o/r02b_chunks.log | tail -40
