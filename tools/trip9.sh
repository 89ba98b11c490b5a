#!/bin/bash
o=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=15 > $o/r02i_pytest.log 2>&1; tail -8 $o/r02i_pytest.log
timeout 900 python tools/sweep.py base base@SDNET_CHUNK_GROUPS=64 base@SDNET_CHUNK_GROUPS=42 base@SDNET_CHUNK_GROUPS=32 base@SDNET_CHUNK_GROUPS=28 base@SDNET_CHUNK_GROUPS=21 --images 128,256 --out $o/r02i_chunks.json 2>&1 | tee $o/r02i_chunks.log | tail -30
python tools/time_suppress.py 256 2>&1 | tail -3
for args in "128 noise" "128 blobs" "128 noise SDNET_CHUNK_GROUPS=42" "128 noise SDNET_CHUNK_GROUPS=28" "1024 noise"; do python tools/trace_warps.py $args 2>&1 | tail -9; done
