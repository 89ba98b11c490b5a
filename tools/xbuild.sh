#!/bin/bash
# Build a macro variant of the decode library for kernel experiments:
#   bash tools/xbuild.sh <name> -DSDNET_X_WAIT=1 ...   ->  structuredetector_b200/csrc/exp/lib_<name>.so
# and run with SDNET_DECODE_LIB=structuredetector_b200/csrc/exp/lib_<name>.so
name=$1; shift
mkdir -p structuredetector_b200/csrc/exp
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -prec-div=true -prec-sqrt=true -ftz=false \
  -Xcompiler -fPIC -shared -Iinclude "$@" -o structuredetector_b200/csrc/exp/lib_$name.so structuredetector_b200/csrc/sdnet_decode.cu
