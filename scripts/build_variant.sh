#!/bin/bash
# usage: scripts/build_variant.sh [extra nvcc flags]   (rebuilds the CUDA library in place)
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -shared -Xcompiler -fPIC,-ffp-contract=off "$@" -o cityseer_b200/libcityseer_b200.so cityseer_b200/csrc/cs_api.cu
