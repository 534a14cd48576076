#!/bin/bash
# build_variant.sh NAME [-DFLAG ...]: experiment build of the library into lib/exp_NAME.so (use with LBM2D_LIB)
name=$1; shift
cd "$(dirname "$0")/.."
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared --expt-relaxed-constexpr \
  -ccbin /usr/bin/g++ "$@" -o 01-lbm-2d_b200/lib/exp_$name.so 01-lbm-2d_b200/csrc/lbm2d_capi.cu -ldl -lpthread
