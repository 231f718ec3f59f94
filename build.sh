#!/bin/bash
# Build the sm_100a shared library in-tree (cimrgp_b200/libcimrgp.so).  nvcc cross-compiles without a GPU.
set -e
cd "$(dirname "$0")"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared \
     -Xptxas -v -o cimrgp_b200/libcimrgp.so cimrgp_b200/csrc/cimrgp.cu 2> cimrgp_b200/csrc/build.log
