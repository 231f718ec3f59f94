#!/bin/bash
# Build the sm_100a shared library in-tree (cimrgp_b200/libcimrgp.so).  nvcc cross-compiles without a GPU.
# Two translation units compiled in parallel: the C ABI with the streaming / small-matrix kernels (cimrgp.cu) and
# the fused ci sweep (chain.cu).
set -e
cd "$(dirname "$0")"
SRC=cimrgp_b200/csrc
OBJ=build/obj
mkdir -p $OBJ
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v"
nvcc $FLAGS -c -o $OBJ/cimrgp.o $SRC/cimrgp.cu 2> $SRC/build.log &
P1=$!
nvcc $FLAGS -c -o $OBJ/chain.o $SRC/chain.cu 2> $SRC/build_chain.log &
P2=$!
wait $P1 || { cat $SRC/build.log | grep -v "^ptxas info" | head -50; exit 1; }
wait $P2 || { cat $SRC/build_chain.log | grep -v "^ptxas info" | head -50; exit 1; }
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o cimrgp_b200/libcimrgp.so $OBJ/cimrgp.o $OBJ/chain.o
