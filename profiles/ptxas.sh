#!/bin/bash
# registers / spills of every kernel of the library (cross-compiles here, no GPU needed)
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xptxas -v -Xcompiler -fPIC -shared \
  -I include -I shapes_b200/csrc -o /tmp/shapes_ptxas.so shapes_b200/csrc/shapes_b200.cu -ldl "$@" 2>&1 \
 | awk '/Compiling entry function/ {name=$0} /bytes stack frame/ {stack=$0} /Used [0-9]+ registers/ {print name " | " stack " | " $0}' \
 | sed -E "s/ptxas info    : Compiling entry function '([^']*)' for 'sm_100a'/\1/; s/ptxas info    : //g" | c++filt | grep -v "cub::" | sed -E 's/\(anonymous namespace\):://g'
