#!/bin/bash
# usage: tools/build_variant.sh NAME -DKNOB=.. ... -> meanfieldvargp_b200/lib/libmfvgp_NAME.so (select with MFVGP_LIB)
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -Xptxas -v "$@" -o meanfieldvargp_b200/lib/libmfvgp_$name.so meanfieldvargp_b200/csrc/mfvgp.cu 2>&1 | grep -A2 "walk_kernelILi128" | tail -2
