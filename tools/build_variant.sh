#!/bin/bash
# usage: tools/build_variant.sh NAME [-DMACRO=VALUE ...]   -> rkstiff_b200/variants/NAME.so (same ABI; select it with RKS_LIB=...)
set -e
NAME=$1; shift
mkdir -p rkstiff_b200/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared "$@" \
     -o rkstiff_b200/variants/$NAME.so rkstiff_b200/csrc/*.cu
echo "built rkstiff_b200/variants/$NAME.so"
