#!/bin/bash
# tools/build_variant.sh NAME [-DSTEP_SEG=.. -DSTEP_NST=..]  ->  gpurun_out/lib_NAME.so  (kernel-parameter experiments)
set -e
name=$1; shift
cd "$(dirname "$0")/../sampling_gpmpc_b200/csrc"
mkdir -p ../../gpurun_out/variants
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -shared -Xcompiler -fPIC "$@" -o ../../variants_$name.so gpmpc_api.cu
