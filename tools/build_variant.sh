#!/bin/bash
# build_variant.sh <tag> [nvcc -D flags...]: csrc/liblanedet_<tag>.so for A/B timing through LANEDET_LIB (tools/gpu_ab.sh)
tag=$1; shift
cd "$(dirname "$0")/../advanced-lane-detection_b200/csrc" || exit 1
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-extended-lambda -Xcompiler -fPIC -shared "$@" \
    -o liblanedet_$tag.so lanedet.cu
