#!/bin/bash
# build_variant.sh <tag> <extra nvcc flags...>  -> csrc/liblanedet_<tag>.so (experiments; select with LANEDET_LIB)
tag=$1; shift
cd "$(dirname "$0")/../advanced-lane-detection_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-extended-lambda -Xcompiler -fPIC -shared "$@" -o liblanedet_$tag.so lanedet.cu
