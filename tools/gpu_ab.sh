#!/bin/bash
# A/B timing of library variants built by tools/build_variant.sh: gpu_ab.sh "<stages>" tag1 tag2 ...
stages=$1; shift
mkdir -p gpurun_out
echo "base"; timeout 120 python tools/stage_times.py --frames 1260 --only "$stages"
for t in "$@"; do
  echo "variant $t"; LANEDET_LIB=$PWD/advanced-lane-detection_b200/csrc/liblanedet_$t.so timeout 120 python tools/stage_times.py --frames 1260 --only "$stages"
done 2>&1 | tee gpurun_out/ab.txt
