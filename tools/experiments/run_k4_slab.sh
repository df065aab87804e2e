#!/bin/bash
# Builds (nvcc cross-compiles without a GPU) and, with RUN=1 on a GPU box, runs the K4 forward A/B experiment.
set -e
cd "$(dirname "$0")/../.."
mkdir -p build gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 tools/experiments/k4_slab.cu -o build/k4_slab \
     -Lhs-pose_b200/lib -lhspose_b200 -Xlinker -rpath -Xlinker "$PWD/hs-pose_b200/lib"
if [ "${RUN:-0}" = "1" ]; then ./build/k4_slab | tee gpurun_out/k4_slab.jsonl; fi
