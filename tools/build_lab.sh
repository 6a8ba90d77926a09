#!/bin/sh
# Builds the kernel lab (developer tool).  Usage: tools/build_lab.sh && gpurun -- tools/kernel_lab
set -e
cd "$(dirname "$0")/.."
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false \
     -o tools/kernel_lab tools/kernel_lab.cu "$@"
