#!/bin/sh
# Builds the kernel labs (developer tools).
#   tools/build_lab.sh            tools/kernel_lab  (walk formulations, gather ceiling, staged window)
#   tools/build_lab.sh replay     tools/replay_lab  (recording walk and replay kernel variants)
# Run on a B200: gpurun -- tools/kernel_lab | tools/replay_lab
set -e
cd "$(dirname "$0")/.."
which=kernel_lab
if [ "$1" = "replay" ]; then which=replay_lab; shift; fi
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false \
     -o tools/$which tools/$which.cu "$@"
