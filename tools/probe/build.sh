#!/bin/bash
# builds the probes next to their sources (binaries are git-ignored, travel with gpurun)
cd "$(dirname "$0")"
for f in *.cu; do
  nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -o "${f%.cu}.bin" "$f" || exit 1
done
