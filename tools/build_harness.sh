#!/bin/bash
# Build the standalone bring-up harnesses (not part of the library): tools/bin/tc3_test, tools/bin/tcwg_test
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/bin
for t in ${@:-tc3_test tcwg_test}; do
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -DS2S_KERNEL_IMPL -Iinclude -o tools/bin/$t tools/$t.cu -lcuda
done
