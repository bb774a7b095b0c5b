#!/bin/bash
# Build a tuning variant of the library: scripts/build_variant.sh NAME -DPSA_FFT_STORE64=0 ...
# -> build/libpsa_NAME.so (git-ignored; travels to the GPU box), selected with PSA_B200_LIB=build/libpsa_NAME.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build
cd psa_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared "$@" \
  -o ../../build/libpsa_$name.so api.cu ingest.cu phase.cu project_simt.cu project_tc2.cu fft.cu fft4.cu post.cu dump.cu -lcudart
echo "built build/libpsa_$name.so ($*)"
