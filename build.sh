#!/bin/bash
# Build every CUDA extension for sm_100a (cross-compiles without a GPU) and the CPU oracle.
set -e
cd "$(dirname "$0")"
mkdir -p vpp_b200/lib oracle/_build
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false \
     -Xcompiler -fPIC -shared ${VPPB_NVCC_EXTRA} \
     -o vpp_b200/lib/libvppb.so vpp_b200/csrc/*.cu
gcc -O2 -ffp-contract=off -fno-fast-math -fPIC -shared -o oracle/_build/libvpp_oracle.so oracle/*.c -lm
gcc -O3 -march=native -fopenmp -DNDEBUG -ffp-contract=off -fPIC -shared -o oracle/_build/libvpp_oracle_omp.so oracle/*.c -lm
echo "built vpp_b200/lib/libvppb.so oracle/_build/libvpp_oracle{,_omp}.so"
