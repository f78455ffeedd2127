#!/bin/bash
# Compile the reference's own headers (where they lie under /root/reference) against the Eigen / iod
# stand-ins of this directory.  Outputs only into oracle/_ref/ (git-ignored, travels to the GPU box).
#   libvppref.so      parity build: serial, -O2, no FP contraction, baseline x86-64 (scalar FAST fallback)
#   libvppref_omp.so  timing build: the reference's benchmark flags (benchmarks/CMakeLists.txt:10,18) -> AVX2 FAST path
set -e
cd "$(dirname "$0")/../.."
mkdir -p oracle/_ref
g++ -std=c++14 -O2 -ffp-contract=off -fno-fast-math -fPIC -shared -w -I oracle/ref_shim -I /root/reference \
    -o oracle/_ref/libvppref.so oracle/ref_shim/vppref.cc
g++ -std=c++14 -O3 -march=native -fopenmp -DNDEBUG -ffp-contract=off -fPIC -shared -w -I oracle/ref_shim -I /root/reference \
    -o oracle/_ref/libvppref_omp.so oracle/ref_shim/vppref.cc
echo "built oracle/_ref/libvppref{,_omp}.so from /root/reference"
