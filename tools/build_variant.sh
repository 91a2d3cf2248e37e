#!/bin/bash
# Build a kernel variant of libhgi_b200.so for tools/ab_bench.py:  tools/build_variant.sh NAME [-DFLAG ...]
# Recompiles the tile kernels with the extra flags into build/NAME/ and links them with the stock objects.
set -e
cd "$(dirname "$0")/../rustyhgi_b200/csrc"
name=$1; shift
out=../../build/$name; mkdir -p $out
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -fmad=false"
for f in hgi_tile_fast hgi_tile_tma; do $NV "$@" -c -o $out/$f.o $f.cu & done; wait
$NV -shared -cudart static -o ../../build/libhgi_$name.so $out/hgi_tile_fast.o $out/hgi_tile_tma.o hgi_capi.o hgi_tile_kernels.o hgi_level_kernels.o hgi_reduce_kernels.o hgi_archive.o -lz
echo built build/libhgi_$name.so
