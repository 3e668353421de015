#!/bin/sh
# Builds libepimcmc.so for sm_100a in-tree (the .so travels to the GPU box with the snapshot).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
for f in epi_kernels epi_capi epi_sampler; do
  $NVCC $FLAGS -c $f.cu -o $f.o &
done
wait
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o libepimcmc.so epi_kernels.o epi_capi.o epi_sampler.o -cudart static
echo "built $(pwd)/libepimcmc.so"
