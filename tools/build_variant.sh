#!/bin/bash
# A/B builds of the exact kernel: tools/build_variant.sh <name> <nvcc -D flags...>  ->  gpurun_out/../dl-dkd_b200/build/libdkd_b200_<name>.so
set -e
cd "$(dirname "$0")/../dl-dkd_b200"
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c csrc/dkd_exact_umma.cu -o build/dkd_exact_umma_$name.o
objs=$(ls build/*.o | grep -v dkd_exact_umma)
nvcc --shared -o variants/libdkd_b200_$name.so $objs build/dkd_exact_umma_$name.o
echo variants/libdkd_b200_$name.so
