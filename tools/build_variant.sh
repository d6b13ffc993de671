#!/bin/bash
# A/B builds of one translation unit: tools/build_variant.sh <name> <file.cu> <nvcc -D flags...>
#   -> dl-dkd_b200/variants/libdkd_b200_<name>.so (the other objects come from dl-dkd_b200/build/)
set -e
cd "$(dirname "$0")/../dl-dkd_b200"
name=$1; src=$2; shift; shift
base=$(basename $src .cu)
mkdir -p variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c csrc/$src -o build/${base}_$name.o
objs=$(ls build/*.o | grep -v "build/${base}" | grep -v "_[a-z]\.o$")
nvcc --shared -o variants/libdkd_b200_$name.so $objs build/${base}_$name.o
echo variants/libdkd_b200_$name.so
