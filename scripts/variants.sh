#!/bin/bash
# A/B builds of libtoygpu.so with extra -D flags: scripts/variants.sh name1 "-DFLAG1" name2 "-DFLAG2" ...
# -> toycluster_b200/variants/libtoygpu_<name>.so ; time them on the GPU with scripts/time_variants.py
set -e
cd "$(dirname "$0")/../toycluster_b200/csrc"
mkdir -p ../variants
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xcompiler -fopenmp $flags \
       -shared -Xlinker -soname=libtoygpu.so -o ../variants/libtoygpu_$name.so toygpu.cu -lgomp &
done
wait
ls -la ../variants
