#!/bin/bash
# experimental build of the same ABI with different compile-time options of ONE source (default sm100_fused2.cu):
#   tools/build_variant.sh t4e8 "-DF2_TEAM_WARPS=4 -DF2_EPI_WARPS=8"   ->  vision-xai-breast-cancer-cad_b200/libbcad_t4e8.so  (use with BCAD_LIB=...)
set -e
name=$1; flags=$2; src=${3:-sm100_fused2.cu}
cd "$(dirname "$0")/../vision-xai-breast-cancer-cad_b200"
python build.py > /dev/null
mkdir -p build/var
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr \
     $flags -c csrc/$src -o build/var/${name}.o
objs=$(ls build/*.o | grep -v "/${src%.cu}.o")
nvcc -shared -o libbcad_${name}.so $objs build/var/${name}.o -gencode arch=compute_100a,code=sm_100a
echo "$PWD/libbcad_${name}.so"
