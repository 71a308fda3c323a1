#!/bin/bash
# experimental build of the same ABI with the cycle-trace accumulators compiled in (-DFZ_TRACE -DF2_TRACE):
#   tools/build_trace.sh && BCAD_LIB=$PWD/vision-xai-breast-cancer-cad_b200/libbcad_trace.so BCAD_DEBUG_SKIP_STORES=64 python bench.py --only-value ...
set -e
cd "$(dirname "$0")/../vision-xai-breast-cancer-cad_b200"
mkdir -p build/trace
pids=()
for f in csrc/*.cu; do
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr \
       -DFZ_TRACE -DF2_TRACE -c "$f" -o "build/trace/$(basename "${f%.cu}").o" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
nvcc -shared -o libbcad_trace.so build/trace/*.o -gencode arch=compute_100a,code=sm_100a
echo "$PWD/libbcad_trace.so"
