#!/bin/bash
# Build tuning variants of libsvb200 into gpurun_variants/ (shipped to the GPU box, not committed).
# usage: tools/build_variants.sh "name:-DFLAG=1 -DOTHER=2" ...
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  objs=""
  for f in supervillain_b200/csrc/*.cu; do
    o="variants/${name}_$(basename ${f%.cu}).o"
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Iinclude -Isupervillain_b200/csrc $flags -c $f -o $o &
    objs="$objs $o"
  done
  wait
  nvcc -Wno-deprecated-gpu-targets -shared -o variants/libsvb200_${name}.so $objs -lcudart_static -lpthread -ldl -lrt
  rm -f $objs
  echo "built variants/libsvb200_${name}.so ($flags)"
done
