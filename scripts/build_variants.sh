#!/bin/bash
# Build library variants for A/B runs on the GPU box:
#   scripts/build_variants.sh "name1:-DFLAG1 -DFLAG2" "name2:-DFLAG3" ...
# -> hermespy-rt_b200/build/var_<name>/libhermespy_rt.so (select with HRT_LIB=...)
set -e
cd "$(dirname "$0")/../hermespy-rt_b200"
make build/compute_paths.o build/scene.o build/materials.o build/host_math.o build/hrt_multi.o > /dev/null
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  d=build/var_$name; mkdir -p $d
  ( nvcc -O3 -std=c++17 $flags -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false \
      -Xcompiler -fPIC,-ffp-contract=off -Xptxas -v -c csrc/hrt_cuda.cu -o $d/hrt_cuda.o 2> $d/ptxas.log &&
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $d/libhermespy_rt.so $d/hrt_cuda.o \
      build/hrt_multi.o build/compute_paths.o build/scene.o build/materials.o build/host_math.o -lm -ldl &&
    echo "built $d ($flags)" ) &
done
wait
