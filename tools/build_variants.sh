#!/bin/bash
# development aid: build kernel variants of libpioran_b200.so into build_abl/ (they travel to the GPU box, git ignores them)
# usage: bash tools/build_variants.sh name1="-DFLAG=1 -DX=2" name2="..."
mkdir -p build_abl
for spec in "$@"; do
  name="${spec%%=*}"; flags="${spec#*=}"
  ( nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared $flags \
      -o build_abl/lib_$name.so pioran.jl_b200/csrc/api.cu 2> build_abl/$name.log && echo "built $name" || echo "FAILED $name" ) &
done
wait
