#!/bin/bash
# Development tool: build side-by-side variants of the CUDA library that differ in the -D switches of ONE source file.
# usage: tools/build_variants.sh <source.cu> name1 "-DA=1 -DB=0" name2 "-DA=0" ...   -> build/variants/libhb_<name>.so
set -e
cd "$(dirname "$0")/.."
SRC=$1; shift
OUT=build/variants; mkdir -p $OUT
python -m hobbit_b200.build > /dev/null
OBJS=$(ls hobbit_b200/build/*.o | grep -v "/${SRC%.cu}.o")
while [ $# -gt 0 ]; do
  name=$1; defs=$2; shift 2
  ( nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -O3 $defs -c hobbit_b200/csrc/$SRC -o $OUT/${name}.o &&
    nvcc -shared -o $OUT/libhb_${name}.so $OUT/${name}.o $OBJS -gencode arch=compute_100a,code=sm_100a && echo "built $name ($defs)" ) &
done
wait
