#!/bin/bash
# Builds experimental variants of libvbc.so (one per compile-time knob setting of one source file) into
# build/variants/libvbc_<tag>.so; run them with VBC_LIBRARY=<path>.  The product library is csrc/Makefile's.
#   tools/build_variants.sh <file.cu> <tag> <defines...> [-- <tag> <defines...>]...
set -e
cd "$(dirname "$0")/../sparsematrixvbcs.jl_b200/csrc"
OUT=../../build/variants
mkdir -p $OUT
ARCH="-gencode arch=compute_100a,code=sm_100a"
FLAGS="-O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC"
make -j8 >/dev/null
SRC=$1; shift
BASE=${SRC%.cu}
OBJS=$(ls *.o | grep -v "^$BASE.o$" | tr '\n' ' ')
while [ $# -gt 0 ]; do
  tag=$1; shift
  defs=()
  while [ $# -gt 0 ] && [ "$1" != "--" ]; do defs+=("$1"); shift; done
  [ "$1" == "--" ] && shift
  ( nvcc $FLAGS "${defs[@]}" -c $SRC -o $OUT/${BASE}_$tag.o -Xptxas -v 2> $OUT/${BASE}_$tag.ptxas.log
    nvcc $ARCH -shared -o $OUT/libvbc_$tag.so $OBJS $OUT/${BASE}_$tag.o -lcudart -ldl ) &
done
wait
ls -la $OUT/*.so
