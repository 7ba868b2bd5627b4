#!/bin/bash
# Builds experimental variants of libvbc.so (one per tuning-knob setting of csrc/spmv.cu) into
# build/variants/libvbc_<tag>.so for tools/tune.py.  The product library is csrc/Makefile's.
set -e
cd "$(dirname "$0")/../sparsematrixvbcs.jl_b200/csrc"
OUT=../../build/variants
mkdir -p $OUT
ARCH="-gencode arch=compute_100a,code=sm_100a"
FLAGS="-O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC"
build() { # tag, defines...
  tag=$1; shift
  ( nvcc $FLAGS "$@" -c spmv.cu -o $OUT/spmv_$tag.o -Xptxas -v 2> $OUT/spmv_$tag.ptxas.log
    nvcc $ARCH -shared -o $OUT/libvbc_$tag.so api.o pack.o csc.o peer.o $OUT/spmv_$tag.o -lcudart ) &
}
make -j4 >/dev/null
build base
build unr2 -DVBC_ADJ_UNR=2
build unr8 -DVBC_ADJ_UNR=8
build minb1 -DVBC_ADJ_MINB=1
build minb4 -DVBC_ADJ_MINB=4
build minb6 -DVBC_ADJ_MINB=6
wait
build minb8_unr2 -DVBC_ADJ_MINB=8 -DVBC_ADJ_UNR=2
build ldg -DVBC_LD_MODE=1
build ldna -DVBC_LD_MODE=2
build wide -DVBC_WIDE_LD=1
build wide_minb4 -DVBC_WIDE_LD=1 -DVBC_ADJ_MINB=4
build wide_unr2 -DVBC_WIDE_LD=1 -DVBC_ADJ_UNR=2
wait
build minb6_unr2 -DVBC_ADJ_MINB=6 -DVBC_ADJ_UNR=2
build wide_unr2_minb5 -DVBC_WIDE_LD=1 -DVBC_ADJ_UNR=2 -DVBC_ADJ_MINB=5
wait
ls -la $OUT/*.so | wc -l
