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
build m4u4 -DVBC_ADJ_MINB=4 -DVBC_ADJ_UNR=4
build m4u8 -DVBC_ADJ_MINB=4 -DVBC_ADJ_UNR=8
build m4u2 -DVBC_ADJ_MINB=4 -DVBC_ADJ_UNR=2
build m3u4 -DVBC_ADJ_MINB=3 -DVBC_ADJ_UNR=4
build m3u8 -DVBC_ADJ_MINB=3 -DVBC_ADJ_UNR=8
wait
build m5u4 -DVBC_ADJ_MINB=5 -DVBC_ADJ_UNR=4
build m5u2 -DVBC_ADJ_MINB=5 -DVBC_ADJ_UNR=2
build m6u2 -DVBC_ADJ_MINB=6 -DVBC_ADJ_UNR=2
build wm4u4 -DVBC_WIDE_LD=1 -DVBC_ADJ_MINB=4 -DVBC_ADJ_UNR=4
build wm4u2 -DVBC_WIDE_LD=1 -DVBC_ADJ_MINB=4 -DVBC_ADJ_UNR=2
build wm3u4 -DVBC_WIDE_LD=1 -DVBC_ADJ_MINB=3 -DVBC_ADJ_UNR=4
wait
ls -la $OUT/*.so | wc -l
