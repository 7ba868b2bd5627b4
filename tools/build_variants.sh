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
    nvcc $ARCH -shared -o $OUT/libvbc_$tag.so api.o pack.o csc.o peer.o spmm.o trsv.o fwdt.o mixed.o hostdp.o $OUT/spmv_$tag.o -lcudart ) &
}
make -j4 >/dev/null
if [ -z "$SPMM_VARIANTS" ]; then
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
fi
# SpMM ring geometry (options 4 / 5 of VBC_OPT_SPMM_SIMT): rows per stage, stages, CTAs per SM
build_spmm() { # tag, defines...
  tag=$1; shift
  ( nvcc $FLAGS "$@" -c spmm.cu -o $OUT/spmm_$tag.o -Xptxas -v 2> $OUT/spmm_$tag.ptxas.log
    nvcc $ARCH -shared -o $OUT/libvbc_spmm_$tag.so api.o pack.o csc.o peer.o spmv.o trsv.o fwdt.o mixed.o hostdp.o $OUT/spmm_$tag.o -lcudart ) &
}
if [ -n "$SPMM_VARIANTS" ]; then
  build_spmm c16s2m2 -DVBC_TMA_CH=16 -DVBC_TMA_ST=2 -DVBC_TMA_MINB=2
  build_spmm c8s2m3 -DVBC_TMA_CH=8 -DVBC_TMA_ST=2 -DVBC_TMA_MINB=3
  build_spmm c8s3m3 -DVBC_TMA_CH=8 -DVBC_TMA_ST=3 -DVBC_TMA_MINB=3
  build_spmm c8s2m4 -DVBC_TMA_CH=8 -DVBC_TMA_ST=2 -DVBC_TMA_MINB=4
  build_spmm c4s4m4 -DVBC_TMA_CH=4 -DVBC_TMA_ST=4 -DVBC_TMA_MINB=4
  wait
fi
ls -la $OUT/*.so | wc -l
