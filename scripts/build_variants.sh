#!/bin/bash
# builds kernel variants into scripts/_bin/libbmb200_<name>.so for A/B runs on the GPU box (BMB200_LIB=...)
set -e
cd "$(dirname "$0")/.."
mkdir -p scripts/_bin
build() { name=$1; shift; nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared "$@" -o scripts/_bin/libbmb200_$name.so bitmagic_b200/csrc/capi.cu -lcudart & }
build ctas2
build ctas3 -DBMB200_CTAS_PER_SM=3
build ctas3u8 -DBMB200_CTAS_PER_SM=3 -DBMB200_BIT_UNROLL=8
build ctas3g32 -DBMB200_CTAS_PER_SM=3 -DBMB200_LANES_PER_BLOCK=32
wait
ls -la scripts/_bin/
