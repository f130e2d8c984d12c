#!/bin/bash
# builds kernel variants into scripts/_bin/libbmb200_<name>.so for A/B runs on the GPU box (BMB200_LIB=...)
set -e
cd "$(dirname "$0")/.."
mkdir -p scripts/_bin
rm -f scripts/_bin/*.so
build() { name=$1; shift; nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared "$@" -o scripts/_bin/libbmb200_$name.so bitmagic_b200/csrc/capi.cu -lcudart -ldl & }
build narrow -DBMB200_AGG_CHUNK_WIDE=1024
build narrow_2slot -DBMB200_AGG_CHUNK_WIDE=1024 -DBMB200_FLAT_SLOTS=2
wait
ls -la scripts/_bin/
