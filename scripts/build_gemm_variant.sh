#!/bin/bash
# Builds eigenkernel_b200/libekb200_<tag>.so: the library with the GEMM engine compiled for another pipeline geometry
# (experiments only; select it with EKB200_LIB=<path>).  Usage: scripts/build_gemm_variant.sh <tag> <BK> <STAGES>
set -e
cd "$(dirname "$0")/../eigenkernel_b200/csrc"
make -j16 > /dev/null
TAG=$1; BK=$2; ST=$3
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DEKB_GEMM_BK=$BK -DEKB_GEMM_STAGES=$ST -c gemm.cu -o build/gemm_$TAG.o
OBJS=$(ls build/*.o | grep -v "build/gemm" | tr '\n' ' ')
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libekb200_$TAG.so $OBJS build/gemm_$TAG.o -lcudart
ls -la ../libekb200_$TAG.so
