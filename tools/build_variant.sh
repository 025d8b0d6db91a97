#!/bin/sh
# build_variant.sh NAME [-DFLAG ...] -- an experimental build of the library as tools/bin/libq_NAME.so (select it with
# QPSK_B200_LIB=$PWD/tools/bin/libq_NAME.so); the host objects are the ones of the normal build.
set -e
cd "$(dirname "$0")/../qpsk_b200"
name=$1; shift
mkdir -p ../tools/bin
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC "$@" -c -o ../tools/bin/q_$name.o csrc/qpsk_b200.cu
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../tools/bin/libq_$name.so ../tools/bin/q_$name.o host/host_design.o host/dropin.o host/stream.o -lm -lpthread
rm -f ../tools/bin/q_$name.o
