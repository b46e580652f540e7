#!/bin/sh
# Lab build of the library (-DEXB_LAB enables the NTT experiment switches); rebuild normally afterwards.
cd "$(dirname "$0")/../exacto_b200/csrc" && nvcc -shared -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -DEXB_LAB $EXB_LAB_FLAGS kernels.cu rns_kernels.cu api.cu host_setup.cpp -o ../libexacto_b200.so
