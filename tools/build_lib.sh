#!/bin/bash
# development helper: build libjme_b200.so (same flags as __graft_entry__.build) with ptxas statistics
cd "$(dirname "$0")/../jmolecularenergy_b200" || exit 1
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-ffp-contract=off -shared \
     -Xptxas -v "$@" -o libjme_b200.so csrc/jme_api.cu csrc/jme_host.cpp 2>&1
