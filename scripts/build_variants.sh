#!/bin/bash
# Experimental builds of the library for A/B runs on the GPU box (scripts/gpu_ab.sh picks them
# up as LIB=<tag> -> build/libplsa_<tag>.so; build/ is git-ignored but travels with gpurun).
#   ftz   -DPLSA_EXP_FTZ_THRESH=1   E-step threshold by flush-to-zero scaling (plsa_kernels.cuh)
set -e
cd "$(dirname "$0")/.."
mkdir -p build
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -shared -Xcompiler -fPIC,-fvisibility=hidden,-mavx2 -diag-suppress 550"
SRC="enstop_b200/csrc/plsa_b200.cu enstop_b200/csrc/host_init.cpp"
nvcc $FLAGS -DPLSA_EXP_FTZ_THRESH=1 -o build/libplsa_ftz.so $SRC -ldl
ls -l build/*.so
