#!/bin/bash
# Experimental builds of the library for A/B runs on the GPU box (scripts/gpu_ab.sh picks them
# up as LIB=<tag> -> build/libplsa_<tag>.so; build/ is git-ignored but travels with gpurun).
#   ftz        -DPLSA_EXP_FTZ_THRESH=1   E-step threshold by flush-to-zero scaling (plsa_kernels.cuh)
#   ftz128_9   the same in 128-thread CTAs, 9 per SM: 56 registers, 36 resident warps instead of
#              32, no spill inside the entry loop (prologue / epilogue: 16-40 bytes)
#   d128_9     default threshold code in 128-thread CTAs, 9 per SM
set -e
cd "$(dirname "$0")/.."
mkdir -p build
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -shared -Xcompiler -fPIC,-fvisibility=hidden,-mavx2 -diag-suppress 550"
SRC="enstop_b200/csrc/plsa_b200.cu enstop_b200/csrc/host_init.cpp"
nvcc $FLAGS -DPLSA_EXP_FTZ_THRESH=1 -o build/libplsa_ftz.so $SRC -ldl &
nvcc $FLAGS -DPLSA_EXP_FTZ_THRESH=1 -DPLSA_PASS_THREADS=128 -DPLSA_PASS_MIN_CTAS=9 -o build/libplsa_ftz128_9.so $SRC -ldl &
nvcc $FLAGS -DPLSA_PASS_THREADS=128 -DPLSA_PASS_MIN_CTAS=9 -o build/libplsa_d128_9.so $SRC -ldl &
wait
ls -l build/*.so
