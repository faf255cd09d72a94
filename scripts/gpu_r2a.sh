#!/bin/bash
# Round 2, first GPU pass: (1) staging microbenchmark (tile / LDGSTS ring / cp.async.bulk ring /
# red.v4), (2) the A/Bs left over from round 1 on C2 (item order, flush-to-zero threshold,
# 128 x 9 CTAs).  Build first: bash scripts/build_variants.sh; nvcc ... microbench_stage.cu
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt
timeout 300 ./build/microbench_stage > gpurun_out/r2_microbench_stage.txt 2>&1
cat gpurun_out/r2_microbench_stage.txt
bash scripts/gpu_ab.sh C2 "d 1 0 1 0" "d 1 0 1 1" "d 1 0 1 2" "ftz 1 0 1 0" "ftz128_9 1 0 1 0" "d128_9 1 0 1 0" "d 1 0 1 0" "ftz 1 0 1 1" 2>&1 | tee gpurun_out/r2a_ab.txt
ENSTOP_B200_LIB=$PWD/build/libplsa_ftz.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_ftz.log
tail -5 gpurun_out/pytest_ftz.log
