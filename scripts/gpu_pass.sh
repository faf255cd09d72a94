#!/bin/bash
# GPU pass: parity tests, bench, launch list, full ncu capture of the three row-pass modes.
# usage: bash scripts/gpu_pass.sh <tag>
TAG=${1:-run}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/pytest_$TAG.log
tail -3 gpurun_out/pytest_$TAG.log
timeout 600 python bench.py --steps 100 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
cat gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-iters 2 \
    > gpurun_out/ncu_launches_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:row_pass -s 9 -c 3 \
    -o gpurun_out/prof_$TAG python bench.py --steps 3 --warmup 3 --no-cpu-baseline --profile-iters 1 \
    > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log | cut -c1-300
