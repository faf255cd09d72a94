#!/bin/bash
# First GPU pass: parity tests, a short bench, the launch list and one full ncu capture.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 50 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
cat gpurun_out/bench_c2.json; tail -5 gpurun_out/bench_c2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-iters 2 \
    > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:row_pass -s 10 -c 2 \
    -o gpurun_out/prof_rowpass python bench.py --steps 3 --warmup 3 --no-cpu-baseline --profile-iters 1 \
    > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
