#!/bin/bash
# full ncu capture of doc+term pass for the given kernel variants
mkdir -p gpurun_out
for V in "$@"; do
  ENSTOP_B200_VARIANT=$V timeout 900 ncu --set full --clock-control none --import-source on -k regex:row_pass -s 9 -c 2 \
    -o gpurun_out/prof_var$V -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --profile-iters 1 > gpurun_out/ncu_var$V.log 2>&1
  tail -1 gpurun_out/ncu_var$V.log | cut -c1-200
done
