#!/bin/bash
# compute-sanitizer over the small parity cases: memcheck (the row pass reads ahead
# of an item's end into zeroed padding; the tiled pass indexes shared memory by corpus data) and
# racecheck (last-arrival reductions, shared-memory tile staging).  Logs -> gpurun_out/sanitizer/.
mkdir -p gpurun_out/sanitizer
SEL='small_cases or split_rows or fused_loglik or device_plan'
for TILED in 0 1; do
  for TOOL in memcheck racecheck; do
    LOG=gpurun_out/sanitizer/${TOOL}_tiled${TILED}.log
    ENSTOP_B200_TILED=$TILED timeout 1500 compute-sanitizer --tool $TOOL --error-exitcode 9 \
      python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" > $LOG 2>&1
    echo "$TOOL tiled=$TILED rc=$?" | tee -a gpurun_out/sanitizer/summary.txt
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $LOG | tail -3 | tee -a gpurun_out/sanitizer/summary.txt
  done
done
